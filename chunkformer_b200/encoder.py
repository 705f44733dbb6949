"""Host-side mirror of the reference encoder interface for the hot path.

`ChunkFormerEncoderB200` exposes the methods the reference's facade calls on `model.encoder`
(chunkformer/modules/encoder.py): `forward_parallel_chunk` (:503-681) and `forward_encoder` (:220-274), with the same
argument names, return tuples and error behaviour, and runs them on the sm_100a kernels through the C ABI
(include/chunkformer_b200.h).  PyTorch is only the owner of device memory and streams here.
"""
import ctypes
from ctypes import POINTER, c_int64, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .geometry import EncoderGeometry
from .plan import Plan

_EMPTY = (0,)


class UploadedFeatures:
    """A batch of utterances on its way to the device (ChunkFormerEncoderB200.upload_async): the flat feature buffer, the
    utterance lengths and the copy events.  Pass it to forward_parallel_chunk in place of the list of tensors."""

    def __init__(self, flat, lens, events, rows_ready):
        self.flat, self.lens, self.events, self.rows_ready = flat, lens, events, rows_ready

    def __len__(self):
        return len(self.lens)


class ChunkFormerEncoderB200:
    """B200 encoder + CTC head built from a reference-layout state_dict (keys `encoder.*`, `ctc.ctc_lo.*`)."""

    def __init__(self, geometry: EncoderGeometry, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        self.geo = geometry
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ChunkFormerEncoderB200 runs on CUDA devices only (no CPU fallback)")
        self._L = _lib.load()
        cfg = _lib.CfConfig(geometry.d_model, geometry.heads, geometry.ffn, geometry.layers, geometry.kernel,
                            geometry.vocab, geometry.feat_dim, 1 if geometry.has_cmvn else 0,
                            1 if geometry.conv_norm == "batch_norm" else 0)
        h = c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self._L.cf_create(ctypes.byref(cfg), int(idx), ctypes.byref(h)), None, "cf_create")
        self._h = h
        for key, t in state_dict.items():
            if not (key.startswith("encoder.") or key.startswith("ctc.ctc_lo.")):
                continue
            t = t.detach().to("cpu", torch.float32).contiguous()
            shape = (c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(self._L.cf_load_tensor(self._h, key.encode(), c_void_p(t.data_ptr()), _lib.CF_F32, t.dim(), shape),
                       self._h, "cf_load_tensor")
        _lib.check(self._L.cf_finalize_weights(self._h), self._h, "cf_finalize_weights")
        self._ws: Optional[torch.Tensor] = None
        self._ctc_ws: Optional[torch.Tensor] = None
        self._last_out = None        # (fp32 output, its bf16 twin) of the latest encode: ctc_greedy(out) then needs no cast
        self._copy_stream = None
        self._live_events = []
        # attributes the reference facade reads (chunkformer_model.py:344-389)
        self.num_blocks = geometry.layers
        self.attention_heads = geometry.heads
        self._output_size = geometry.d_model
        self.cnn_module_kernel = geometry.kernel
        self.subsampling_rate = 8

    def output_size(self) -> int:
        return self._output_size

    def set_option(self, name: str, value: int) -> None:
        """Per-handle tuning knob of the library (cf_set_option), e.g. "fused_layernorm" 0 / 1 for A/B measurements."""
        _lib.check(self._L.cf_set_option(self._h, name.encode(), int(value)), self._h, "cf_set_option")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.cf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, plan: Plan) -> torch.Tensor:
        need = int(self._L.cf_workspace_bytes(self._h, plan.handle))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def upload_async(self, xs: Sequence[torch.Tensor], lens: Optional[Sequence[int]] = None) -> UploadedFeatures:
        """Start copying ragged utterances into one flat device buffer [sum T_i, feat] and return at once.

        The copies run on a side stream, in pieces of about 32 MB, each followed by an event; forward_parallel_chunk hands the
        events to the library (cf_encode_feature_events), whose front-end waits only for the rows each slab of chunks reads, so
        the encoder starts on the first rows while the rest is still crossing PCIe (pinned host tensors copy asynchronously;
        pageable ones are staged by the driver): what the reference's blocking xs.to(device), chunkformer_model.py:395-401,
        cannot do.  The buffer is allocated from the side stream's pool, so an upload never waits for encoder work queued
        earlier.  (Starting the upload of batch k + 1 before batch k is encoded is possible with this call, but measured slower
        than uploading inside the step: the copy then competes with the attention / FFN kernels instead of hiding under the
        front-end; profiles/README.md.)"""
        lens = [int(x.shape[0]) for x in xs] if lens is None else [int(t) for t in lens]
        F = self.geo.feat_dim
        for x, t in zip(xs, lens):
            if x.dim() != 2 or x.shape[1] != F or x.shape[0] < t or t < 0:
                raise ValueError("every utterance must be (T_i, feat_dim) with T_i >= xs_origin_lens[i] >= 0")
        total = sum(lens)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        piece = max(1, (32 << 20) // (4 * F))    # rows per piece
        rows_ready, events = [], []
        row = 0
        with torch.cuda.stream(cs):
            flat = torch.empty((total, F), dtype=torch.float32, device=self.device)   # a block of the copy stream's pool
            for x, t in zip(xs, lens):
                if x.dtype != torch.float32:
                    x = x.float()
                for a in range(0, t, piece):
                    b = min(t, a + piece)
                    flat[row + a:row + b].copy_(x[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    events.append(ev)
                    rows_ready.append(row + b)
                row += t
        return UploadedFeatures(flat, lens, events, rows_ready)

    def _arm_feature_events(self, up: UploadedFeatures) -> torch.Tensor:
        """Hand the copy events of `up` to the library for the cf_encode call that follows, and tie the buffer's lifetime to
        the consuming stream."""
        up.flat.record_stream(torch.cuda.current_stream(self.device))
        self._live_events = up.events            # keep the events alive until the next call
        n = len(up.events)
        if n:
            rr = (c_int64 * n)(*up.rows_ready)
            evp = (c_void_p * n)(*[c_void_p(e.cuda_event) for e in up.events])
            _lib.check(self._L.cf_encode_feature_events(self._h, n, rr, evp), self._h, "cf_encode_feature_events")
        return up.flat

    def encode_plan(self, plan: Plan, feats: torch.Tensor, att_cache=None, cnn_cache=None, trunc: int = 0,
                    out_dtype=torch.float32, want_bf16: bool = False, workspace: Optional[torch.Tensor] = None):
        """Run cf_encode on a prepared plan and a flat device feature buffer. Returns (out (rows, d), out_bf16|None).
        workspace: a caller-owned uint8 device buffer of cf_workspace_bytes (StreamingGraph pins its plan's tables in one)."""
        d = self.geo.d_model
        out = torch.empty((plan.rows, d), dtype=out_dtype, device=self.device)
        out16 = torch.empty((plan.rows, d), dtype=torch.bfloat16, device=self.device) \
            if (want_bf16 and out_dtype != torch.bfloat16) else None
        ws = self._workspace(plan) if workspace is None else workspace
        rc = self._L.cf_encode(self._h, plan.handle, c_void_p(feats.data_ptr()), _lib.ptr(att_cache), _lib.ptr(cnn_cache),
                               int(trunc), c_void_p(out.data_ptr()),
                               _lib.CF_F32 if out_dtype == torch.float32 else _lib.CF_BF16, _lib.ptr(out16),
                               c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        _lib.check(rc, self._h, "cf_encode")
        self._last_out = (out, out16) if out16 is not None else None
        return out, (out if out_dtype == torch.bfloat16 else out16)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_parallel_chunk(self, xs, xs_origin_lens: torch.Tensor, chunk_size: int = -1,
                               left_context_size: int = -1, right_context_size: int = -1,
                               att_cache: torch.Tensor = torch.zeros((0, 0, 0)),
                               cnn_cache: torch.Tensor = torch.zeros((0, 0)), truncated_context_size: int = 0,
                               offset: torch.Tensor = torch.zeros(0)
                               ) -> Tuple[torch.Tensor, torch.Tensor, List[int], torch.Tensor, torch.Tensor, torch.Tensor]:
        """Masked-batch encoder forward; drop-in for ChunkFormerEncoder.forward_parallel_chunk (encoder.py:503-681).

        xs: list of (T_i, feat) tensors (host or device), or the UploadedFeatures of an earlier upload_async call.
        Returns (xs (n, c, d) fp32, xs_lens int32 (B), n_chunks list, att_cache, cnn_cache, offset); like the reference,
        `offset` is advanced in place and the caches come back empty ((L,0,0,0) / (L,0,0)) when none were passed."""
        if chunk_size <= 0 or left_context_size < 0 or right_context_size < 0:
            raise ValueError("forward_parallel_chunk needs chunk_size > 0 and non-negative context sizes")
        lens = [int(v) for v in xs_origin_lens.tolist()]
        if len(lens) != len(xs):
            raise ValueError("xs and xs_origin_lens disagree on the batch size")
        if isinstance(xs, UploadedFeatures):
            if xs.lens != lens:
                raise ValueError("xs_origin_lens differs from the lengths the features were uploaded with")
        else:
            for x, t in zip(xs, lens):
                if x.dim() != 2 or x.shape[1] != self.geo.feat_dim or x.shape[0] < t or t < 0:
                    raise ValueError("every utterance must be (T_i, feat_dim) with T_i >= xs_origin_lens[i] >= 0")
        if offset.shape[0] == 0:
            offset = torch.zeros(len(xs), dtype=torch.long, device=xs_origin_lens.device)
        offs = [int(v) for v in offset.tolist()]
        L, H, d, lo = self.geo.layers, self.geo.heads, self.geo.d_model, self.geo.kernel // 2
        # the reference hands cache slice i to layer i whenever the cache tensors have a leading dimension
        # (encoder.py:655-661); with left_context_size == 0 the K/V part is empty but the conv cache still carries over
        streaming = (att_cache.dim() == 4 and att_cache.size(0) > 0) or (cnn_cache.dim() == 3 and cnn_cache.size(0) > 0)
        new_att = new_cnn = None
        if streaming:
            if len(xs) != 1:
                raise ValueError("streaming caches need a single utterance per call")
            if tuple(cnn_cache.shape) != (L, d, lo) or \
                    (left_context_size > 0 and tuple(att_cache.shape) != (L, left_context_size, H, 2 * d // H)):
                raise ValueError("cache shapes must be (L, left_context, H, 2*d_k) and (L, d, kernel//2)")
            if left_context_size > 0:
                new_att = att_cache.to(self.device, torch.float32).contiguous().clone()
            else:                                # no K/V rows to carry; the library still wants a (dummy) buffer
                new_att = torch.zeros((L, 1, H, 2 * d // H), device=self.device)
            new_cnn = cnn_cache.to(self.device, torch.float32).contiguous().clone()
        plan = Plan(chunk_size, left_context_size, right_context_size, lens, offs, self.geo.kernel)
        # everything is validated: start (or pick up) the upload and arm its events for the cf_encode call that follows
        feats = self._arm_feature_events(xs if isinstance(xs, UploadedFeatures) else self.upload_async(xs, lens))
        out, _ = self.encode_plan(plan, feats, new_att, new_cnn, truncated_context_size, want_bf16=True)
        xs_lens = torch.as_tensor(plan.enc_lens, dtype=torch.int32, device=xs_origin_lens.device)
        offset += xs_lens.to(offset.dtype)
        if not streaming:
            new_att = torch.zeros((L, 0, 0, 0), device=self.device)
            new_cnn = torch.zeros((L, 0, 0), device=self.device)
        elif left_context_size == 0:
            new_att = torch.zeros((L, 0, H, 2 * d // H), device=self.device)
        return out.view(plan.n, chunk_size, d), xs_lens, plan.n_chunks, new_att, new_cnn, offset

    # ------------------------------------------------------------------------------------------------------------
    # frame-synchronous streaming
    @torch.no_grad()
    def forward_chunk(self, xs: torch.Tensor, att_cache: torch.Tensor = torch.zeros((0, 0, 0, 0, 0)),
                      cnn_cache: torch.Tensor = torch.zeros((0, 0, 0, 0)), chunk_size: int = 0, left_context_size: int = 0,
                      right_context_size: int = 0, offset: int = 0, donate_caches: bool = False):
        """One streaming step for B concurrent streams; drop-in for ChunkFormerEncoder.forward_chunk (encoder.py:310-390).

        xs (B, 8 (c + r - 1) + 15, feat) = the new c encoder frames of every stream plus r frames of right context; att_cache
        (L, B, H, l, 2 d_k) and cnn_cache (L, B, d, 7) as returned by the previous step (empty = zeros = start of the streams);
        offset = encoder frames already consumed.  Returns (out (B, c + r, d), ones mask (B, 1, c + r), new att_cache,
        new cnn_cache); the caller keeps the first c rows of every step but the last (forward_chunk_by_chunk does).

        A step is, per stream, the masked-chunk path on ONE chunk of c + r frames with that stream's caches: the reference
        embeds and attends chunk + right context as a single chunk with left context l (encoder.py:341-347), cuts the conv
        module at the chunk grid (chunk_size = c inside convolution.py:150-167) and ends the returned caches where the chunk
        ends (encoder.py:376-385); oracle.forward_chunk restates it, pinned against the reference (tests/golden/stream.npz,
        stream_right.npz).  All B streams share one encoder pass (cf_encode_streams): every stream carries its left context
        as placeholder rows that cf_encode fills from the caches.  With r = 0 (every shipped streaming preset,
        apps/realtime-asr/config.py:86-110) the tcgen05 attention kernels run; chunk + r is in general not a tile-friendly
        size, those steps take the CUDA-core attention fallback.

        donate_caches=True: the passed device fp32 caches are updated in place and returned (a serving loop hands the returned
        caches straight back in, so the copies the reference's functional style implies, 1 GB per step for CTC-large at 256
        streams, are pure waste); the default keeps the reference's behaviour of leaving the arguments untouched."""
        c, l, r = int(chunk_size), int(left_context_size), int(right_context_size)
        if c <= 0 or r < 0 or l < self.geo.kernel // 2 or xs.dim() != 3:
            raise ValueError("forward_chunk needs xs (B, T, feat), chunk_size > 0, right_context_size >= 0 and "
                             "left_context_size >= kernel // 2")
        B, T, _ = xs.shape
        cc = c + r                                   # frames embedded and attended as one chunk
        if T != 8 * (cc - 1) + 15:
            raise ValueError(f"forward_chunk expects {8 * (cc - 1) + 15} input frames per step for chunk_size {c} and "
                             f"right_context_size {r}, got {T}")
        L, H, d, lo = self.geo.layers, self.geo.heads, self.geo.d_model, self.geo.kernel // 2
        if att_cache.numel() == 0:
            att_cache = torch.zeros((L, B, H, l, 2 * d // H), device=self.device)
        if cnn_cache.numel() == 0:
            cnn_cache = torch.zeros((L, B, d, lo), device=self.device)
        if tuple(att_cache.shape) != (L, B, H, l, 2 * d // H) or tuple(cnn_cache.shape) != (L, B, d, lo):
            raise ValueError("cache shapes must be (L, B, H, left_context, 2*d_k) and (L, B, d, kernel//2)")
        # one pass for all streams: stream s = an utterance of `ph` placeholder chunks (zeros in, their K/V and conv rows are
        # overwritten with the stream's caches inside cf_encode) + the real chunk; the negative plan offset masks the part of
        # the left context that is not filled yet (valid cache rows = min(offset, l))
        ph = -(-max(l, lo) // cc)

        def own(t):
            if donate_caches and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous():
                return t
            return t.to(self.device, torch.float32).contiguous().clone()
        new_att, new_cnn = own(att_cache), own(cnn_cache)
        t_tot = ph * 8 * cc + T
        x_dev = torch.zeros((B, t_tot, xs.shape[2]), dtype=torch.float32, device=self.device)
        x_dev[:, ph * 8 * cc:] = xs.to(self.device, torch.float32)
        plan = Plan(cc, l, 0, [t_tot] * B, [-(ph * cc - min(int(offset), l))] * B, self.geo.kernel)
        _lib.check(self._L.cf_encode_streams(self._h, B, ph, c), self._h, "cf_encode_streams")
        o, _ = self.encode_plan(plan, x_dev.view(B * t_tot, -1), new_att, new_cnn, 0)
        if int(self._L.cf_encode_output_rows(self._h)) == B * cc:      # compact step: only the real chunks were computed
            out = o[:B * cc].view(B, cc, d)
        else:
            out = o.view(B, (ph + 1) * cc, d)[:, ph * cc:].contiguous()
        return out, torch.ones((B, 1, cc), dtype=torch.bool, device=self.device), new_att, new_cnn

    @torch.no_grad()
    def forward_chunk_by_chunk(self, xs: torch.Tensor, xs_lens: torch.Tensor, chunk_size: int = 0, left_context_size: int = 0,
                               right_context_size: int = 0):
        """Streaming simulation over whole utterances; drop-in for ChunkFormerEncoder.forward_chunk_by_chunk
        (encoder.py:392-459).  Returns (out (B, steps * c [+ r rows of the last step], d), masks (B, 1, T'))."""
        c, l, r = int(chunk_size), int(left_context_size), int(right_context_size)
        B, T, _ = xs.shape
        size, stride = 8 * (c - 1) + 15 + 8 * r, 8 * c
        pad = stride - ((T - size) % stride)
        xp = torch.nn.functional.pad(xs.to(self.device, torch.float32), (0, 0, 0, pad))
        att = torch.zeros((0, 0, 0, 0, 0))
        cnn = torch.zeros((0, 0, 0, 0))
        outs, offset = [], 0
        for i in range(0, xp.shape[1] - size + stride, stride):
            o, _, att, cnn = self.forward_chunk(xp[:, i:i + size], att, cnn, c, l, r, offset, donate_caches=True)
            outs.append(o[:, :c] if i + size < xp.shape[1] else o)      # encoder.py:449
            offset += c
        out = torch.cat(outs, dim=1)

        def calc(n):                                     # subsampling.py:270-288
            for _ in range(3):
                n = (n - 3) // 2 + 1
            return max(n, 0)
        enc_lens = torch.tensor([calc(int(t) + pad) for t in xs_lens.tolist()])
        masks = (torch.arange(int(enc_lens.max())).unsqueeze(0) < enc_lens.unsqueeze(1)).unsqueeze(1).to(self.device)
        return out, masks

    @torch.no_grad()
    def forward_encoder(self, xs: torch.Tensor, xs_lens: torch.Tensor, chunk_size: int = 0, left_context_size: int = 0,
                        right_context_size: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Padded-batch encoder forward; drop-in for ChunkFormerEncoder.forward_encoder (encoder.py:220-274).
        Returns (xs (B, T', d) fp32, masks (B, 1, T') bool)."""
        if xs.dim() != 3:
            raise ValueError("xs must be (B, T, feat)")
        B, T, _ = xs.shape
        Tp = (T - 15) // 8 + 1
        if chunk_size is None or chunk_size <= 0:
            # full attention = one chunk spanning the utterance (attention.py:411 / encoder.py:490-493)
            chunk_size, left_context_size, right_context_size = Tp, 0, 0
        lens = [int(v) for v in xs_lens.tolist()]
        plan = Plan(chunk_size, left_context_size, right_context_size, lens, None, self.geo.kernel, padded_T=T)
        feats = xs.to(self.device, torch.float32).contiguous().view(B * T, -1)
        out, _ = self.encode_plan(plan, feats)
        nck = plan.n_chunks[0]
        out = out.view(B, nck * chunk_size, self.geo.d_model)[:, :Tp]
        enc_lens = torch.as_tensor(plan.enc_lens, device=self.device).clamp_min(0)
        masks = (torch.arange(Tp, device=self.device).unsqueeze(0) < enc_lens.unsqueeze(1)).unsqueeze(1)
        return out, masks

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def fbank(self, waveform: torch.Tensor, num_mel_bins: int = 80, frame_length: int = 25, frame_shift: int = 10,
              sample_frequency: int = 16000) -> torch.Tensor:
        """Kaldi-compatible log-mel filterbank on the device; drop-in for the call the reference makes right before the path,
        torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=..., frame_length=..., frame_shift=..., dither=0.0,
        energy_floor=0.0, sample_frequency=...) (chunkformer_model.py:307-315).  waveform (1, n) or (n,) in 16-bit range."""
        w = waveform.reshape(-1).to(self.device, torch.float32).contiguous()
        T = int(self._L.cf_fbank_num_frames(w.numel(), int(sample_frequency), int(frame_length), int(frame_shift)))
        out = torch.empty((T, num_mel_bins), dtype=torch.float32, device=self.device)
        if T:
            _lib.check(self._L.cf_fbank(self._h, c_void_p(w.data_ptr()), w.numel(), int(sample_frequency), int(num_mel_bins),
                                        int(frame_length), int(frame_shift), c_void_p(out.data_ptr()), self._stream()),
                       self._h, "cf_fbank")
        return out

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def ctc_greedy(self, enc: torch.Tensor, want_margin: bool = False, want_logp: bool = False):
        """argmax(log_softmax(ctc_lo(enc))) (ctc.py:73-91). enc (..., d) fp32 or bf16 on the device.
        Returns tokens int64 (...,) [, margin fp32 (...,)] [, logp fp32 (..., V)]."""
        if self.geo.vocab <= 0:
            raise ValueError("model has no CTC head")
        lead = enc.shape[:-1]
        e = enc.reshape(-1, self.geo.d_model)
        if e.device != self.device:
            e = e.to(self.device)
        last = self._last_out
        if last is not None and e.dtype == torch.float32 and e.data_ptr() == last[0].data_ptr() and e.numel() == last[0].numel():
            e = last[1]                           # the bf16 twin cf_encode wrote next to this fp32 output
        if e.dtype not in (torch.float32, torch.bfloat16):
            e = e.float()
        e = e.contiguous()
        rows = e.shape[0]
        tokens = torch.empty(rows, dtype=torch.int64, device=self.device)
        margin = torch.empty(rows, dtype=torch.float32, device=self.device) if want_margin else None
        logp = torch.empty((rows, self.geo.vocab), dtype=torch.float32, device=self.device) if want_logp else None
        dt = _lib.CF_F32 if e.dtype == torch.float32 else _lib.CF_BF16      # fp32 rows are rounded to bf16 by the library's own kernel
        need = int(self._L.cf_ctc_workspace_bytes(self._h, rows, dt))
        if self._ctc_ws is None or self._ctc_ws.numel() < need:
            self._ctc_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        rc = self._L.cf_ctc_greedy(self._h, c_void_p(e.data_ptr()), dt, rows, c_void_p(tokens.data_ptr()), _lib.ptr(margin),
                                   _lib.ptr(logp), c_void_p(self._ctc_ws.data_ptr()), self._ctc_ws.numel(), self._stream())
        _lib.check(rc, self._h, "cf_ctc_greedy")
        res = [tokens.view(lead)]
        if want_margin:
            res.append(margin.view(lead))
        if want_logp:
            res.append(logp.view(*lead, self.geo.vocab))
        return res[0] if len(res) == 1 else tuple(res)

    @torch.no_grad()
    def ctc_prefix_beam_search(self, enc: torch.Tensor, enc_lens, beam_size: int = 10, blank_id: int = 0):
        """CTC prefix beam search on encoder outputs: log_softmax(ctc_lo(enc)) by the library's CTC kernels, the first beam prune
        (top-k per frame) on the device, the prefix search on the host (postprocess.prefix_beam_search_topk).  What
        ASRModel.decode(methods=["ctc_prefix_beam_search"]) does behind the encoder (modules/asr_model.py:313-325 ->
        modules/search.py:131-249).  enc: (B, T, d) fp32 / bf16 on the device (e.g. from forward_encoder); enc_lens: (B,) valid
        frames.  Returns one postprocess.DecodeResult per utterance (tokens, score, times, nbest, nbest_scores, nbest_times).
        One utterance at a time, so that only T x V log-probabilities exist at once."""
        from .postprocess import DecodeResult, prefix_beam_search_topk
        if enc.dim() != 3:
            raise ValueError("enc must be (B, T, d)")
        k = min(int(beam_size), self.geo.vocab)
        out = []
        for b in range(enc.shape[0]):
            n = int(enc_lens[b])
            if n <= 0:
                out.append(DecodeResult([], 0.0, [], [[]], [0.0], [[]]))
                continue
            _, logp = self.ctc_greedy(enc[b, :n], want_logp=True)
            top_logp, top_idx = logp.topk(k, dim=1)
            nbest, scores, times = prefix_beam_search_topk(top_logp.cpu().numpy(), top_idx.cpu().numpy(), n, k, blank_id)
            out.append(DecodeResult(nbest[0], scores[0], times[0], nbest, scores, times))
        return out

    @torch.no_grad()
    def ctc_compact(self, tokens: torch.Tensor, seg_start: Sequence[int], seg_len: Sequence[int], mode: int = 0,
                    blank_id: int = 0):
        """Device-side CTC compaction of greedy token ids (utils/model_utils.py:23-32, :186-196), so that only the kept
        tokens are copied to the host.  tokens: device int64, flat rows; utterance s owns rows [seg_start[s], seg_start[s] +
        seg_len[s]).  mode 0 = remove_duplicates_and_blank, mode 1 = drop blanks only.
        Returns one (token ids int64, frame indices int32) pair of host tensors per utterance."""
        t = tokens.reshape(-1).to(self.device, torch.int64).contiguous()
        rows, n = t.numel(), len(seg_start)
        st = torch.tensor([int(v) for v in seg_start], dtype=torch.int64).to(self.device, non_blocking=True)
        ln = torch.tensor([max(int(v), 0) for v in seg_len], dtype=torch.int32).to(self.device, non_blocking=True)
        out_tok = torch.empty(max(rows, 1), dtype=torch.int64, device=self.device)
        out_fr = torch.empty(max(rows, 1), dtype=torch.int32, device=self.device)
        offs = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        ws = torch.empty(int(self._L.cf_ctc_compact_workspace_bytes(rows)), dtype=torch.uint8, device=self.device)
        rc = self._L.cf_ctc_compact(c_void_p(t.data_ptr()), rows, c_void_p(st.data_ptr()), c_void_p(ln.data_ptr()), n, int(mode),
                                    int(blank_id), c_void_p(out_tok.data_ptr()), c_void_p(out_fr.data_ptr()),
                                    c_void_p(offs.data_ptr()), c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        _lib.check(rc, None, "cf_ctc_compact")
        o = offs.cpu().tolist()
        total = o[-1]
        tok_h, fr_h = out_tok[:total].cpu(), out_fr[:total].cpu()
        return [(tok_h[o[s]:o[s + 1]], fr_h[o[s]:o[s + 1]]) for s in range(n)]


class StreamingGraph:
    """Frame-synchronous streaming of B concurrent streams with the steady-state step captured in a CUDA graph.

    A streaming step of ChunkFormerEncoder.forward_chunk (encoder.py:310-390) is about 70 small launches per layer; below a few
    dozen streams its time is the launches (3.2 ms per step for CTC-large at B = 1 on a B200), not the work.  Once every stream's
    left context is filled (offset >= left_context_size) the step no longer changes: same plan, same buffers, caches updated in
    place.  This class runs the first steps through forward_chunk (same caches, donated) and then replays one captured graph
    of [copy the new frames in, cf_encode on the pinned plan (cf_plan_pin: no host-dependent operation), greedy CTC].
    Results are those of forward_chunk (tests/test_gpu_encoder.py::test_streaming_graph_equals_forward_chunk).

        sg = StreamingGraph(enc, B, chunk_size=16, left_context_size=64)
        for frames in source:                     # frames (B, 8 (c - 1) + 15, feat), host or device
            out, tokens = sg.step(frames)         # out (B, c, d) fp32, tokens (B, c) int64: views of static buffers

    right_context_size > 0 is not offered here (those steps take forward_chunk)."""

    def __init__(self, enc: "ChunkFormerEncoderB200", n_streams: int, chunk_size: int, left_context_size: int,
                 with_ctc: bool = True):
        c, l = int(chunk_size), int(left_context_size)
        geo = enc.geo
        lo = geo.kernel // 2
        if n_streams <= 0 or c <= 0 or l < lo:
            raise ValueError("StreamingGraph needs n_streams > 0, chunk_size > 0 and left_context_size >= kernel // 2")
        self.enc, self.B, self.c, self.l = enc, int(n_streams), c, l
        self.with_ctc = bool(with_ctc) and geo.vocab > 0
        self.T = 8 * (c - 1) + 15
        L, H, d = geo.layers, geo.heads, geo.d_model
        dev = enc.device
        self.att = torch.zeros((L, self.B, H, l, 2 * d // H), device=dev)
        self.cnn = torch.zeros((L, self.B, d, lo), device=dev)
        self.x_in = torch.zeros((self.B, self.T, geo.feat_dim), device=dev)
        self.offset = 0
        self._ph = -(-max(l, lo) // c)
        self._t_tot = self._ph * 8 * c + self.T
        self._x_dev = torch.zeros((self.B, self._t_tot, geo.feat_dim), device=dev)
        self._plan = Plan(c, l, 0, [self._t_tot] * self.B, [-(self._ph * c - l)] * self.B, geo.kernel)
        need = int(enc._L.cf_workspace_bytes(enc._h, self._plan.handle))
        self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._graph = None
        self._out = self._tok = None
        self._stream = torch.cuda.Stream(dev)

    def reset(self) -> None:
        """Start new streams: empty caches, offset 0 (the captured graph is kept)."""
        self.att.zero_()
        self.cnn.zero_()
        self.offset = 0

    def _steady_step(self):
        enc, B, c, d = self.enc, self.B, self.c, self.enc.geo.d_model
        self._x_dev[:, self._ph * 8 * c:] = self.x_in
        _lib.check(enc._L.cf_encode_streams(enc._h, B, self._ph, c), enc._h, "cf_encode_streams")
        o, _ = enc.encode_plan(self._plan, self._x_dev.view(B * self._t_tot, -1), self.att, self.cnn, 0, workspace=self._ws)
        if int(enc._L.cf_encode_output_rows(enc._h)) == B * c:           # compact step: only the real chunks were computed
            out = o[:B * c].view(B, c, d)
        else:
            out = o.view(B, (self._ph + 1) * c, d)[:, self._ph * c:].contiguous()
        tok = enc.ctc_greedy(out) if self.with_ctc else None
        return out, tok

    def _capture(self) -> None:
        enc = self.enc
        cur = torch.cuda.current_stream(enc.device)
        st = self._stream
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            _lib.check(enc._L.cf_plan_pin(enc._h, self._plan.handle, c_void_p(self._ws.data_ptr()), self._ws.numel(),
                                          c_void_p(st.cuda_stream)), enc._h, "cf_plan_pin")
            # one eager run on private copies of the caches: position tables, shared-memory opt-ins and workspaces exist afterwards
            keep = (self.att, self.cnn)
            self.att, self.cnn = self.att.clone(), self.cnn.clone()
            self._steady_step()
            self.att, self.cnn = keep
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            self._out, self._tok = self._steady_step()
        self._graph = g
        cur.wait_stream(st)

    @torch.no_grad()
    def step(self, xs: torch.Tensor):
        """Consume the next chunk of every stream: xs (B, 8 (c - 1) + 15, feat).  Returns (out (B, c, d), tokens (B, c) or None);
        both are overwritten by the next call."""
        if tuple(xs.shape) != (self.B, self.T, self.enc.geo.feat_dim):
            raise ValueError(f"StreamingGraph.step expects xs of shape {(self.B, self.T, self.enc.geo.feat_dim)}")
        if self.offset < self.l:
            # left context still filling: the plan masks the missing part, so these steps differ from one another
            out, _, self.att, self.cnn = self.enc.forward_chunk(xs, self.att, self.cnn, self.c, self.l, 0, self.offset,
                                                                donate_caches=True)
            self.offset += self.c
            return out, (self.enc.ctc_greedy(out) if self.with_ctc else None)
        if self._graph is None:
            self._capture()
        self.x_in.copy_(xs, non_blocking=True)
        self._graph.replay()
        self.offset += self.c
        return self._out, self._tok

