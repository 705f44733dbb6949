"""Transducer greedy search on the device: host-side mirror of the reference's search interface for `chunkformer-rnnt-*`
models (chunkformer/transducer/search/greedy_search.py: `optimized_search`, `batch_greedy_search`), bound to
`cf_rnnt_*` of the C ABI (include/chunkformer_b200.h).  PyTorch only owns the device buffers.
"""
import ctypes
from ctypes import POINTER, c_int32, c_int64, c_void_p
from typing import Dict, List, Sequence, Tuple

import torch

from . import lib as _lib


class RnntConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("vocab", "embed", "hidden", "layers", "pred_out", "enc_dim", "join_dim", "blank")]


class TransducerGreedyB200:
    """LSTM predictor + joint + greedy control built from a reference-layout state_dict (keys `predictor.*`, `joint.*`).

    Supported: `predictor: rnn` with `rnn_type: lstm`, `joint_mode: add`, `prejoin_linear: true`, `postjoin_linear: false`,
    `activation: tanh`, no HAT joint — the configuration of every shipped rnnt YAML (examples/asr/rnnt/conf/*.yaml:25-45)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], blank: int = 0, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TransducerGreedyB200 runs on CUDA devices only (no CPU fallback)")
        sd = {k: v for k, v in state_dict.items() if k.startswith("predictor.") or k.startswith("joint.")}
        for k in ("joint.post_ffn.weight", "joint.blank_pred.2.weight", "predictor.rnn.weight_ih_l0_reverse"):
            if k in sd:
                raise ValueError(f"unsupported transducer head (found {k}): only LSTM predictor + add/tanh joint are built")
        layers = 0
        while f"predictor.rnn.weight_ih_l{layers}" in sd:
            layers += 1
        if layers == 0 or "predictor.embed.weight" not in sd or "joint.ffn_out.weight" not in sd:
            raise ValueError("state_dict holds no LSTM predictor / joint (keys predictor.rnn.weight_ih_l0, joint.ffn_out.weight)")
        if sd["predictor.rnn.weight_ih_l0"].shape[0] != 4 * sd["predictor.rnn.weight_hh_l0"].shape[1]:
            raise ValueError("predictor.rnn is not an LSTM (gate rows != 4 * hidden)")
        self.vocab, self.embed = (int(v) for v in sd["predictor.embed.weight"].shape)
        self.hidden = int(sd["predictor.rnn.weight_hh_l0"].shape[1])
        self.layers, self.blank = layers, int(blank)
        self.pred_out = int(sd["predictor.projection.weight"].shape[0])
        self.join_dim, self.enc_dim = (int(v) for v in sd["joint.enc_ffn.weight"].shape)
        self._L = _lib.load()
        cfg = RnntConfig(self.vocab, self.embed, self.hidden, self.layers, self.pred_out, self.enc_dim, self.join_dim, self.blank)
        h = c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._check(self._L.cf_rnnt_create(ctypes.byref(cfg), int(idx), ctypes.byref(h)), "cf_rnnt_create")
        self._h = h
        for key, t in sd.items():
            t = t.detach().to("cpu", torch.float32).contiguous()
            shape = (c_int64 * t.dim())(*t.shape)
            self._check(self._L.cf_rnnt_load_tensor(self._h, key.encode(), c_void_p(t.data_ptr()), t.dim(), shape), "cf_rnnt_load_tensor")
        self._check(self._L.cf_rnnt_finalize_weights(self._h), "cf_rnnt_finalize_weights")
        self.last_iterations = 0

    def set_option(self, name: str, value: int) -> None:
        """Per-handle option of the library (cf_rnnt_set_option): "persistent" 1 / 0 = the search as one persistent
        cooperative kernel / one launch per phase."""
        self._check(self._L.cf_rnnt_set_option(self._h, name.encode(), int(value)), "cf_rnnt_set_option")

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.cf_rnnt_last_error(getattr(self, "_h", None)).decode("utf-8", "replace")
            raise RuntimeError(f"chunkformer_b200 {what} failed ({rc}): {msg}")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.cf_rnnt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def search_flat(self, enc: torch.Tensor, seg_start: Sequence[int], seg_len: Sequence[int], n_steps: int = 64,
                    capacity: int = 0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """enc: (rows, enc_dim) encoder rows on the device; utterance b owns rows [seg_start[b], seg_start[b] + seg_len[b]).
        Returns per utterance (symbols int64, frame of each symbol int32) on the host, in emission order."""
        e = enc.reshape(-1, self.enc_dim).to(self.device, torch.float32).contiguous()
        rows, B = e.shape[0], len(seg_start)
        if B == 0:
            return []
        lens = [max(int(v), 0) for v in seg_len]
        cap = int(capacity) if capacity > 0 else max(16, min(max(lens) * n_steps, 4 * max(lens) + 64))
        while True:
            toks = torch.empty((B, cap), dtype=torch.int64, device=self.device)
            frames = torch.empty((B, cap), dtype=torch.int32, device=self.device)
            counts = torch.empty(B, dtype=torch.int32, device=self.device)
            ws = torch.empty(int(self._L.cf_rnnt_workspace_bytes(self._h, rows, B)), dtype=torch.uint8, device=self.device)
            st = (c_int64 * B)(*[int(v) for v in seg_start])
            ln = (c_int32 * B)(*lens)
            iters = c_int64(0)
            rc = self._L.cf_rnnt_greedy(self._h, c_void_p(e.data_ptr()), rows, st, ln, B, int(n_steps), cap,
                                        c_void_p(toks.data_ptr()), c_void_p(frames.data_ptr()), c_void_p(counts.data_ptr()),
                                        ctypes.byref(iters), c_void_p(ws.data_ptr()), ws.numel(),
                                        c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc == _lib.CF_ERR_WORKSPACE and capacity <= 0 and cap < max(lens) * n_steps:
                cap = min(cap * 4, max(lens) * n_steps)          # rare: more than ~4 symbols per frame on average
                continue
            self._check(rc, "cf_rnnt_greedy")
            break
        self.last_iterations = int(iters.value)
        n = counts.cpu().tolist()
        th, fh = toks.cpu(), frames.cpu()
        return [(th[b, :n[b]].clone(), fh[b, :n[b]].clone()) for b in range(B)]

    def optimized_search(self, encoder_out: torch.Tensor, encoder_out_lens, n_steps: int = 64) -> torch.Tensor:
        """greedy_search.py:6-80: (B, T, E) padded encoder output -> (B, T * n_steps) int64 grid (blank = nothing emitted)."""
        B, T, _ = encoder_out.shape
        lens = [min(int(v), T) for v in (encoder_out_lens.tolist() if torch.is_tensor(encoder_out_lens) else encoder_out_lens)]
        pairs = self.search_flat(encoder_out.reshape(B * T, -1), [b * T for b in range(B)], lens, n_steps)
        grid = torch.full((B, T, n_steps), self.blank, dtype=torch.int64)
        for b, (tok, fr) in enumerate(pairs):
            if tok.numel():
                f = fr.to(torch.int64)
                first = torch.ones_like(f, dtype=torch.bool)
                first[1:] = f[1:] != f[:-1]
                run_start = torch.cummax(torch.where(first, torch.arange(f.numel()), torch.zeros_like(f)), 0).values
                grid[b, f, torch.arange(f.numel()) - run_start] = tok      # k-th symbol of its frame
        return grid.reshape(B, T * n_steps)

    def batch_greedy_search(self, encoder_out: torch.Tensor, encoder_out_lens, n_steps: int = 64) -> List[List[int]]:
        """greedy_search.py:84-99: the emitted symbols of every utterance."""
        B, T, _ = encoder_out.shape
        lens = [min(int(v), T) for v in (encoder_out_lens.tolist() if torch.is_tensor(encoder_out_lens) else encoder_out_lens)]
        return [tok.tolist() for tok, _ in self.search_flat(encoder_out.reshape(B * T, -1), [b * T for b in range(B)], lens, n_steps)]

    def greedy_search(self, encoder_out: torch.Tensor, encoder_out_lens, n_steps: int = 64) -> List[List[int]]:
        """greedy_search.py:147-170 (the unified entry point: `basic_greedy_search` for one utterance, `batch_greedy_search`
        otherwise; both emit the same symbols, so one device path serves them)."""
        return self.batch_greedy_search(encoder_out, encoder_out_lens, n_steps)

    basic_greedy_search = greedy_search
