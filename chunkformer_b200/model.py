"""Drop-in facade: the reference's `ChunkFormerModel` method surface (chunkformer/chunkformer_model.py:57-640) on top of
the B200 encoder.  Same method names, argument names / defaults, return types and errors; the checkpoint directory
layout (config.yaml, global_cmvn, pytorch_model.{bin,pt,ckpt}, vocab.txt, label_mapping.json) is read as is.

`model: transducer` checkpoints decode through the CUDA greedy search (transducer.py).  Out of scope here:
Hugging Face Hub download (no network) and audio-file decoding through pydub; `endless_decode` / `batch_decode` /
`classify_audio` accept either a path to a wav file (torchaudio) or an already extracted (T, 80) fbank tensor.
`batch_decode(..., devices=[...])` replaces the reference's arrival-order admission by duration-balanced scheduling over
several GPUs of one box."""
import json
import math
import os
from typing import Dict, List, Optional, Union

import torch

from .encoder import ChunkFormerEncoderB200
from .geometry import EncoderGeometry
from .transducer import TransducerGreedyB200
from .postprocess import get_output, get_output_with_timestamps, get_output_with_timestamps_compact, ids_to_text


def load_cmvn_json(path: str):
    """utils/cmvn.py:23-46: mean_stat / var_stat / frame_num -> (mean, 1/std)."""
    with open(path) as f:
        st = json.load(f)
    n = st["frame_num"]
    mean = [m / n for m in st["mean_stat"]]
    istd = []
    for m, v in zip(mean, st["var_stat"]):
        var = max(v / n - m * m, 1.0e-20)
        istd.append(1.0 / math.sqrt(var))
    return torch.tensor(mean, dtype=torch.float32), torch.tensor(istd, dtype=torch.float32)


def read_symbol_table(path: str) -> Dict[str, int]:
    table = {}
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            arr = line.strip().split()
            assert len(arr) == 2
            table[arr[0]] = int(arr[1])
    return table


class _CTCHead:
    """Mirror of chunkformer/modules/ctc.py:73-91 on the fused GEMM + argmax / log-softmax kernels."""

    def __init__(self, encoder: ChunkFormerEncoderB200):
        self._enc = encoder

    def log_softmax(self, hs_pad: torch.Tensor) -> torch.Tensor:
        return self._enc.ctc_greedy(hs_pad, want_logp=True)[1]

    def argmax(self, hs_pad: torch.Tensor) -> torch.Tensor:
        return self._enc.ctc_greedy(hs_pad)


class ChunkFormerModel:
    def __init__(self, config: dict, state_dict: Dict[str, torch.Tensor], device: Union[str, torch.device] = "cuda:0"):
        self.config = dict(config)
        self.device = torch.device(device)
        self.model_type = self.config.get("model", "asr_model")
        enc_conf = self.config.get("encoder_conf", {})
        self.is_classification = self.model_type == "classification" or "tasks" in self.config.get("model_conf", {})
        vocab = 0 if self.is_classification else int(self.config.get("output_dim", 0))
        has_cmvn = "encoder.global_cmvn.mean" in state_dict
        if vocab and "ctc.ctc_lo.weight" not in state_dict:
            vocab = 0
        geo = EncoderGeometry.from_encoder_conf(enc_conf, self.config.get("input_dim", 80), vocab, has_cmvn)
        self.geometry = geo
        self.encoder = ChunkFormerEncoderB200(geo, state_dict, self.device)
        self._state_dict = state_dict               # kept for weight replicas on further GPUs (batch_decode(devices=...))
        self._replicas = {}
        self.ctc = _CTCHead(self.encoder) if vocab else None
        # chunkformer-rnnt-*: LSTM predictor + joint behind the encoder (utils/init_model.py:118-133)
        self.transducer = None
        if self.model_type == "transducer" and "predictor.embed.weight" in state_dict:
            blank = int(self.config.get("ctc_conf", {}).get("ctc_blank_id", 0))
            self.transducer = TransducerGreedyB200(state_dict, blank=blank, device=self.device)
        self.char_dict: Optional[Dict[int, str]] = None
        self.label_mapping = None
        self.tasks = None
        self._heads = {}
        if self.is_classification:
            self.tasks = dict(self.config.get("model_conf", {}).get("tasks", {}))
            for name in self.tasks:
                w = state_dict[f"classification_heads.{name}.linear.weight"].to(self.device, torch.float32)
                b = state_dict[f"classification_heads.{name}.linear.bias"].to(self.device, torch.float32)
                self._heads[name] = (w, b)

    # ---------------------------------------------------------------------------------------------- loading
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, config: Optional[dict] = None,
                        cache_dir: Optional[str] = None, force_download: bool = False, device="cuda:0", **kwargs):
        """chunkformer_model.py:106-207, local directories only (there is no network here)."""
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise ValueError(f"No config found in {path}")
        if config is None:
            cfg_path = os.path.join(path, "config.yaml")
            if not os.path.exists(cfg_path):
                raise ValueError(f"No config found in {path}")
            import yaml
            with open(cfg_path, "r") as f:
                config = yaml.load(f, Loader=yaml.FullLoader)
        ckpt = None
        candidates = ["pytorch_model.bin", "pytorch_model.pt", "pytorch_model.ckpt"]
        for cand in candidates:
            if os.path.exists(os.path.join(path, cand)):
                ckpt = os.path.join(path, cand)
                break
        if ckpt is None:
            raise ValueError(f"No checkpoint found in {path}. Expected one of: {candidates}")
        sd = torch.load(ckpt, map_location="cpu", mmap=True, weights_only=True)   # utils/checkpoint.py:26-41
        cmvn_path = os.path.join(path, "global_cmvn")
        if os.path.exists(cmvn_path) and "encoder.global_cmvn.mean" not in sd:
            mean, istd = load_cmvn_json(cmvn_path)
            sd["encoder.global_cmvn.mean"], sd["encoder.global_cmvn.istd"] = mean, istd
        model = cls(config, sd, device)
        vocab_path = os.path.join(path, "vocab.txt")
        if os.path.exists(vocab_path):
            model.char_dict = {v: k for k, v in read_symbol_table(vocab_path).items()}
        lm = os.path.join(path, "label_mapping.json")
        if os.path.exists(lm):
            with open(lm) as f:
                model.label_mapping = json.load(f)
        return model

    def forward(self, **kwargs):
        raise NotImplementedError(
            "Forward method is not implemented. If you want to use ChunkFormer for feature extraction from pretrained "
            "model please use 'chunkformer.encode()' instead. Or if you want transcription please use "
            "'endless_decode' or 'batch_decode'.")

    def get_encoder(self):
        return self.encoder

    def get_ctc(self):
        return self.ctc

    def get_tasks(self):
        return self.tasks if self.is_classification else None

    # ---------------------------------------------------------------------------------------------- features
    def extract_features(self, waveform: torch.Tensor, sample_rate: Optional[int] = None):
        """16-bit-range mono waveform (1, n) or (n,) at the model's sample rate -> fbank (T, num_mel_bins) on the device, by
        the CUDA kernel behind cf_fbank (Kaldi-compatible; same values as the reference's kaldi.fbank call,
        chunkformer_model.py:307-315, within 2e-3)."""
        fbank_conf = self.config.get("fbank_conf", {})
        sr = sample_rate or self.config.get("resample_conf", {}).get("resample_rate", 16000)
        x = self.encoder.fbank(waveform, num_mel_bins=fbank_conf.get("num_mel_bins", 80),
                               frame_length=fbank_conf.get("frame_length", 25), frame_shift=fbank_conf.get("frame_shift", 10),
                               sample_frequency=sr)
        return x, int(x.shape[0])

    def _load_audio_and_extract_features(self, audio):
        """chunkformer_model.py:276-318. Accepts an fbank tensor (T, feat) or a wav path; the file is decoded and resampled by
        torchaudio on the host (the reference uses pydub), the fbank itself runs on the GPU (cf_fbank)."""
        if torch.is_tensor(audio):
            return audio, int(audio.shape[0])
        import torchaudio
        sr = self.config.get("resample_conf", {}).get("resample_rate", 16000)
        wav, in_sr = torchaudio.load(audio)
        wav = wav.mean(0, keepdim=True)
        if in_sr != sr:
            wav = torchaudio.functional.resample(wav, in_sr, sr)
        wav = torch.round(wav * (1 << 15)).clamp_(-32768, 32767)     # 16-bit samples as floats, like set_sample_width(2)
        return self.extract_features(wav, sr)

    # ---------------------------------------------------------------------------------------------- encode
    @torch.no_grad()
    def encode(self, xs: torch.Tensor, xs_lens: torch.Tensor, chunk_size: Optional[int] = None,
               left_context_size: Optional[int] = None, right_context_size: Optional[int] = None, **kwargs):
        """chunkformer_model.py:256-274 -> forward_encoder; returns (feats (B, T', d), lens (B,))."""
        out, masks = self.encoder.forward_encoder(xs, xs_lens, chunk_size or 0, left_context_size or 0, right_context_size or 0)
        return out, masks.squeeze(1).sum(-1)

    # ---------------------------------------------------------------------------------------------- long-form
    @torch.no_grad()
    def endless_decode(self, audio_path, chunk_size: Optional[int] = 64, left_context_size: Optional[int] = 128,
                       right_context_size: Optional[int] = 128, total_batch_duration: int = 1800,
                       return_timestamps: bool = True, max_silence_duration: float = 0.5):
        """chunkformer_model.py:320-459: sequential segments with K/V + conv caches carried between them."""
        if self.model_type == "asr_model" and self.ctc is None or self.model_type != "asr_model" and self.transducer is None:
            raise ValueError("endless_decode needs a CTC head (asr_model) or an LSTM predictor + joint (transducer)")
        c = chunk_size if chunk_size is not None else 64
        l = left_context_size if left_context_size is not None else 128
        r = right_context_size if right_context_size is not None else 128
        geo = self.geometry
        sub, lo = 8, geo.kernel // 2
        max_len = int(total_batch_duration // 0.01) // 2
        multiply_n = max_len // c // sub
        trunc = c * multiply_n
        rr = max(r, lo)
        rel_right = (rr + max(c, rr) * (geo.layers - 1)) * sub
        xs, xs_len = self._load_audio_and_extract_features(audio_path)
        offset = torch.zeros(1, dtype=torch.int)
        att_cache = torch.zeros((geo.layers, l, geo.heads, 2 * geo.d_k), device=self.device)
        cnn_cache = torch.zeros((geo.layers, geo.d_model, lo), device=self.device)
        outs = []
        for idx, _ in enumerate(range(0, xs_len, trunc * sub)):
            start = trunc * sub * idx
            end = min(trunc * sub * (idx + 1) + 7, xs_len)
            x = xs[start:end + rel_right]
            out, enc_len, _, att_cache, cnn_cache, offset = self.encoder.forward_parallel_chunk(
                xs=[x], xs_origin_lens=torch.tensor([x.shape[0]], dtype=torch.int), chunk_size=c, left_context_size=l,
                right_context_size=r, att_cache=att_cache, cnn_cache=cnn_cache, truncated_context_size=trunc, offset=offset)
            out = out.reshape(1, -1, out.shape[-1])[:, :max(int(enc_len[0]), 0)]
            last = not (trunc * sub * idx + rel_right < xs_len)
            if not last:
                out = out[:, :trunc]                      # drop the rows computed only as right context
            offset = offset - enc_len + out.shape[1]
            outs.append(out)
            if last:
                break
        enc = torch.cat(outs, dim=1)
        if self.model_type != "asr_model":
            # chunkformer_model.py:440-447: optimized_search on the whole recording, (1, T', n_steps) symbol grid
            n_frames = enc.shape[1]
            if self.char_dict is None:
                return self.transducer.optimized_search(enc, [n_frames]).reshape(1, n_frames, -1)
            (tok, frames), = self.transducer.search_flat(enc[0], [0], [n_frames])
            res = get_output_with_timestamps_compact(frames, tok, n_frames, self.char_dict, self.model_type,
                                                     max_silence_duration) if n_frames > 0 else []
            if not return_timestamps:
                res = " ".join(item["decode"] for item in res).strip()
            return res
        tokens = self.ctc.argmax(enc).reshape(1, -1, 1)
        if self.char_dict is not None:
            # only the non-blank frames cross PCIe (cf_ctc_compact mode 1); segmentation runs on that short list and gives
            # exactly what get_output_with_timestamps gives on the full frame sequence (tests/test_postprocess.py)
            n_frames = tokens.shape[1]
            (tok, frames), = self.encoder.ctc_compact(tokens, [0], [n_frames], mode=1)
            res = get_output_with_timestamps_compact(frames, tok, n_frames, self.char_dict, self.model_type,
                                                     max_silence_duration) if n_frames > 0 else []
            if not return_timestamps:
                res = " ".join(item["decode"] for item in res).strip()
            return res
        return tokens

    # ---------------------------------------------------------------------------------------------- batch
    def _decode_group_enqueue(self, encoder, transducer, xs, lens, c, l, r):
        """Encoder + greedy CTC ids (or the encoder rows for a transducer) of one masked batch, left on the device."""
        out, enc_lens, n_chunks, _, _, _ = encoder.forward_parallel_chunk(
            xs=xs, xs_origin_lens=torch.tensor(lens, dtype=torch.int), chunk_size=c, left_context_size=l,
            right_context_size=r, offset=torch.zeros(len(xs), dtype=torch.int))
        starts, row = [], 0
        for n in n_chunks:
            starts.append(row * c)
            row += int(n)
        valid = [max(int(m), 0) for m in enc_lens]
        tokens = encoder.ctc_greedy(out) if self.model_type == "asr_model" else None
        return dict(encoder=encoder, transducer=transducer, out=out, tokens=tokens, n_chunks=n_chunks, starts=starts, valid=valid)

    def _decode_group_collect(self, g):
        """Host side of one group: device-side compaction / transducer search, ids -> text."""
        if self.model_type != "asr_model":
            # chunkformer_model.py:532-543: batch_greedy_search; the flat chunk rows are searched in place
            out = g["out"]
            pairs = g["transducer"].search_flat(out.reshape(-1, out.shape[-1]), g["starts"], g["valid"])
            hyps = [tok.tolist() for tok, _ in pairs]
            return get_output(hyps, self.char_dict, self.model_type) if self.char_dict is not None else hyps
        if self.char_dict is not None:
            # CTC collapse on the device (cf_ctc_compact mode 0): one short id list per utterance comes back
            pairs = g["encoder"].ctc_compact(g["tokens"], g["starts"], g["valid"], mode=0)
            return [ids_to_text(tok.tolist(), self.char_dict).strip() for tok, _ in pairs]
        hyps = g["tokens"].split(g["n_chunks"], dim=0)
        return [h.flatten()[:m] for h, m in zip(hyps, g["valid"])]

    def _replica(self, device):
        """Encoder (+ transducer head) replica on another GPU of this box: weights are replicated, work is sharded
        (SURVEY.md 8e)."""
        device = torch.device(device)
        if device == self.device:
            return self.encoder, self.transducer
        if device not in self._replicas:
            enc = ChunkFormerEncoderB200(self.geometry, self._state_dict, device)
            tr = None
            if self.transducer is not None:
                blank = int(self.config.get("ctc_conf", {}).get("ctc_blank_id", 0))
                tr = TransducerGreedyB200(self._state_dict, blank=blank, device=device)
            self._replicas[device] = (enc, tr)
        return self._replicas[device]

    @torch.no_grad()
    def batch_decode(self, audio_paths: List, chunk_size: Optional[int] = 64, left_context_size: Optional[int] = 128,
                     right_context_size: Optional[int] = 128, total_batch_duration: int = 1800, devices: Optional[List] = None):
        """chunkformer_model.py:461-552: greedy arrival-order admission, one masked batch per group.

        devices=[...] (an extension): the admission is replaced by duration-balanced scheduling over several
        GPUs of this box.  `total_batch_duration` stays the per-GPU budget of one masked batch; utterances are dealt to the
        devices longest first by chunk count (shard.partition_by_chunks), every device runs its share as one masked batch,
        all devices of a round are enqueued before any result is read back, and the results come back in input order.
        Utterances are independent in a masked batch, so the texts equal the single-device ones."""
        if self.model_type == "asr_model" and self.ctc is None or self.model_type != "asr_model" and self.transducer is None:
            raise ValueError("batch_decode needs a CTC head (asr_model) or an LSTM predictor + joint (transducer)")
        c = chunk_size if chunk_size is not None else 64
        l = left_context_size if left_context_size is not None else 128
        r = right_context_size if right_context_size is not None else 128
        budget0 = int(total_batch_duration // 0.01) // 2
        if devices is not None and len(devices) > 0:
            return self._batch_decode_balanced(audio_paths, c, l, r, budget0, [torch.device(d) for d in devices])
        def groups():                      # the reference's arrival-order admission (chunkformer_model.py:496-504)
            xs, lens, budget = [], [], budget0
            for i, a in enumerate(audio_paths):
                x, n = self._load_audio_and_extract_features(a)
                xs.append(x)
                lens.append(n)
                budget -= n
                if budget <= 0 or i == len(audio_paths) - 1:
                    yield xs, lens
                    xs, lens, budget = [], [], budget0
        decodes = []
        for xs, lens in groups():
            g = self._decode_group_enqueue(self.encoder, self.transducer, xs, lens, c, l, r)
            decodes.extend(self._decode_group_collect(g))
        return decodes

    def _batch_decode_balanced(self, audio_paths, c, l, r, budget0, devices):
        from .shard import chunks_of, partition_by_chunks
        feats = [self._load_audio_and_extract_features(a) for a in audio_paths]
        lens = [n for _, n in feats]
        order = sorted(range(len(lens)), key=lambda i: -chunks_of(lens[i], c))
        # rounds: longest first until the round holds one budget per device (a single utterance may exceed it, as in the
        # reference, where a group closes only after the budget is spent)
        rounds, cur, frames = [], [], 0
        for i in order:
            cur.append(i)
            frames += lens[i]
            if frames >= budget0 * len(devices):
                rounds.append(cur)
                cur, frames = [], 0
        if cur:
            rounds.append(cur)
        results = [None] * len(lens)
        for rd in rounds:
            bins = partition_by_chunks([lens[i] for i in rd], c, len(devices))
            groups = []
            for dev, b in zip(devices, bins):                   # enqueue every device before reading anything back
                if not b:
                    continue
                idx = [rd[k] for k in b]
                enc, tr = self._replica(dev)
                groups.append((idx, self._decode_group_enqueue(enc, tr, [feats[i][0] for i in idx], [lens[i] for i in idx], c, l, r)))
            for idx, g in groups:
                for i, hyp in zip(idx, self._decode_group_collect(g)):
                    results[i] = hyp
        return results

    # ---------------------------------------------------------------------------------------------- classification
    @torch.no_grad()
    def classify_audio(self, audio_path, chunk_size: Optional[int] = -1, left_context_size: Optional[int] = -1,
                       right_context_size: Optional[int] = -1):
        """chunkformer_model.py:554-640 + classification_model.py:174-281 (masked mean pool, per-task Linear)."""
        if not self.is_classification:
            raise ValueError("This model is not a classification model. Use ASR decoding methods instead.")
        xs, n = self._load_audio_and_extract_features(audio_path)
        c, l, r = chunk_size, left_context_size, right_context_size
        if c is None or l is None or r is None or c < 0 or l < 0 or r < 0:
            c = l = r = 0                                       # encoder.py:490-493: full attention
        out, masks = self.encoder.forward_encoder(xs.unsqueeze(0), torch.tensor([n]), c, l, r)
        m = masks.transpose(1, 2).float()
        pooled = (out * m).sum(1) / (m.sum(1) + 1e-10)
        result = {}
        for name, (w, b) in self._heads.items():
            logits = pooled @ w.T + b
            pred = int(torch.argmax(logits, -1)[0])
            prob = float(torch.softmax(logits, -1)[0, pred])
            label = str(pred)
            if self.label_mapping and name in self.label_mapping:
                label = self.label_mapping[name].get(str(pred), str(pred))
            result[name] = {"label": label, "label_id": pred, "prob": prob}
        return result
