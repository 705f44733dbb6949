"""Host-side token post-processing (stays Python on the host, as in the reference: chunkformer/utils/model_utils.py).

Same observable behaviour as `remove_duplicates_and_blank` (:23-32), `class2str` (:135-139), `get_output` (:164-171) and
`get_output_with_timestamps` (:174-222): CTC collapse, id -> text with the sentencepiece space marker, and
silence-based segmentation with hh:mm:ss:ms stamps (80 ms per encoder frame)."""
import math
from typing import Dict, List, Sequence

import torch


def ctc_collapse(tokens: Sequence[int], blank_id: int = 0) -> List[int]:
    out: List[int] = []
    prev = None
    for t in tokens:
        t = int(t)
        if t != prev and t != blank_id:
            out.append(t)
        prev = t
    return out


def ids_to_text(ids: Sequence[int], char_dict: Dict[int, str]) -> str:
    return "".join(char_dict[int(i)] for i in ids).replace("▁", " ")


def format_ms(ms: int) -> str:
    h, rem = divmod(int(ms), 3600 * 1000)
    m, rem = divmod(rem, 60 * 1000)
    s, rem = divmod(rem, 1000)
    return f"{h:02}:{m:02}:{s:02}:{rem:03}"


def get_output(hyps, char_dict: Dict[int, str], model_type: str = "asr_model") -> List[str]:
    res = []
    for hyp in hyps:
        ids = [int(v) for v in (hyp.tolist() if torch.is_tensor(hyp) else hyp)]
        if model_type == "asr_model":
            ids = ctc_collapse(ids)
        res.append(ids_to_text(ids, char_dict).strip())
    return res


def get_output_with_timestamps(hyps, char_dict: Dict[int, str], model_type: str, max_silence_duration: float):
    """hyps: iterable of (T', k) token tensors (k = 1 for CTC). A segment closes after `max_silence_duration // 0.08`
    consecutive blank frames; its start is pulled back by up to 2 frames (midpoint to the previous segment's end)."""
    results = []
    max_silence = max_silence_duration // 0.08
    for tokens in hyps:
        tokens = tokens.cpu()
        rows = tokens.reshape(tokens.shape[0], -1).tolist()
        start = end = prev_end = -1
        silence = 0
        pending: List[int] = []
        segs = []
        t = -1
        for t, row in enumerate(rows):
            nonblank = [v for v in row if v != 0]
            if not nonblank:
                silence += 1
            else:
                if start == -1 and end == -1:
                    start = max(math.ceil((t + prev_end) / 2), t - 2) if prev_end != -1 else max(t - 2, 0)
                silence = 0
                pending.extend(nonblank)
            if silence == max_silence and start != -1:
                end = t
                prev_end = end
                segs.append({"decode": get_output([pending], char_dict, model_type)[0],
                             "start": format_ms(start * 80), "end": format_ms(end * 80)})
                pending, start, end, silence = [], -1, -1, 0
        if start != -1 and end == -1 and pending:
            segs.append({"decode": get_output([pending], char_dict, model_type)[0],
                         "start": format_ms(start * 80), "end": format_ms(t * 80)})
        results.append(segs)
    return results


def get_output_with_timestamps_compact(frames: Sequence[int], tokens: Sequence[int], n_frames: int, char_dict: Dict[int, str],
                                       model_type: str, max_silence_duration: float):
    """`get_output_with_timestamps` for ONE hypothesis given only its non-blank symbols (non-decreasing frame indices and
    token ids, as produced on the device by `cf_ctc_compact` mode 1 or by the transducer search, which may emit several
    symbols on one frame) and the number of encoder frames: the same
    segments, texts and stamps as the frame-by-frame loop (utils/model_utils.py:174-222), in O(#non-blank) host work.

    A segment that saw its last non-blank frame at `last` closes at frame `last + max_silence` when no other non-blank
    frame arrives up to and including that frame and the frame exists; otherwise it runs to the last frame."""
    max_silence = max_silence_duration // 0.08
    frames = [int(f) for f in (frames.tolist() if torch.is_tensor(frames) else frames)]
    tokens = [int(t) for t in (tokens.tolist() if torch.is_tensor(tokens) else tokens)]
    segs = []
    prev_end = -1
    i, n = 0, len(frames)
    while i < n:
        t0 = frames[i]
        start = max(math.ceil((t0 + prev_end) / 2), t0 - 2) if prev_end != -1 else max(t0 - 2, 0)
        pending = [tokens[i]]
        last = t0
        i += 1
        while i < n and frames[i] - last <= max_silence:
            pending.append(tokens[i])
            last = frames[i]
            i += 1
        end = last + max_silence
        if max_silence >= 0 and end <= n_frames - 1:
            end = int(end)
            prev_end = end
            segs.append({"decode": get_output([pending], char_dict, model_type)[0],
                         "start": format_ms(start * 80), "end": format_ms(end * 80)})
        else:
            # the silence counter never reaches the threshold before the stream ends: every remaining frame joins this segment
            while i < n:
                pending.append(tokens[i])
                i += 1
            segs.append({"decode": get_output([pending], char_dict, model_type)[0],
                         "start": format_ms(start * 80), "end": format_ms((n_frames - 1) * 80)})
    return segs


# ------------------------------------------------------------------------------------------------------------------
# CTC prefix beam search
# ------------------------------------------------------------------------------------------------------------------
class DecodeResult:
    """Fields of the reference's DecodeResult that ctc_prefix_beam_search fills (modules/search.py:35-64, :231-246)."""

    def __init__(self, tokens, score=0.0, times=None, nbest=None, nbest_scores=None, nbest_times=None):
        self.tokens, self.score, self.times = tokens, score, times if times is not None else []
        self.nbest = nbest if nbest is not None else [tokens]
        self.nbest_scores = nbest_scores if nbest_scores is not None else [score]
        self.nbest_times = nbest_times if nbest_times is not None else [self.times]


_NEG = -float("inf")


def _lse2(a: float, b: float) -> float:
    """log(exp(a) + exp(b)) as utils/common.py:201-209 computes it for a two-element list."""
    if a == _NEG and b == _NEG:
        return _NEG
    m = a if a > b else b
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def prefix_beam_search_topk(top_logp, top_idx, num_t: int, beam_size: int, blank_id: int = 0):
    """CTC prefix beam search of one utterance on per-frame top-k lists (best first), i.e. on what the reference keeps after its
    first prune `logp.topk(beam_size)` (modules/search.py:167-169); the device computes the lists, only T x beam values cross to
    the host.  top_logp / top_idx: (T, >= beam_size) array-likes.  Returns (nbest tokens, nbest scores, nbest times).

    Drop-in for the per-utterance body of the reference's ctc_prefix_beam_search (search.py:144-229) without a context graph: the
    same hypotheses in the same order (candidate order = descending probability, prefix order = the previous beam, ties of the
    second prune resolved by insertion order as Python's stable sort does), so it returns what the reference returns once
    PrefixScore.score() can call log_add (see tests/golden/make_golden_beam.py).  A hypothesis is a list
    [s, ns, v_s, v_ns, cur_token_prob, times_s, times_ns] (blank-ending / non-blank-ending log-probability, Viterbi scores and
    their frame lists)."""
    cur = [((), [0.0, _NEG, 0.0, 0.0, _NEG, [], []])]
    for t in range(int(num_t)):
        lp, ix = top_logp[t], top_idx[t]
        nxt = {}
        for j in range(beam_size):
            u, prob = int(ix[j]), float(lp[j])
            for prefix, h in cur:
                s, ns, v_s, v_ns = h[0], h[1], h[2], h[3]
                vit_blank = v_s > v_ns                       # PrefixScore.viterbi_score() / times()
                vit, vit_times = (v_s, h[5]) if vit_blank else (v_ns, h[6])
                if u == blank_id:
                    n = nxt.get(prefix)
                    if n is None:
                        n = nxt[prefix] = [_NEG, _NEG, _NEG, _NEG, _NEG, [], []]
                    n[0] = _lse2(n[0], _lse2(s, ns) + prob)
                    n[2] = vit + prob
                    n[5] = list(vit_times)
                elif prefix and u == prefix[-1]:
                    n = nxt.get(prefix)                      # *uu -> *u
                    if n is None:
                        n = nxt[prefix] = [_NEG, _NEG, _NEG, _NEG, _NEG, [], []]
                    n[1] = _lse2(n[1], ns + prob)
                    if n[3] < v_ns + prob:
                        n[3] = v_ns + prob
                        if n[4] < prob:
                            n[4] = prob
                            n[6] = list(h[6])
                            n[6][-1] = t
                    ext = prefix + (u,)                      # *u-u -> *uu
                    n = nxt.get(ext)
                    if n is None:
                        n = nxt[ext] = [_NEG, _NEG, _NEG, _NEG, _NEG, [], []]
                    n[1] = _lse2(n[1], s + prob)
                    if n[3] < v_s + prob:
                        n[3] = v_s + prob
                        n[4] = prob
                        n[6] = h[5] + [t]
                else:
                    ext = prefix + (u,)
                    n = nxt.get(ext)
                    if n is None:
                        n = nxt[ext] = [_NEG, _NEG, _NEG, _NEG, _NEG, [], []]
                    n[1] = _lse2(n[1], _lse2(s, ns) + prob)
                    if n[3] < vit + prob:
                        n[3] = vit + prob
                        n[4] = prob
                        n[6] = vit_times + [t]
        cur = sorted(nxt.items(), key=lambda kv: _lse2(kv[1][0], kv[1][1]), reverse=True)[:beam_size]
    nbest = [list(p) for p, _ in cur]
    scores = [_lse2(h[0], h[1]) for _, h in cur]
    times = [(h[5] if h[2] > h[3] else h[6]) for _, h in cur]
    return nbest, scores, times


def ctc_prefix_beam_search(ctc_probs: torch.Tensor, ctc_lens: torch.Tensor, beam_size: int, context_graph=None,
                           blank_id: int = 0) -> List[DecodeResult]:
    """Drop-in for chunkformer.modules.search.ctc_prefix_beam_search (search.py:131-249): ctc_probs (B, T, V) log-probabilities on
    any device, ctc_lens (B,).  The first prune (top-k per frame) runs where the tensor lives, the prefix search on the host."""
    if context_graph is not None:
        raise NotImplementedError("context biasing is outside the hot path (DESIGN.md section 4)")
    if ctc_probs.dim() != 3:
        raise ValueError("ctc_probs must be (B, T, V)")
    k = min(int(beam_size), ctc_probs.shape[2])
    top_logp, top_idx = ctc_probs.float().topk(k, dim=2)
    top_logp, top_idx = top_logp.cpu().numpy(), top_idx.cpu().numpy()
    out = []
    for b in range(ctc_probs.shape[0]):
        nbest, scores, times = prefix_beam_search_topk(top_logp[b], top_idx[b], int(ctc_lens[b]), k, blank_id)
        out.append(DecodeResult(nbest[0], scores[0], times[0], nbest, scores, times))
    return out
