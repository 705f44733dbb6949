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
