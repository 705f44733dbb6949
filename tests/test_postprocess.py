"""Host-side token post-processing against golden outputs of the reference's utils/model_utils.py."""
import json
import os

import torch

from chunkformer_b200.postprocess import get_output, get_output_with_timestamps
from oracle import chunkformer_oracle as O


def test_postprocess_matches_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "postprocess.json")))
    cd = {int(k): v for k, v in g["char_dict"].items()}
    for case in g["cases"]:
        t = torch.tensor(case["tokens"], dtype=torch.long)
        assert get_output([t], cd, "asr_model")[0] == case["text"]
        assert get_output_with_timestamps([t.reshape(-1, 1)], cd, "asr_model", case["max_silence"])[0] == case["segments"]
        assert O.ctc_collapse(case["tokens"]) == [i for i in __import__("chunkformer_b200.postprocess", fromlist=["x"]).ctc_collapse(case["tokens"])]


def test_segment_and_batch_arithmetic():
    # endless_decode segment geometry at the reference defaults (chunkformer_model.py:359-371)
    trunc, rel_right, segs = O.endless_segments(5759998, 64, 128, 17, 14400)
    assert trunc == 1406 * 64 and rel_right == 17408
    assert len(segs) == 9 and segs[0][0] == 0 and segs[-1][2]
    assert O.batch_groups([100, 200, 90000, 50, 800000, 10], 1800) == [[0, 1, 2], [3, 4], [5]]


def _nonblank(tokens):
    return [i for i, t in enumerate(tokens) if t != 0], [t for t in tokens if t != 0]


def test_compact_timestamps_match_reference_golden(golden_dir):
    """The O(#non-blank) segmentation fed by the device compaction gives the reference's segments (golden cases)."""
    from chunkformer_b200.postprocess import get_output_with_timestamps_compact
    g = json.load(open(os.path.join(golden_dir, "postprocess.json")))
    cd = {int(k): v for k, v in g["char_dict"].items()}
    for case in g["cases"]:
        fr, tk = _nonblank(case["tokens"])
        assert get_output_with_timestamps_compact(fr, tk, len(case["tokens"]), cd, "asr_model", case["max_silence"]) == case["segments"]


def test_compact_timestamps_match_frame_loop_random():
    from chunkformer_b200.postprocess import get_output_with_timestamps_compact
    gen = torch.Generator().manual_seed(7)
    cd = {i: f"▁w{i}" if i % 3 == 0 else f"s{i}" for i in range(12)}
    for case in range(300):
        n = int(torch.randint(1, 120, (1,), generator=gen))
        density = float(torch.rand(1, generator=gen))
        toks = torch.randint(1, 12, (n,), generator=gen)
        toks[torch.rand(n, generator=gen) > density] = 0
        for ms in (0.0, 0.05, 0.08, 0.17, 0.5, 1.0, -0.5):
            want = get_output_with_timestamps([toks.reshape(-1, 1)], cd, "asr_model", ms)[0]
            fr, tk = _nonblank(toks.tolist())
            assert get_output_with_timestamps_compact(fr, tk, n, cd, "asr_model", ms) == want, (case, ms, toks.tolist())


def test_compact_timestamps_multi_symbol_frames():
    """Transducer grids (T', n_steps): several symbols on one frame, no CTC collapse (model type != asr_model)."""
    from chunkformer_b200.postprocess import get_output_with_timestamps_compact
    gen = torch.Generator().manual_seed(11)
    cd = {i: f"▁w{i}" if i % 3 == 0 else f"s{i}" for i in range(12)}
    for case in range(200):
        n, k = int(torch.randint(1, 80, (1,), generator=gen)), 3
        grid = torch.zeros((n, k), dtype=torch.long)
        for t in range(n):
            if float(torch.rand(1, generator=gen)) < 0.35:
                m = int(torch.randint(1, k + 1, (1,), generator=gen))
                grid[t, :m] = torch.randint(1, 12, (m,), generator=gen)
        fr = [t for t in range(n) for j in range(k) if int(grid[t, j]) != 0]
        tk = [int(grid[t, j]) for t in range(n) for j in range(k) if int(grid[t, j]) != 0]
        for ms in (0.0, 0.08, 0.2, 0.5):
            want = get_output_with_timestamps([grid], cd, "transducer", ms)[0]
            assert get_output_with_timestamps_compact(fr, tk, n, cd, "transducer", ms) == want, (case, ms)
