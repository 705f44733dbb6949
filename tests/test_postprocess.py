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
