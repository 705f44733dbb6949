"""Golden outputs of the reference's host-side token post-processing (utils/model_utils.py) on random token streams."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

import_reference()
from chunkformer.utils.model_utils import get_output, get_output_with_timestamps  # noqa: E402

rng = np.random.RandomState(5)
char_dict = {0: "<blank>", 1: "<unk>"}
for i in range(2, 40):
    char_dict[i] = ("▁" if i % 3 == 0 else "") + chr(ord("a") + i % 26) + ("x" if i % 7 == 0 else "")
cases = []
for k in range(40):
    T = int(rng.randint(1, 400))
    toks = []
    while len(toks) < T:
        if rng.rand() < 0.5:
            toks += [0] * int(rng.randint(1, 15))
        else:
            toks += [int(rng.randint(1, 40))] * int(rng.randint(1, 4))
    toks = toks[:T]
    msd = float(rng.choice([0.5, 0.2, 1.0, 0.08]))
    t = torch.tensor(toks, dtype=torch.long)
    cases.append({"tokens": toks, "max_silence": msd,
                  "text": get_output([t], char_dict, "asr_model")[0],
                  "segments": get_output_with_timestamps([t.reshape(-1, 1)], char_dict, "asr_model", msd)[0]})
json.dump({"char_dict": {str(k): v for k, v in char_dict.items()}, "cases": cases},
          open(os.path.join(HERE, "postprocess.json"), "w"), ensure_ascii=False)
print("postprocess.json", len(cases))
