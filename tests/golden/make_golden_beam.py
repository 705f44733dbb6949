"""Golden outputs of the reference's CTC prefix beam search (modules/search.py:131-249) on random log-probabilities.

As shipped, the reference's function raises TypeError on its first frame: PrefixScore.score() (search.py:88) calls
log_add(self.s, self.ns) while utils/common.py:201 defines log_add(arr) with ONE list argument.  The goldens come from the
unmodified function with that one helper made tolerant of both call forms (log_add(a, b) == log_add([a, b])), which is the
evident intent (WeNet's implementation, which this is taken from, has exactly that helper); nothing else is touched."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

import_reference()
import chunkformer.modules.search as search  # noqa: E402
from chunkformer.utils.common import log_add as _log_add  # noqa: E402

# confirm the crash of the shipped function first (so the repair is not applied to something that works)
try:
    search.ctc_prefix_beam_search(torch.log_softmax(torch.randn(1, 3, 5), -1), torch.tensor([3]), 2)
    crashed = False
except TypeError as e:
    crashed = "log_add" in str(e)
assert crashed, "the reference's ctc_prefix_beam_search no longer raises: regenerate without the repair"
search.log_add = lambda *a: _log_add(list(a[0]) if (len(a) == 1 and isinstance(a[0], (list, tuple))) else list(a))

g = torch.Generator().manual_seed(11)
cases, arrays = [], {}
for k, (B, T, V, beam, peaky) in enumerate([(3, 40, 30, 4, 2.0), (2, 120, 200, 10, 4.0), (4, 25, 12, 3, 1.0), (1, 200, 160, 8, 6.0),
                                            (2, 60, 40, 1, 3.0), (3, 80, 64, 16, 0.5)]):
    logits = torch.randn((B, T, V), generator=g) * peaky
    logits[..., 0] += peaky                     # blank-heavy like a CTC model
    # runs of repeated labels so that the "u == last" branch is exercised
    for b in range(B):
        t = 0
        while t < T:
            run = int(torch.randint(1, 5, (1,), generator=g))
            lab = int(torch.randint(0, V, (1,), generator=g))
            logits[b, t:t + run, lab] += 2.0 * peaky
            t += run + int(torch.randint(0, 3, (1,), generator=g))
    logp = torch.log_softmax(logits, -1)
    lens = torch.tensor([T - 3 * b for b in range(B)])
    res = search.ctc_prefix_beam_search(logp, lens, beam)
    arrays[f"logp_{k}"] = logp.numpy().astype(np.float32)
    cases.append({"lens": lens.tolist(), "beam": beam,
                  "results": [{"tokens": r.tokens, "score": r.score, "times": r.times, "nbest": r.nbest,
                               "nbest_scores": r.nbest_scores, "nbest_times": r.nbest_times} for r in res]})
    print(k, [len(r.tokens) for r in res], [round(r.score, 3) for r in res])
np.savez_compressed(os.path.join(HERE, "beam.npz"), cases=json.dumps(cases), **arrays)
print("wrote beam.npz")
