"""CPU oracle (TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else) of the reference's CTC prefix beam search,
chunkformer/modules/search.py:131-249 with PrefixScore (:67-111), restated frame by frame in plain Python.

Pinned against tests/golden/beam.npz, which tests/golden/make_golden_beam.py produced with the UNMODIFIED reference function after
making its log_add helper accept the two-argument call of PrefixScore.score() (search.py:88 vs utils/common.py:201: as shipped the
function raises TypeError on its first frame).  No context graph (out of scope)."""
import math
from collections import defaultdict
from typing import List, Sequence

NEG = -float("inf")


def log_add(*a: float) -> float:                      # utils/common.py:201-209
    if all(x == NEG for x in a):
        return NEG
    m = max(a)
    return m + math.log(sum(math.exp(x - m) for x in a))


class _Score:                                         # search.py:67-111 without the context fields
    def __init__(self, s=NEG, ns=NEG, v_s=NEG, v_ns=NEG):
        self.s, self.ns, self.v_s, self.v_ns = s, ns, v_s, v_ns
        self.cur_token_prob = NEG
        self.times_s: List[int] = []
        self.times_ns: List[int] = []

    def score(self):
        return log_add(self.s, self.ns)

    def viterbi_score(self):
        return self.v_s if self.v_s > self.v_ns else self.v_ns

    def times(self):
        return self.times_s if self.v_s > self.v_ns else self.times_ns


def ctc_prefix_beam_search(logp: Sequence[Sequence[float]], num_t: int, beam_size: int, blank_id: int = 0):
    """One utterance.  logp[t][v]: log-probabilities as Python floats (the reference reads them with .item()).
    Returns (nbest token lists, nbest scores, nbest times), best first (search.py:144-246)."""
    cur = [(tuple(), _Score(s=0.0, ns=NEG, v_s=0.0, v_ns=0.0))]
    for t in range(num_t):
        row = logp[t]
        order = sorted(range(len(row)), key=lambda v: (-row[v], v))[:beam_size]      # torch.topk: descending, ties by index
        nxt = defaultdict(_Score)
        for u in order:
            prob = row[u]
            for prefix, ps in cur:
                last = prefix[-1] if len(prefix) > 0 else None
                if u == blank_id:
                    n = nxt[prefix]
                    n.s = log_add(n.s, ps.score() + prob)
                    n.v_s = ps.viterbi_score() + prob
                    n.times_s = ps.times().copy()
                elif u == last:
                    n1 = nxt[prefix]
                    n1.ns = log_add(n1.ns, ps.ns + prob)
                    if n1.v_ns < ps.v_ns + prob:
                        n1.v_ns = ps.v_ns + prob
                        if n1.cur_token_prob < prob:
                            n1.cur_token_prob = prob
                            n1.times_ns = ps.times_ns.copy()
                            n1.times_ns[-1] = t
                    n2 = nxt[prefix + (u,)]
                    n2.ns = log_add(n2.ns, ps.s + prob)
                    if n2.v_ns < ps.v_s + prob:
                        n2.v_ns = ps.v_s + prob
                        n2.cur_token_prob = prob
                        n2.times_ns = ps.times_s.copy()
                        n2.times_ns.append(t)
                else:
                    n = nxt[prefix + (u,)]
                    n.ns = log_add(n.ns, ps.score() + prob)
                    if n.v_ns < ps.viterbi_score() + prob:
                        n.v_ns = ps.viterbi_score() + prob
                        n.cur_token_prob = prob
                        n.times_ns = ps.times().copy()
                        n.times_ns.append(t)
        cur = sorted(nxt.items(), key=lambda kv: kv[1].score(), reverse=True)[:beam_size]
    return [list(p) for p, _ in cur], [sc.score() for _, sc in cur], [sc.times() for _, sc in cur]
