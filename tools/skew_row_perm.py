import random, itertools
PITCH = 144
def lds_wavefronts(pi):
    tot = 0
    for k in range(32):
        banks = {}
        for lam in range(32):
            byte = pi[lam] * PITCH + 2 * (31 - lam + k)
            w = byte >> 2
            banks.setdefault(w % 32, set()).add(w)
        tot += max(len(v) for v in banks.values())
    return tot
def sts_wavefronts(pi):
    tot = 0
    for q in range(8):
        for grp in range(4):
            banks = {}
            for lam in range(grp * 8, grp * 8 + 8):
                byte = pi[lam] * PITCH + 16 * q
                for j in range(4):
                    w = (byte >> 2) + j
                    banks.setdefault(w % 32, set()).add(w)
            tot += max(len(v) for v in banks.values())
    return tot
ident = list(range(32))
print("identity: lds", lds_wavefronts(ident), "(ideal 32)  sts", sts_wavefronts(ident), "(ideal 32)")
best = None
random.seed(1)
# structured: pi = groups of 8 with mod-8 bijection: pi[8u + j] = 8*gperm[u] + rot(j, u)
for gperm in itertools.permutations(range(4)):
    for trial in range(3000):
        pi = []
        for u in range(4):
            p8 = list(range(8)); random.shuffle(p8)
            pi += [8 * gperm[u] + x for x in p8]
        c = lds_wavefronts(pi)
        if best is None or c < best[0]:
            best = (c, pi[:], sts_wavefronts(pi))
print("best random structured:", best)
# local search from best
cur = best[1][:]; curc = best[0]
for it in range(200000):
    a, b = random.sample(range(32), 2)
    if a // 8 != b // 8 and (cur[a] % 8 != cur[b] % 8): continue
    cur[a], cur[b] = cur[b], cur[a]
    if sts_wavefronts(cur) > 32: cur[a], cur[b] = cur[b], cur[a]; continue
    c = lds_wavefronts(cur)
    if c <= curc: curc = c
    else: cur[a], cur[b] = cur[b], cur[a]
print("after local search:", curc, cur, sts_wavefronts(cur))
