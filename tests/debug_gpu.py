import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, orc
from genozip_b200 import Engine
eng = Engine(0)
rng = np.random.default_rng(99)
cases = []
for t in range(120):
    big = t % 3 == 0
    k = int(rng.integers(8, 60)) if big else int(rng.integers(2, 40))
    n = int(rng.integers(50000, 300000)) if big else int(rng.integers(64, 6000))
    p = rng.dirichlet(np.full(k, 0.05 if big else rng.uniform(0.05, 2.0))) + 1e-5
    p /= p.sum()
    cases.append(("RANB", rng.choice(np.arange(k, dtype=np.uint8) + 40, size=n, p=p).astype(np.uint8)))
comp = eng.compress(cases)
outs = eng.uncompress([(nm, c, d.size) for (nm, d), c in zip(cases, comp)])
sizes = sorted([d.size for _, d in cases], reverse=True)
for i, ((nm, d), o) in enumerate(zip(cases, outs)):
    if not np.array_equal(o, d):
        bad = np.where(o != d)[0]
        print("case", i, "n", d.size, "q4", d.size >> 2, "r", d.size & 3, "nbad", bad.size, "first", bad[:12], "last", bad[-5:], "rank in sizes", sizes.index(d.size), "flags", hex(comp[i][0]))
        print("  got ", o[bad[:12]], "want", d[bad[:12]])
print("neighbours by size:", [s for s in sizes if 3000 < s < 6000])
