"""debug helper (GPU box): low-entropy order-1 rANS streams through the hot-transition path vs the oracle"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc
from genozip_b200 import Engine

eng = Engine(0)
rng = np.random.default_rng(5)
bad = 0
for n in (100, 1000, 4099, 20000, 450001, 3000000):
    for p, sym in ((0.001, 78), (0.02, 1), (0.0, 0), (0.3, 5)):
        x = np.zeros(n, np.uint8)
        x[rng.random(n) < p] = sym
        if p == 0.02:
            x[rng.random(n) < 0.001] = 200
        for codec in ("RANB", "RANb", "RANW"):
            want = orc.compress("port", "rans", x, orc.ORDER[codec])
            got = eng.compress([(codec, x)])[0]
            ok = np.array_equal(got, want)
            out = eng.uncompress([(codec, want, n)])[0]
            ok2 = np.array_equal(out, x)
            if not (ok and ok2):
                bad += 1
            print(n, p, codec, "enc", ok, "dec", ok2, got.size, want.size, flush=True)
# several leaves per warp job + a big one, mixed hot / not hot
items = []
for i in range(40):
    n = int(rng.integers(50, 60000))
    x = np.zeros(n, np.uint8) if i % 3 else rng.integers(0, 9, n).astype(np.uint8)
    x[rng.random(n) < 0.002] = 78
    items.append(("RANB", x))
got = eng.compress(items)
for (c, x), g in zip(items, got):
    w = orc.compress("port", "rans", x, orc.ORDER[c])
    if not np.array_equal(g, w):
        bad += 1; print("batch enc mismatch", x.size)
outs = eng.uncompress([(c, g, x.size) for (c, x), g in zip(items, got)])
for (c, x), o in zip(items, outs):
    if not np.array_equal(o, x):
        bad += 1; print("batch dec mismatch", x.size)
print("bad", bad)
