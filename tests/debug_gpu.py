import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, orc
from datagen import stream, KINDS
from genozip_b200 import Engine
eng = Engine(0)
def show(name, data, tag):
    print("case", tag, name, data.size, flush=True)
    c = eng.compress([(name, data)])[0]
    kind = "rans" if name.startswith("RAN") else "arith"
    w = orc.compress("ref" if orc.have_ref() else "port", kind, data, orc.ORDER[name])
    ok = c.size == w.size and np.array_equal(c, w)
    if not ok:
        m = min(c.size, w.size); d = int(np.argmax(c[:m] != w[:m])) if m and (c[:m] != w[:m]).any() else m
        print(f"{tag} {name} n={data.size} got_len={c.size} want_len={w.size} first_diff={d}")
        print("  got ", c[max(0,d-14):d+10].tobytes().hex())
        print("  want", w[max(0,d-14):d+10].tobytes().hex())
    return ok
for name in ("ARTB", "ARTb"):
    for dk in KINDS:
        for n in (777, 50021):
            show(name, stream(dk, n, 11), dk)
rng = np.random.default_rng(1)
for i in range(200):
    name = ["RANB","RANW","RANb","RANw","ARTB","ARTW","ARTb","ARTw"][i % 8]
    dk = KINDS[int(rng.integers(0, len(KINDS)))]
    n = int(rng.integers(1, 40000))
    show(name, stream(dk, n, 100 + i), f"mixed{i}:{dk}")
print("done")
