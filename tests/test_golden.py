"""Golden vectors (tests/golden/hts_golden.json, produced by tests/golden/make_golden.py from the reference's own compiled
htscodecs): the CPU restatement must reproduce every one of them — with or without oracle/_ref at hand — and so must the
CUDA path (-m gpu)."""
import hashlib, json, os
import numpy as np, pytest
import orc
from datagen import stream

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hts_golden.json")))["cases"]
KIND = {"R": "rans", "A": "arith"}


def _check(c, got):
    assert got.size == c["len"], (c["codec"], c["kind"], c["n"], got.size, c["len"])
    assert got[:24].tobytes().hex() == c["head"], (c["codec"], c["kind"], c["n"])
    assert hashlib.sha256(got.tobytes()).hexdigest() == c["sha256"], (c["codec"], c["kind"], c["n"])


@pytest.mark.parametrize("codec", ["RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw"])
def test_restatement_reproduces_golden_vectors(codec):
    n = 0
    for c in GOLD:
        if c["codec"] != codec:
            continue
        data = stream(c["kind"], c["n"], c["seed"])
        _check(c, orc.compress("port", KIND[codec[0]], data, orc.ORDER[codec]))
        n += 1
    assert n > 100


@pytest.mark.gpu
def test_cuda_path_reproduces_golden_vectors(pytestconfig):
    from genozip_b200 import Engine
    eng = Engine(0)
    try:
        gold = GOLD[::3]                                     # every third vector: all codecs, kinds and sizes still occur
        if pytestconfig.getoption("--simt") and os.environ.get("GZB_SIMT_QUICK"):
            gold = GOLD[::17]                                # (the emulator's quick pass inside the CPU suite)
        items = [(c["codec"], stream(c["kind"], c["n"], c["seed"])) for c in gold]
        for i in range(0, len(items), 256):
            got = eng.compress(items[i:i + 256])
            for c, g in zip(gold[i:i + 256], got):
                _check(c, g)
            sel = [k for k, c in enumerate(gold[i:i + 256]) if c["n"]]
            back = eng.uncompress([(gold[i + k]["codec"], got[k], gold[i + k]["n"]) for k in sel])
            for k, b in zip(sel, back):
                assert np.array_equal(b, items[i + k][1])
    finally:
        eng.close()
