"""gzb_assign_codecs: codec_assign_best_codec's size criterion (src/codec.c:234-389) — the sample of every stream compressed with the
eight simple codecs in one batch, sizes only.  The sizes must be those of the reference's own test compressions (oracle/_ref = the
reference's htscodecs objects), the choice the smallest section (ties to the lower Codec value, CODEC_NONE against the bare sample)."""
import numpy as np
import pytest

import orc
from datagen import stream

pytestmark = pytest.mark.gpu

NAMES = ("RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw")
ID = {"NONE": 1, "RANB": 6, "RANW": 7, "RANb": 8, "RANw": 9, "ARTB": 16, "ARTW": 17, "ARTb": 18, "ARTw": 19}


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


def _want(data):
    sample = data[:99999]
    if sample.size < 50:
        return None, {}
    which = "ref" if orc.have_ref() else "port"
    sizes = {nm: int(orc.compress(which, "rans" if nm.startswith("RAN") else "arith", sample, orc.ORDER[nm]).size) for nm in NAMES}
    best, best_size = "NONE", sample.size
    for nm in NAMES:                                   # ascending Codec value: the first of equal sizes stays
        if sizes[nm] + 28 < best_size:
            best, best_size = nm, sizes[nm] + 28
    return best, sizes


def test_assign_matches_reference_sizes(eng):
    rng = np.random.default_rng(5)
    bufs = [stream("qual", 150000, 1),                 # longer than the sample: only the first 99,999 bytes count
            stream("u32le", 30000, 2), stream("skew8", 20000, 3), stream("text", 5000, 4),
            rng.integers(0, 256, 4000, dtype=np.uint8),                          # incompressible: CODEC_NONE
            np.zeros(49, np.uint8), np.zeros(50, np.uint8), np.zeros(0, np.uint8),   # below / at MIN_LEN_FOR_COMPRESSION, empty
            np.tile(np.arange(4, dtype=np.uint8), 3000)]                          # PACK territory
    got = eng.assign_codecs(bufs)
    assert len(got) == len(bufs)
    for b, (best, sizes) in zip(bufs, got):
        wbest, wsizes = _want(b)
        assert best == wbest, (b.size, best, wbest, sizes, wsizes)
        assert sizes == wsizes
    assert got[4][0] == "NONE" and got[5][0] is None and got[7][0] is None and got[6][0] is not None
