"""gzb_local_transpose_batch: dyn_int_transpose (src/dyn_int.c:45-105) and its PIZ inverse BGEN_transpose_u*_buf (src/buffer.c:364-391).
CPU: the restatement against the reference's compiled dyn_int.c (oracle/_ref); the PIZ direction as the inverse of that plus the byte swap.
GPU (-m gpu, also --simt): the kernel against both."""
import numpy as np
import pytest

import orc

CASES = [(np.uint8, 1000, 7), (np.uint16, 333, 255), (np.uint32, 4097, 33), (np.uint32, 1, 1), (np.uint8, 64, 64), (np.uint16, 31, 100),
         (np.uint32, 100, 1)]


def mats(seed=3):
    rng = np.random.default_rng(seed)
    return [(rng.integers(0, np.iinfo(dt).max, rows * cols, dtype=dt), cols) for dt, rows, cols in CASES]


def test_port_matches_reference():
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    for a, cols in mats():
        p, pt = orc.local_transpose(a, cols)
        r, rt = orc.ref_dyn_int_transpose(a, cols)
        assert pt and rt and np.array_equal(p, r), (a.dtype, a.size, cols)
        # the number of samples of the VCF header as the column count (local.n_cols = 0, dyn_int.c:59-62): not limited to 255
        r2, rt2 = orc.ref_dyn_int_transpose(a, 0, cols)
        assert rt2 and np.array_equal(p, r2)
    # not a rectangle: left alone (:75-78)
    a = np.arange(1000, dtype=np.uint16)
    for fn in (lambda x: orc.local_transpose(x, 7), lambda x: orc.ref_dyn_int_transpose(x, 7)):
        out, tr = fn(a)
        assert not tr and np.array_equal(out, a)


def test_port_piz_is_the_inverse():
    for a, cols in mats(5):
        t, _ = orc.local_transpose(a, cols)
        be = t.byteswap()                                   # what is in the file: big endian (BGEN_u*_buf ran before the transpose, zip.c:167-214)
        back, ok = orc.local_transpose(be, cols, piz=True)
        assert ok and np.array_equal(back, a)


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_transpose(eng):
    items = mats(7) + [(np.arange(1000, dtype=np.uint16), 7), (np.zeros(0, np.uint8), 5)]
    got = eng.local_transpose(items)
    for (a, cols), (g, tr) in zip(items, got):
        w, wt = orc.local_transpose(a, cols)
        assert tr == wt and np.array_equal(g, w), (a.dtype, a.size, cols)
        if orc.have_gz_ref() and a.size:
            r, rt = orc.ref_dyn_int_transpose(a, cols)
            assert tr == rt and np.array_equal(g, r)
    # PIZ: from the file's bytes back to the lines' order and the host's endianness
    piz_items = [(orc.local_transpose(a, cols)[0].byteswap(), cols) for a, cols in mats(7)]
    back = eng.local_transpose(piz_items, piz=True)
    for (a, cols), (b, tr) in zip(mats(7), back):
        assert tr and np.array_equal(b, a)
    from genozip_b200.lib import GzbError
    with pytest.raises(GzbError):
        eng.local_transpose([(np.arange(1000, dtype=np.uint16), 7)], piz=True)
