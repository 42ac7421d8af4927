"""zip_generate_local's in-place transforms (src/zip.c:167-213) and their PIZ inverses: the numpy restatement pinned against the reference's
own INTERLACE / DEINTERLACE / BGEN macros (oracle/_ref), the CUDA path (gzb_local_transform_batch) against both — every width, the
extreme values of every signed type, empty and odd-sized buffers, several buffers in one batch."""
import numpy as np, pytest
import orc

S = {8: np.int8, 16: np.int16, 32: np.int32, 64: np.int64}
U = {16: np.uint16, 32: np.uint32, 64: np.uint64}


def _cases():
    rng = np.random.default_rng(11)
    out = []
    for bits, dt in S.items():
        info = np.iinfo(dt)
        edge = np.array([0, 1, -1, 2, -2, info.max, info.min, info.max - 1, info.min + 1], dt)
        rnd = rng.integers(info.min, info.max, 70001, dtype=dt, endpoint=True)
        small = rng.integers(-40, 40, 1000).astype(dt)
        for a in (edge, rnd, small, np.zeros(0, dt)):
            out.append((f"interlace{bits}", a))
            out.append((f"deinterlace{bits}", a.view({8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}[bits])))
    for bits, dt in U.items():
        out.append((f"swap{bits}", rng.integers(0, np.iinfo(dt).max, 200003, dtype=dt, endpoint=True)))
    return out


def test_restatement_is_the_reference():
    if not orc.have_gz_ref():
        pytest.skip("oracle/_ref/libgz_ref.so not built here")
    for op, a in _cases():
        assert np.array_equal(orc.local_transform(op, a).view(np.uint8), orc.ref_local_transform(op, a).view(np.uint8)), op
    for bits, dt in S.items():                                              # PIZ undoes ZIP
        a = np.random.default_rng(bits).integers(np.iinfo(dt).min, np.iinfo(dt).max, 5000, dtype=dt, endpoint=True)
        assert np.array_equal(orc.local_transform(f"deinterlace{bits}", orc.local_transform(f"interlace{bits}", a)).view(dt), a)


@pytest.mark.gpu
def test_gpu_local_transforms():
    from genozip_b200 import Engine
    eng = Engine(0)
    cases = _cases()
    got = eng.local_transform(cases)
    for (op, a), g in zip(cases, got):
        want = orc.ref_local_transform(op, a) if orc.have_gz_ref() else orc.local_transform(op, a)
        assert np.array_equal(g.view(np.uint8), want.view(np.uint8)), op
    eng.close()
