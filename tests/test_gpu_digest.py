"""gzb_adler32_batch against zlib's adler32 — the function the reference calls for z_digest (src/compressor.c:151,161; its adler32 is
the vendored zlib / libdeflate one, same definition): empty, tiny, misaligned, chunk-boundary and multi-chunk buffers in one batch,
host and device pointers."""
import zlib
import numpy as np, pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _bufs():
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 15, 16, 17, 255, 4096, 65535, 65536, 65537, 131072, 200001, 1 << 20, (3 << 20) + 7]
    out = [rng.integers(0, 256, n, dtype=np.uint8) for n in sizes]
    out.append(np.full(500000, 255, np.uint8))                             # the largest sums: the modular folding must not overflow
    out.append(np.zeros(70000, np.uint8))
    big = rng.integers(0, 256, (1 << 20) + 64, dtype=np.uint8)
    out += [big[k:k + 100000 + k] for k in (1, 3, 7, 13)]                    # misaligned starts
    return out


def test_adler32_batch_matches_zlib(engine):
    bufs = _bufs()
    got = engine.adler32(bufs)
    want = [zlib.adler32(b.tobytes(), 1) & 0xffffffff for b in bufs]
    assert got == want


def test_adler32_of_compressed_sections(engine):
    """z_digest of section bodies as comp_compress computes it: adler32 of exactly the bytes the codec wrote"""
    from datagen import stream
    secs = [("RANB", stream("qual", 100000, 1)), ("ARTb", stream("skew8", 30000, 2)), ("RANW", stream("u32le", 50000, 3))]
    comp = engine.compress(secs)
    assert engine.adler32(comp) == [zlib.adler32(c.tobytes(), 1) & 0xffffffff for c in comp]
