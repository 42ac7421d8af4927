"""HOMP (src/codec_homp.c) and T0 (src/codec_t0.c): Ultima's homopolymer codecs.  CPU: the restatement against the reference's compiled
objects (oracle/_ref), both directions.  GPU (-m gpu, also --simt): gzb_homp_condense / gzb_homp_expand against both."""
import numpy as np
import pytest

import orc

HOMP, T0 = 0, 1


def ultima_like(n_lines, seed, mode, noise=0.03, read_len=(20, 200)):
    """reads with long homopolymer runs; a quality (resp. t0) string that follows Ultima's rule per run — a palindrome that ends in 'I's once an 'I'
    appears (resp. one character per run) — broken now and then"""
    rng = np.random.default_rng(seed)
    parts, so, qo, lens = [np.frombuffer(b"header\n", np.uint8)], [], [], []
    pos = parts[0].size
    for _ in range(n_lines):
        L = int(rng.integers(read_len[0], read_len[1]))
        seq, s = [], []
        while len(seq) < L:
            h = min(L - len(seq), int(rng.choice([1, 1, 1, 2, 3, 4, 6, 9, 14])))
            seq += [int(rng.choice(np.frombuffer(b"ACGT", np.uint8)))] * h
            if mode == T0:
                run = [int(rng.integers(48, 58))] * h
            else:
                half = [int(rng.choice(np.frombuffer(b"5:?DI", np.uint8), p=[.1, .15, .2, .25, .3])) for _ in range((h + 1) // 2)]
                for k in range(1, len(half)):
                    if half[k - 1] == ord("I"):
                        half[k] = ord("I")
                run = half + half[:h // 2][::-1]
            if h > 1 and rng.random() < noise:
                run[int(rng.integers(0, h))] = ord("#")
            s += run
        seq, s = np.array(seq[:L], np.uint8), np.array(s[:L], np.uint8)
        qo.append(pos); parts.append(seq); pos += L
        so.append(pos); parts.append(s); pos += L
        lens.append(L)
    return np.concatenate(parts), np.array(so, np.uint64), np.array(lens, np.uint32), np.array(qo, np.uint64)


@pytest.mark.parametrize("mode", [HOMP, T0])
@pytest.mark.parametrize("seed", [1, 2])
def test_port_matches_reference(mode, seed):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    txt, so, sl, qo = ultima_like(300, seed, mode, noise=0.03 if seed == 1 else 0.5)
    p = orc.hp_condense(mode, txt, so, sl, qo, "port")
    r = orc.hp_condense(mode, txt, so, sl, qo, "ref")
    assert np.array_equal(p[0], r[0]) and np.array_equal(p[1], r[1])
    assert p[0].size < sl.sum()                                        # something was condensed
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(so, sl)])
    for lib in ("port", "ref"):
        back = orc.hp_expand(mode, p[0], txt, qo, sl, lib)
        assert back is not None and np.array_equal(back[0], want), lib
    assert orc.hp_expand(mode, p[0][:-1], txt, qo, sl, "port") is None   # one byte short


def test_port_homp_short_and_missing():
    """a quality of length <= 1 is left alone (:137); on the way back a ' ' is a line without quality (:241-244)"""
    txt = np.frombuffer(b"AACC ?GGGTT55I55", np.uint8).copy()
    so = np.array([4, 5, 11], np.uint64); sl = np.array([1, 1, 5], np.uint32); qo = np.array([0, 2, 6], np.uint64)
    p = orc.hp_condense(HOMP, txt, so, sl, qo, "port")
    assert bytes(p[0][:2]) == b" ?" and list(p[1][:2]) == [1, 1]
    if orc.have_gz_ref():
        r = orc.hp_condense(HOMP, txt, so, sl, qo, "ref")
        assert np.array_equal(p[0], r[0])
    lens = np.array([4, 1, 5], np.uint32)                               # the first line has no quality: its SEQ has 4 bases
    back = orc.hp_expand(HOMP, p[0], txt, qo, lens, "port")
    assert back is not None and back[1][0] == 1 and back[0][0] == ord("*") and bytes(back[0][4:5]) == b"?" and bytes(back[0][5:10]) == bytes(txt[11:16])


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [HOMP, T0])
def test_gpu_homp_t0(eng, mode):
    cases = [ultima_like(500, 5, mode), ultima_like(129, 6, mode, noise=0.6), ultima_like(64, 7, mode, read_len=(1, 12)), ultima_like(40, 8, mode, read_len=(900, 2500))]
    got = eng.hp_condense(mode, cases)
    for (txt, so, sl, qo), g in zip(cases, got):
        w = orc.hp_condense(mode, txt, so, sl, qo, "port")
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]), "GPU != restatement"
        if orc.have_gz_ref():
            r = orc.hp_condense(mode, txt, so, sl, qo, "ref")
            assert np.array_equal(g[0], r[0]), "GPU != reference"
    back = eng.hp_expand(mode, [(g[0], c[0], c[3], c[2]) for c, g in zip(cases, got)])
    for (txt, so, sl, qo), b in zip(cases, back):
        want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(so, sl)])
        assert np.array_equal(b[0], want), "GPU expand mismatch"
    from genozip_b200.lib import GzbError
    with pytest.raises(GzbError):
        eng.hp_expand(mode, [(got[0][0][:-1], cases[0][0], cases[0][3], cases[0][2])])
