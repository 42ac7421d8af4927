"""SMUX (src/codec_smux.c), MGI's quality codec: QUAL multiplexed by the base at the same position into 5 channels.
CPU: the restatement against the reference's compiled codec_smux.c (oracle/_ref), both directions, incl. reverse-complemented reads and
reads without quality.  GPU (-m gpu, also --simt): gzb_smux_mux / gzb_smux_demux against both."""
import numpy as np
import pytest

import orc


def mgi_like(n_lines, seed, sam=False, n_frac=0.01, n_mono=True, missing=False, read_len=(20, 160)):
    rng = np.random.default_rng(seed)
    parts, qoff, soff, lens, rev = [np.frombuffer(b"@HD\n", np.uint8)], [], [], [], []
    pos = parts[0].size
    for li in range(n_lines):
        L = int(rng.integers(read_len[0], read_len[1]))
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        isn = rng.random(L) < n_frac
        seq = np.where(isn, ord("N"), seq).astype(np.uint8)
        qual = (33 + np.clip(seq.astype(np.int32) % 7 * 5 + rng.integers(0, 8, L), 0, 60)).astype(np.uint8)
        qual = np.where(isn, ord("!") if n_mono else rng.integers(33, 36, L), qual).astype(np.uint8)
        r = bool(sam and rng.random() < 0.4)
        miss = missing and sam and rng.random() < 0.1
        soff.append(pos); parts.append(seq); pos += L
        qoff.append(pos)
        if miss:
            parts.append(np.frombuffer(b" ", np.uint8)); pos += 1
        else:
            parts.append(qual); pos += L
        lens.append((1 if miss else L, L)); rev.append(r)
    txt = np.concatenate(parts)
    ql = np.array([a for a, _ in lens], np.uint32); sl = np.array([b for _, b in lens], np.uint32)
    return txt, np.array(qoff, np.uint64), ql, np.array(soff, np.uint64), sl, (np.array(rev, np.uint8) if sam else None)


def present(chan, count, n_param):
    """what goes to the file: the fifth channel is dropped when it is monochar"""
    cnt = count.copy()
    if n_param:
        chan = chan[:int(count[:4].sum())]; cnt[4] = 0
    return chan, cnt


@pytest.mark.parametrize("seed,sam,n_mono,missing", [(1, False, True, False), (2, True, True, False), (3, True, False, False), (4, True, True, True)])
def test_port_matches_reference(seed, sam, n_mono, missing):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    txt, qoff, ql, soff, sl, rev = mgi_like(300, seed, sam, n_mono=n_mono, missing=missing)
    p = orc.smux_mux(txt, qoff, ql, soff, sl, rev, "port")
    r = orc.smux_mux(txt, qoff, ql, soff, sl, rev, "ref")
    assert np.array_equal(p[0], r[0]) and np.array_equal(p[1], r[1]) and p[2] == r[2]
    if not missing:                                                      # (a read without quality may put its blank into the fifth channel)
        assert bool(p[2]) == n_mono
    ch, cnt = present(*p)
    out_off = np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64)
    back_p = orc.smux_demux(txt, soff, sl, rev, out_off, int(sl.sum()), ch, cnt, p[2], "port")
    back_r = orc.smux_demux(txt, soff, sl, rev, out_off, int(sl.sum()), ch, cnt, p[2], "ref")
    assert back_p is not None and back_r is not None
    for o, n, q, a, b in zip(out_off, sl, qoff, ql, range(len(sl))):
        o, n = int(o), int(n)
        if a == 1 and n != 1:                                            # a read without quality: '*'
            assert back_p[o] == ord("*") and back_r[o] == ord("*")
        else:
            assert np.array_equal(back_p[o:o + n], txt[int(q):int(q) + n]) and np.array_equal(back_r[o:o + n], txt[int(q):int(q) + n])


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_smux(eng):
    cases = [mgi_like(400, 11), mgi_like(333, 12, sam=True), mgi_like(64, 13, sam=True, n_mono=False, read_len=(1, 40)), mgi_like(50, 14, sam=True, missing=True)]
    got = eng.smux_mux(cases)
    for c, g in zip(cases, got):
        w = orc.smux_mux(*c, lib="port")
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and g[2] == w[2], "GPU != restatement"
        if orc.have_gz_ref():
            r = orc.smux_mux(*c, lib="ref")
            assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1]) and g[2] == r[2], "GPU != reference codec_smux.c"
    items, wants = [], []
    for (txt, qoff, ql, soff, sl, rev), g in list(zip(cases, got))[:3]:
        ch, cnt = present(*g)
        out_off = np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64)
        items.append((txt, soff, sl, rev, out_off, int(sl.sum()), ch, cnt, g[2]))
        wants.append(np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qoff, ql)]))
    for b, w in zip(eng.smux_demux(items), wants):
        assert np.array_equal(b, w), "GPU SMUX demux mismatch"
    from genozip_b200.lib import GzbError
    txt, qoff, ql, soff, sl, rev = cases[3]                              # reads without quality: refused by the bulk form
    ch, cnt = present(*got[3])
    with pytest.raises(GzbError):
        eng.smux_demux([(txt, soff, sl, rev, np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64), int(sl.sum()), ch, cnt, got[3][2])])
