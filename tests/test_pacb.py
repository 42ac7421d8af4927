"""PACB (src/codec_pacb.c), PacBio's quality codec: QUAL multiplexed by the read's number of passes and the base's surroundings.
CPU: the restatement against the reference's compiled codec_pacb.c (oracle/_ref), both directions, with and without np:i, incl. reads without
quality.  GPU (-m gpu, also --simt): gzb_pacb_mux / gzb_pacb_demux against both."""
import numpy as np
import pytest

import orc


def hifi_like(n_lines, seed, max_np, missing=False, read_len=(30, 400)):
    rng = np.random.default_rng(seed)
    parts, qoff, soff, ql, sl, np0 = [np.frombuffer(b"@PG\n", np.uint8)], [], [], [], [], []
    pos = parts[0].size
    for _ in range(n_lines):
        L = int(rng.integers(read_len[0], read_len[1]))
        seq = []
        while len(seq) < L:
            seq += [int(rng.choice(np.frombuffer(b"ACGT", np.uint8)))] * int(rng.choice([1, 1, 1, 2, 2, 3, 5]))
        seq = np.array(seq[:L], np.uint8)
        n0 = int(rng.integers(0, max_np))
        qual = (33 + np.clip(20 + 3 * n0 + rng.integers(0, 10, L) - 8 * (np.concatenate([[0], seq[1:] == seq[:-1]])), 0, 60)).astype(np.uint8)
        miss = missing and rng.random() < 0.1
        soff.append(pos); parts.append(seq); pos += L
        qoff.append(pos)
        if miss:
            parts.append(np.frombuffer(b" ", np.uint8)); pos += 1; ql.append(1)
        else:
            parts.append(qual); pos += L; ql.append(L)
        sl.append(L); np0.append(n0)
    return np.concatenate(parts), np.array(qoff, np.uint64), np.array(ql, np.uint32), np.array(soff, np.uint64), np.array(sl, np.uint32), np.array(np0, np.uint8)


@pytest.mark.parametrize("seed,max_np,missing", [(1, 1, False), (2, 12, False), (3, 12, True)])
def test_port_matches_reference(seed, max_np, missing):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    txt, qoff, ql, soff, sl, np0 = hifi_like(200, seed, max_np, missing)
    n0 = np0 if max_np > 1 else None
    p = orc.pacb_mux(txt, qoff, ql, soff, n0, max_np, "port")
    r = orc.pacb_mux(txt, qoff, ql, soff, n0, max_np, "ref")
    assert np.array_equal(p[1], r[1]) and np.array_equal(p[0], r[0])
    assert (p[1][:7 * max_np] > 0).sum() >= 6                         # the channels are in use
    out_off = np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64)
    bp = orc.pacb_demux(txt, soff, sl, n0, max_np, out_off, int(sl.sum()), p[0], p[1], "port")
    br = orc.pacb_demux(txt, soff, sl, n0, max_np, out_off, int(sl.sum()), p[0], p[1], "ref")
    if missing:
        # the ' ' of a read without quality is stored under K of a ONE-base sequence (:135) but fetched under K of the whole read (:306): when the
        # read starts with a homopolymer the reference's own decoder looks in another channel.  Restatement and reference must agree on the outcome.
        assert (bp is None) == (br is None) and (bp is None or np.array_equal(bp, br))
        return
    assert bp is not None and br is not None
    for o, n, q, a in zip(out_off, sl, qoff, ql):
        o, n = int(o), int(n)
        if a == 1 and n != 1:
            assert bp[o] == ord("*") and br[o] == ord("*")
        else:
            assert np.array_equal(bp[o:o + n], txt[int(q):int(q) + n]) and np.array_equal(br[o:o + n], txt[int(q):int(q) + n])


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_pacb(eng):
    raw = [hifi_like(300, 11, 1), hifi_like(257, 12, 12), hifi_like(40, 13, 12, read_len=(1, 6)), hifi_like(60, 14, 12, missing=True)]
    cases = [(t, qo, ql, so, (n0 if mnp > 1 else None), mnp) for (t, qo, ql, so, sl, n0), mnp in zip(raw, (1, 12, 12, 12))]
    got = eng.pacb_mux(cases)
    for c, g in zip(cases, got):
        w = orc.pacb_mux(*c, lib="port")
        assert np.array_equal(g[1], w[1]) and np.array_equal(g[0], w[0]), "GPU != restatement"
        if orc.have_gz_ref():
            r = orc.pacb_mux(*c, lib="ref")
            assert np.array_equal(g[1], r[1]) and np.array_equal(g[0], r[0]), "GPU != reference codec_pacb.c"
    items, wants = [], []
    for (txt, qoff, ql, soff, sl, np0), c, g in list(zip(raw, cases, got))[:3]:
        out_off = np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64)
        items.append((txt, soff, sl, c[4], c[5], out_off, int(sl.sum()), g[0], g[1]))
        wants.append(np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qoff, ql)]))
    for b, w in zip(eng.pacb_demux(items), wants):
        assert np.array_equal(b, w), "GPU PACB demux mismatch"
    from genozip_b200.lib import GzbError
    txt, qoff, ql, soff, sl, np0 = raw[3]                              # reads without quality: refused by the bulk form
    with pytest.raises(GzbError):
        eng.pacb_demux([(txt, soff, sl, np0, 12, np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64), int(sl.sum()), got[3][0], got[3][1])])
