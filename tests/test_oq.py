"""OQ codec (src/codec_oq.c): the original qualities of a read multiplexed by its QUAL into 94 channels.
CPU: the restatement (oracle/gz_port.c) against the reference's compiled codec_oq.c (oracle/_ref), both directions.
GPU (-m gpu, also --simt): gzb_oq_mux / gzb_oq_demux against both."""
import numpy as np
import pytest

import orc


def sam_like(n_lines, seed, read_len=(30, 160), mono_channels=True, bam_terms=False):
    """a text with, per line, a QUAL string and an OQ string of the same length somewhere in it; OQ depends mostly on QUAL (recalibration),
    one channel exactly (monochar), some noise elsewhere"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(read_len[0], read_len[1], n_lines).astype(np.uint32)
    parts, qoff, ooff = [np.frombuffer(b"@HD\tVN:1.6\n", np.uint8)], [], []
    pos = parts[0].size
    for L in lens:
        L = int(L)
        qual = rng.choice(np.frombuffer(b"#,:F", np.uint8), L, p=[0.02, 0.08, 0.1, 0.8])
        oq = (qual.astype(np.int32) - rng.integers(0, 3, L) * (qual != ord("#"))).astype(np.uint8)   # '#' -> always '#': a monochar channel
        if not mono_channels:
            oq = np.where(rng.random(L) < 0.3, ord("5"), oq).astype(np.uint8)
        filler = rng.integers(65, 91, int(rng.integers(0, 20))).astype(np.uint8)
        qoff.append(pos); parts.append(qual); pos += L
        parts.append(filler); pos += filler.size
        ooff.append(pos); parts.append(oq); pos += L
    txt = np.concatenate(parts)
    return txt, np.array(qoff, np.uint64), lens, np.array(ooff, np.uint64)


def present(chan, count, mono):
    """what goes to the file: the channels that are not monochar, back to back; their counts"""
    cnt = np.where(mono != 0, 0, count).astype(np.uint32)
    at = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
    keep = [chan[at[q]:at[q + 1]] for q in range(94) if cnt[q]]
    return (np.concatenate(keep) if keep else np.zeros(0, np.uint8)), cnt


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_port_matches_reference(seed):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    txt, qoff, lens, ooff = sam_like(400, seed, mono_channels=seed != 3)
    seq_len = lens.copy()
    if seed == 2:
        seq_len[::7] = 0                                # lines the mux pass skips (:94) while the count pass counts them (:61-72)
    p = orc.oq_mux(txt, qoff, lens, ooff, seq_len, "port")
    r = orc.oq_mux(txt, qoff, lens, ooff, seq_len, "ref")
    assert p is not None and r is not None
    for a, b, nm in zip(p, r, ("channels", "count", "monochars")):
        assert np.array_equal(a, b), nm
    if seed == 1:
        assert p[2][ord("#") - 33] == ord("#")          # the monochar channel was found
    if seed == 2:
        return                                          # (skipped lines cannot be reconstructed: their OQ was never stored)
    ch, cnt = present(*p)
    out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(ooff, lens)])
    for lib in ("port", "ref"):
        got = orc.oq_demux(txt, qoff, lens, out_off, int(lens.sum()), 33, ch, cnt, p[2], lib)
        assert got is not None and np.array_equal(got, want), lib
    # a channel one byte short: "channel is out of data" (:152)
    k = int(np.flatnonzero(cnt)[0]); cnt2 = cnt.copy(); cnt2[k] -= 1
    at = int(cnt[:k].sum())
    ch2 = np.concatenate([ch[:at + int(cnt2[k])], ch[at + int(cnt[k]):]])
    assert orc.oq_demux(txt, qoff, lens, out_off, int(lens.sum()), 33, ch2, cnt2, p[2], "port") is None
    assert orc.oq_demux(txt, qoff, lens, out_off, int(lens.sum()), 33, ch2, cnt2, p[2], "ref") is None


def test_port_bam_terms():
    """keys as BAM values (sam_diff = 0, :131)"""
    txt, qoff, lens, ooff = sam_like(100, 9)
    p = orc.oq_mux(txt, qoff, lens, ooff, None, "port")
    ch, cnt = present(*p)
    t2 = txt.copy()
    for o, l in zip(qoff, lens):
        t2[int(o):int(o) + int(l)] -= 33
    out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(ooff, lens)])
    libs = ("port", "ref") if orc.have_gz_ref() else ("port",)
    for lib in libs:
        assert np.array_equal(orc.oq_demux(t2, qoff, lens, out_off, int(lens.sum()), 0, ch, cnt, p[2], lib), want)


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_oq_batch(eng):
    cases = [sam_like(300, 11), sam_like(1000, 12, mono_channels=False), sam_like(64, 13, read_len=(1, 40)), sam_like(33, 14, read_len=(500, 3000))]
    seqs = [c[2].copy() for c in cases]
    seqs[1][::5] = 0
    got = eng.oq_mux([(c[0], c[1], c[2], c[3], s) for c, s in zip(cases, seqs)])
    for (txt, qoff, lens, ooff), s, g in zip(cases, seqs, got):
        w = orc.oq_mux(txt, qoff, lens, ooff, s, "port")
        for a, b, nm in zip(g, w, ("channels", "count", "monochars")):
            assert np.array_equal(a, b), f"{nm}: GPU != restatement"
        if orc.have_gz_ref():
            r = orc.oq_mux(txt, qoff, lens, ooff, s, "ref")
            for a, b, nm in zip(g, r, ("channels", "count", "monochars")):
                assert np.array_equal(a, b), f"{nm}: GPU != reference codec_oq.c"
    # back: every VBlock whose lines were all multiplexed
    items, wants = [], []
    for i, ((txt, qoff, lens, ooff), g) in enumerate(zip(cases, got)):
        if i == 1:
            continue
        ch, cnt = present(*g)
        out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        items.append((txt, qoff, lens, out_off, int(lens.sum()), 33, ch, cnt, g[2]))
        wants.append(np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(ooff, lens)]))
    back = eng.oq_demux(items)
    for b, w in zip(back, wants):
        assert np.array_equal(b, w), "GPU OQ demux mismatch"


@pytest.mark.gpu
def test_gpu_oq_errors(eng):
    from genozip_b200.lib import GzbError
    txt, qoff, lens, ooff = sam_like(50, 21)
    bad = txt.copy(); bad[int(qoff[3]) + 2] = 31                       # a QUAL character below '!'
    with pytest.raises(GzbError):
        eng.oq_mux([(bad, qoff, lens, ooff, None)])
    g = eng.oq_mux([(txt, qoff, lens, ooff, None)])[0]
    ch, cnt = present(*g)
    k = int(np.flatnonzero(cnt)[0]); cnt2 = cnt.copy(); cnt2[k] -= 1
    at = int(cnt[:k].sum())
    ch2 = np.concatenate([ch[:at + int(cnt2[k])], ch[at + int(cnt[k]):]])
    out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    with pytest.raises(GzbError):                                       # a channel out of data (:152)
        eng.oq_demux([(txt, qoff, lens, out_off, int(lens.sum()), 33, ch2, cnt2, g[2])])
