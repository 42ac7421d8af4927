"""TMPL (src/codec_tmpl.c), Element's quality codec: QUAL multiplexed by a per-position template.  CPU: the restatement against the
reference's compiled codec_tmpl.c (oracle/_ref) — which finds the template itself (codec_tmpl_segconf_finalize) and then compresses.
GPU (-m gpu, also --simt): gzb_tmpl_mux / gzb_tmpl_demux against the restatement and the reference; demux as the inverse."""
import numpy as np
import pytest

import orc


def element_like(n_lines, seed, tmpl_len=150, longer=True):
    """qualities that follow a per-position profile (a plateau, a decay at the end) with noise; some reads shorter, a few longer than the template"""
    rng = np.random.default_rng(seed)
    profile = np.array([ord("5") if i < tmpl_len * 0.13 else ord("+") if i >= tmpl_len * 0.93 else ord("?") if i >= tmpl_len * 0.73 else ord("I") for i in range(tmpl_len)], np.uint8)
    parts, qoff, lens = [np.frombuffer(b"@x\n", np.uint8)], [], []
    pos = parts[0].size
    for _ in range(n_lines):
        L = tmpl_len if rng.random() < 0.8 else int(rng.integers(1, tmpl_len + (25 if longer else 0)))
        q = np.resize(profile, L).copy() if L > tmpl_len else profile[:L].copy()
        if L > tmpl_len:
            q[tmpl_len:] = rng.integers(40, 60, L - tmpl_len)
        noise = rng.random(L) < 0.2
        q[noise] = rng.integers(35, 75, int(noise.sum()))
        qoff.append(pos); parts.append(q); pos += L; lens.append(L)
        filler = rng.integers(65, 91, int(rng.integers(0, 9))).astype(np.uint8); parts.append(filler); pos += filler.size
    return np.concatenate(parts), np.array(qoff, np.uint64), np.array(lens, np.uint32)


@pytest.mark.parametrize("seed", [1, 2])
def test_port_matches_reference(seed):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    txt, qoff, lens = element_like(500, seed, longer=False)          # (std_seq_len is the longest read of the segconf data)
    r = orc.ref_tmpl_mux(txt, qoff, lens, 150)
    assert r is not None, "the reference found no dominant template"
    tmpl, rch, rcnt = r
    assert (tmpl[25:80] == ord("I")).all()                            # the profile's plateau
    pch, pcnt = orc.tmpl_mux(txt, qoff, lens, tmpl)
    assert np.array_equal(pcnt, rcnt) and np.array_equal(pch, rch)


def test_port_round_trip_with_excess():
    txt, qoff, lens = element_like(300, 5)
    tmpl = np.full(150, ord("I"), np.uint8); tmpl[:20] = ord("5"); tmpl[140:] = ord("+")
    ch, cnt = orc.tmpl_mux(txt, qoff, lens, tmpl)
    assert cnt[94] == int(np.maximum(lens.astype(np.int64) - 150, 0).sum()) and cnt[94] > 0
    out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    back = orc.tmpl_demux(lens, out_off, int(lens.sum()), tmpl, ch, cnt)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qoff, lens)])
    assert back is not None and np.array_equal(back, want)
    cnt2 = cnt.copy(); k = int(np.flatnonzero(cnt)[0]); cnt2[k] -= 1
    at = int(cnt[:k].sum()); ch2 = np.concatenate([ch[:at + int(cnt2[k])], ch[at + int(cnt[k]):]])
    assert orc.tmpl_demux(lens, out_off, int(lens.sum()), tmpl, ch2, cnt2) is None


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_tmpl(eng):
    cases = []
    for seed, tl in ((11, 150), (12, 100), (13, 33)):
        txt, qoff, lens = element_like(400, seed, tmpl_len=tl)
        tmpl = orc.ref_tmpl_mux(*element_like(400, seed, tmpl_len=tl, longer=False), tl)[0] if orc.have_gz_ref() else np.full(tl, ord("I"), np.uint8)
        cases.append((txt, qoff, lens, tmpl))
    got = eng.tmpl_mux(cases)
    for c, g in zip(cases, got):
        w = orc.tmpl_mux(*c)
        assert np.array_equal(g[1], w[1]) and np.array_equal(g[0], w[0]), "GPU != restatement"
    if orc.have_gz_ref():                                              # and against the reference's own compress, on reads no longer than the template
        txt, qoff, lens = element_like(500, 1, longer=False)
        tmpl, rch, rcnt = orc.ref_tmpl_mux(txt, qoff, lens, 150)
        g = eng.tmpl_mux([(txt, qoff, lens, tmpl)])[0]
        assert np.array_equal(g[1], rcnt) and np.array_equal(g[0], rch), "GPU != reference codec_tmpl.c"
    items, wants = [], []
    for (txt, qoff, lens, tmpl), g in zip(cases, got):
        out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        items.append((lens, out_off, int(lens.sum()), tmpl, g[0], g[1]))
        wants.append(np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qoff, lens)]))
    for b, w in zip(eng.tmpl_demux(items), wants):
        assert np.array_equal(b, w), "GPU TMPL demux mismatch"
