#!/usr/bin/env python
"""One pass of every entry point either side of the codec path on a batch of realistic size, for an ncu launch list:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_oq|k_hp|k_b250|k_local|k_normq|k_adler" \
        --csv --log-file gpurun_out/r02_aux_kernels.csv python tools/aux_kernels.py --vblocks 32

Host buffers (the call times printed include PCIe); the kernel times are ncu's.  Prints the algorithmic bytes of every kernel family."""
import argparse, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vblocks", type=int, default=32)
    ap.add_argument("--reads", type=int, default=92000)
    ap.add_argument("--read-len", type=int, default=150)
    a = ap.parse_args()
    from genozip_b200 import Engine
    eng = Engine(0)
    rng = np.random.default_rng(1)
    V, R, L = a.vblocks, a.reads, a.read_len
    n = R * L
    res = {}

    def timed(name, fn, nbytes):
        t0 = time.time(); out = fn(); dt = time.time() - t0
        res[name] = {"call_ms": round(dt * 1e3, 1), "algorithmic_bytes": int(nbytes)}
        return out

    # one VBlock's text: SEQ | QUAL | OQ, line tables into it
    vbs = []
    for v in range(V):
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
        qual = rng.choice(np.frombuffer(b"#,:F", np.uint8), n, p=[.02, .06, .12, .8])
        oq = (qual - (rng.random(n) < 0.2)).astype(np.uint8)
        txt = np.concatenate([seq, qual, oq])
        off = (np.arange(R, dtype=np.uint64) * L)
        ln = np.full(R, L, np.uint32)
        vbs.append((txt, off, off + np.uint64(n), off + np.uint64(2 * n), ln))
    rev = (np.arange(R) % 3 == 0).astype(np.uint8)
    loc = timed("normq_gather", lambda: eng.normq_gather([(t, qo, ln, rev) for t, so, qo, oo, ln in vbs]), 2 * n * V)
    timed("normq_reconstruct", lambda: eng.normq_reconstruct([(l, vbs[0][4], rev) for l in loc]), 2 * n * V)
    mux = timed("oq_mux", lambda: eng.oq_mux([(t, qo, ln, oo, None) for t, so, qo, oo, ln in vbs]), 3 * n * V)
    oo_out = (np.arange(R, dtype=np.uint64) * L)
    def present(ch, cnt, mono):
        c2 = np.where(mono != 0, 0, cnt).astype(np.uint32); at = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        keep = [ch[at[q]:at[q + 1]] for q in range(94) if c2[q]]
        return (np.concatenate(keep) if keep else np.zeros(0, np.uint8)), c2
    timed("oq_demux", lambda: eng.oq_demux([(t, qo, ln, oo_out, n, 33) + present(*m) + (m[2],) for (t, so, qo, oo, ln), m in zip(vbs, mux)]), 3 * n * V)
    sm = timed("smux_mux", lambda: eng.smux_mux([(t, qo, ln, so, ln, None) for t, so, qo, oo, ln in vbs]), 3 * n * V)
    timed("smux_demux", lambda: eng.smux_demux([(t, so, ln, None, oo_out, n, m[0], m[1], m[2]) for (t, so, qo, oo, ln), m in zip(vbs, sm)]), 3 * n * V)
    tmpl = np.full(L, ord("F"), np.uint8)
    tm = timed("tmpl_mux", lambda: eng.tmpl_mux([(t, qo, ln, tmpl) for t, so, qo, oo, ln in vbs]), 2 * n * V)
    timed("tmpl_demux", lambda: eng.tmpl_demux([(ln, oo_out, n, tmpl, m[0], m[1]) for (t, so, qo, oo, ln), m in zip(vbs, tm)]), 2 * n * V)
    pc = timed("pacb_mux", lambda: eng.pacb_mux([(t, qo, ln, so, None, 1) for t, so, qo, oo, ln in vbs]), 3 * n * V)
    timed("pacb_demux", lambda: eng.pacb_demux([(t, so, ln, None, 1, oo_out, n, m[0], m[1]) for (t, so, qo, oo, ln), m in zip(vbs, pc)]), 3 * n * V)
    hp = timed("homp_condense", lambda: eng.hp_condense(0, [(t, qo, ln, so) for t, so, qo, oo, ln in vbs]), 3 * n * V)
    timed("homp_expand", lambda: eng.hp_expand(0, [(h[0], t, so, ln) for (t, so, qo, oo, ln), h in zip(vbs, hp)]), 3 * n * V)
    # b250: R words per context, 8 contexts per VBlock
    items = []
    for v in range(V * 8):
        wi = rng.integers(0, 3000, R)
        lens = np.where(wi <= 126, 1, 2); pos = np.concatenate([[0], np.cumsum(lens)[:-1]])
        b = np.zeros(int(lens.sum()), np.uint8)
        one = wi <= 126
        b[pos[one]] = wi[one]
        b[pos[~one]] = (wi[~one] - 127) & 0xff; b[pos[~one] + 1] = 0x80 | ((wi[~one] - 127) >> 8)
        items.append((b, np.zeros(0, np.int32), 4000, True))
    timed("b250_generate", lambda: eng.b250_generate(items), sum(2 * i[0].size for i in items))
    mats = [(rng.integers(0, 1 << 32, R * 16, dtype=np.uint32), 16) for _ in range(V)]
    timed("local_transpose", lambda: eng.local_transpose(mats), sum(2 * m.nbytes for m, _ in mats))
    timed("adler32", lambda: eng.adler32([t for t, *_ in vbs]), 3 * n * V)
    timed("assign_codecs", lambda: eng.assign_codecs([t[n:2 * n] for t, *_ in vbs]), 99999 * 8 * V)
    print(json.dumps({"vblocks": V, "reads": R, "read_len": L, "entries": res}))


if __name__ == "__main__":
    main()
