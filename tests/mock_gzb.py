"""A stand-in for libgzb200.so on a machine WITHOUT a GPU — test infrastructure only.

It exposes the entry points genozip_b200/fastq_path.py calls and computes them with the CPU checkers (tests/orc.py), reading
and writing the caller's buffers through the very descriptor arrays the real library would receive.  With it the host
driver (descriptor set-up, stream bookkeeping, per-pipeline threads, buffer sizing, byte accounting) runs end to end in
the CPU test suite; the CUDA library itself is exercised by the -m gpu tests.  `est_size` / `packed_len` come from the real
library (they need no device).

`install()` goes one step further: it puts a DryEngine — the REAL genozip_b200.lib.Engine marshalling code on top of MockLib —
in place of the package's Engine, so that `pytest -m gpu --dry-gpu` runs the GPU parity tests' own logic (descriptor set-up,
expectations, comparisons with the reference objects) on a machine without a GPU.  That checks the TESTS and the binding,
never the kernels; tests/test_gpu_tests_dry_run.py does it inside the CPU suite."""
import ctypes as C
import numpy as np

import orc
from genozip_b200.lib import load, CODEC, GZB_DEVICE_PTRS, GZB_OUT_DEVICE, GZB_IN_DEVICE
from genozip_b200.lib import Section as lib_Section

NAME = {v: k for k, v in CODEC.items()}


def _view(ptr, n, dtype=np.uint8):
    """numpy view of n items at address ptr (the 'device' memory of the mock is host memory)"""
    if not n:
        return np.zeros(0, dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * nbytes).from_address(int(ptr)), dtype=dtype)


class MockLib:
    def __init__(self):
        self.real = load()
        self.calls = []
        self.err = ""

    # ---- no device needed
    def gzb_acgt_packed_len(self, n):
        return self.real.gzb_acgt_packed_len(n)

    def gzb_est_size(self, codec, n):
        return self.real.gzb_est_size(codec, n)

    def gzb_engine_stream(self, h):
        return None

    def gzb_last_kernel_ms(self, h, which):
        return 0.0

    def gzb_last_error(self, h):
        return b"mock: " + self.err.encode()

    def gzb_kernel_launches(self, h):
        return len(self.calls)

    def gzb_last_chain_ms(self, h):
        return 0.0

    def gzb_engine_sync(self, h):
        return 0

    def gzb_engine_destroy(self, h):
        return 0

    def gzb_acgt_pack(self, h, seq, n, packed, x, allz, flags):
        pk, xs, z = orc.acgt_pack(_view(seq, n))
        _view(packed, pk.size)[:] = pk
        _view(x, n)[:] = xs
        allz._obj.value = int(z)
        return 0

    def gzb_acgt_unpack(self, h, packed, x, n, out, flags):
        plen = int(self.gzb_acgt_packed_len(n))
        _view(out, n)[:] = orc.acgt_unpack(_view(packed, plen).copy(), None if x is None else _view(x, n).copy(), n)
        return 0

    # ---- PBWT, LONGR
    def gzb_pbwt_encode(self, h, ht, n_lines, w, runs, runs_cap, nr, fgrc, fgrc_cap, nf, flags):
        r, f = orc.pbwt_encode(_view(ht, n_lines * w).reshape(n_lines, w))
        assert r.size <= runs_cap and f.size <= fgrc_cap
        _view(runs, r.size, np.uint32)[:] = r; _view(fgrc, f.size, np.uint32)[:] = f
        nr._obj.value, nf._obj.value = r.size, f.size
        return 0

    def gzb_pbwt_decode(self, h, runs, n_runs, fgrc, n_fgrc, n_lines, ht, size, hl, flags):
        _view(ht, size)[:] = orc.pbwt_decode(_view(runs, n_runs, np.uint32), _view(fgrc, n_fgrc, np.uint32), n_lines, size)
        hl._obj.value = size
        return 0

    def gzb_pbwt_encode_batch(self, h, arr, n, flags):
        rc = 0
        for i in range(n):
            a = arr[i]
            r, f = orc.pbwt_encode(_view(a.ht, a.n_lines * a.ht_per_line).reshape(a.n_lines, a.ht_per_line))
            if r.size > a.runs_cap or f.size > a.fgrc_cap:
                a.status = -3; rc = rc or -3; self.err = "PBWT: RUNS/FGRC capacity too small"
                continue
            _view(a.runs, r.size, np.uint32)[:] = r; _view(a.fgrc, f.size, np.uint32)[:] = f
            a.n_runs, a.n_fgrc, a.status = r.size, f.size, 0
        return rc

    def gzb_pbwt_decode_batch(self, h, arr, n, flags):
        for i in range(n):
            a = arr[i]
            runs, fgrc = _view(a.runs, a.n_runs, np.uint32), _view(a.fgrc, a.n_fgrc, np.uint32)
            size = int(fgrc[-2]) | (int(fgrc[-1]) << 32)
            if int(runs.sum(dtype=np.uint64)) < size or size > a.ht_cap:
                a.status = -4; self.err = "PBWT: runs do not cover the matrix"
                return -4
            _view(a.ht, size)[:] = orc.pbwt_decode(runs, fgrc, a.n_lines, size)
            a.ht_len, a.status = size, 0
        return 0

    def _longr(self, a):
        n = a.n_lines
        ln = _view(a.len, n, np.uint32)
        return (_view(a.txt, a.txt_len), _view(a.seq_off, n, np.uint64), _view(a.qual_off, n, np.uint64), ln,
                _view(a.is_rev, n) if a.is_rev else None, np.frombuffer(bytes(a.value_to_bin), np.uint8).copy(), int(ln.sum()))

    def gzb_longr_encode(self, h, arr, n, flags):
        for i in range(n):
            a = arr[i]
            txt, so, qo, ln, rv, v2b, tot = self._longr(a)
            ql = _view(a.qual_len, a.n_lines, np.uint32) if a.qual_len else None
            vals, lb = orc.longr_encode(txt, so, qo, ln, rv, v2b) if ql is None else orc.longr_encode(txt, so, qo, ql, rv, v2b, seq_lens=ln)
            _view(a.values, vals.size)[:] = vals; _view(a.lens_be, 65536, np.uint32)[:] = lb
        return 0

    def gzb_longr_decode(self, h, arr, n, flags):
        """line by line, like codec_longr_reconstruct: a line whose first value is 255 has no quality and takes that one value"""
        for i in range(n):
            a = arr[i]
            txt, so, qo, ln, rv, v2b, tot = self._longr(a)
            lb = _view(a.lens_be, 65536, np.uint32)
            nvals = int(lb.byteswap().sum(dtype=np.uint64))
            if nvals != (a.n_bases or tot):
                self.err = "LONGR: channel lengths do not add up to the number of qualities"
                return -4
            out, miss = orc.longr_decode_lines(txt, so, ln, rv, v2b, _view(a.values, nvals), lb)
            _view(a.qual_out, tot)[:] = out
            if a.missing:
                _view(a.missing, a.n_lines)[:] = miss
        return 0

    def gzb_longr_calculate_bins(self, h, arr, flags, v2b):
        a = arr[0]
        txt, so, qo, ln, rv, _, tot = self._longr(a)
        ql = _view(a.qual_len, a.n_lines, np.uint32) if a.qual_len else ln
        parts = [txt[int(o): int(o) + int(l)] for o, l in zip(qo, ql) if l and not (l == 1 and txt[int(o)] == 32)]
        if not parts:
            return 1
        _view(v2b, 256)[:] = orc.longr_bins(np.concatenate(parts))
        return 0

    # ---- ACGT
    def gzb_acgt_pack_batch(self, h, arr, n, flags):
        self.calls.append(("acgt_pack_batch", n, flags))
        for i in range(n):
            a = arr[i]
            pk, x, allz = orc.acgt_pack(_view(a.seq, a.n_bases))
            _view(a.packed, pk.size)[:] = pk
            if a.x:
                _view(a.x, a.n_bases)[:] = x
            a.x_all_zero = int(allz)
        return 0

    def gzb_acgt_unpack_batch(self, h, arr, n, flags):
        self.calls.append(("acgt_unpack_batch", n, flags))
        for i in range(n):
            a = arr[i]
            plen = int(self.gzb_acgt_packed_len(a.n_bases))
            x = _view(a.x, a.n_bases).copy() if a.x else None
            _view(a.seq, a.n_bases)[:] = orc.acgt_unpack(_view(a.packed, plen).copy(), x, a.n_bases)
        return 0

    # ---- DOMQ
    def _domq(self, a):
        off = _view(a.line_off, a.n_lines, np.uint64); ln = _view(a.line_len, a.n_lines, np.uint32)
        return orc.domq_encode(_view(a.txt, a.txt_len), off, ln)

    def gzb_domq_prepare(self, h, arr, n, flags):
        self.calls.append(("domq_prepare", n, flags))
        self._enc = []
        for i in range(n):
            a = arr[i]
            e = self._domq(a)
            self._enc.append(e)
            a.num_norm_qs, a.num_doms, a.has_diverse = e["num_norm_qs"], e["num_doms"], e["has_diverse"]
            C.memmove(a.denorm, e["denorm"].ctypes.data, e["denorm"].size)
            C.memmove(a.normalize, e["normalize"].ctypes.data, min(e["normalize"].size, 95 * 95))
            _view(a.line_dom, a.n_lines)[:] = e["line_dom"]; _view(a.line_diverse, a.n_lines)[:] = e["line_diverse"]
        return 0

    def gzb_domq_split(self, h, arr, n, flags):
        self.calls.append(("domq_split", n, flags))
        for i in range(n):
            a, e = arr[i], self._enc[i]
            for fld, k in (("qual", "qual"), ("runs", "runs"), ("mplx", "mplx"), ("divr", "divr")):
                assert e[k].size <= getattr(a, fld + "_cap")
                _view(getattr(a, fld), e[k].size)[:] = e[k]
                setattr(a, fld + "_len", e[k].size)
        return 0

    def gzb_domq_reconstruct(self, h, arr, n, flags):
        self.calls.append(("domq_reconstruct", n, flags))
        for i in range(n):
            a = arr[i]
            enc = dict(qual=_view(a.qual, a.qual_len).copy(), runs=_view(a.runs, a.runs_len).copy(), mplx=_view(a.mplx, a.mplx_len).copy(),
                       divr=_view(a.divr, a.divr_len).copy(), denorm=_view(a.denorm, a.denorm_len).copy(), num_norm_qs=a.num_norm_qs)
            ln = _view(a.line_len, a.n_lines, np.uint32)
            out = orc.domq_decode(enc, ln)
            assert out.size <= a.out_cap
            _view(a.out, out.size)[:] = out
        return 0

    # ---- simple codecs
    def gzb_compress_sections(self, h, secs, n, flags):
        self.calls.append(("compress", n, flags))
        for i in range(n):
            s = secs[i]
            name = NAME[s.codec]
            data = _view(s.in_, s.in_len).copy()
            c = orc.compress("port", "rans" if name.startswith("RAN") else "arith", data, orc.ORDER[name])
            if c.size > s.out_cap:                                          # soft fail (compressor.c:90)
                s.out_len = 0; s.status = 1
                continue
            _view(s.out, c.size)[:] = c
            s.out_len = c.size; s.status = 0
        return 0

    def gzb_compress_sections_packed(self, h, secs, n, arena, cap, used, flags):
        self.calls.append(("compress_packed", n, flags))
        secs = C.cast(secs, C.POINTER(lib_Section)) if not hasattr(secs, "__getitem__") else secs
        outs = []
        for i in range(n):
            s = secs[i]
            name = NAME[s.codec]
            outs.append(orc.compress("port", "rans" if name.startswith("RAN") else "arith", _view(s.in_, s.in_len).copy(), orc.ORDER[name]))
        total = sum((c.size + 15) & ~15 for c in outs)
        if used is not None:
            used._obj.value = total
        if total > cap:
            for i in range(n):
                secs[i].status = 1; secs[i].out_len = 0
            return 1
        off = 0
        for i, c in enumerate(outs):
            _view(int(arena) + off, c.size)[:] = c
            secs[i].out = int(arena) + off; secs[i].out_len = c.size; secs[i].status = 0
            off += (c.size + 15) & ~15
        return 0

    def gzb_copy_batch(self, h, cps, n):
        for i in range(n):
            c = cps[i]
            if c.len:
                _view(c.dst, c.len)[:] = _view(c.src, c.len)
        return 0

    def gzb_local_transform_batch(self, h, items, n, flags):
        names = {v: k for k, v in orc.LT_OPS.items()}
        for i in range(n):
            it = items[i]
            w = {1: 2, 2: 4, 3: 8, 4: 1, 5: 2, 6: 4, 7: 8, 8: 1, 9: 2, 10: 4, 11: 8}[it.op]
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w]
            if it.n_elems:
                v = _view(it.data, it.n_elems, dt)
                v[:] = orc.local_transform(names[it.op], v.copy()).view(dt)
            it.status = 0
        return 0

    # ---- the steps either side of the codecs, and the round-2 quality codecs
    def gzb_b250_generate_batch(self, h, items, n, flags):
        rc = 0
        for i in range(n):
            a = items[i]
            r = orc.b250_generate(_view(a.b250, a.len).copy(), _view(a.ni2wi, a.n_new, np.int32).copy(), a.ol_len, bool(a.one_up_ok))
            if r is None:
                a.status = -4; rc = -4; self.err = "b250: the backward scan does not end at the start of the buffer"; continue
            out, nw = r
            if out.size:
                _view(a.out, a.len)[a.len - out.size:] = out
            a.out_len = out.size; a.n_words = nw; a.status = 0
        return rc

    def gzb_local_transpose_batch(self, h, items, n, flags):
        for i in range(n):
            a = items[i]
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[a.width]
            a.transposed = 0; a.status = 0
            if not a.n_elems:
                continue
            if a.n_elems % a.cols:
                if a.dir == 1:
                    a.status = -4; self.err = "transposed local is not a rectangle"; return -4
                continue
            out, tr = orc.local_transpose(_view(a.data, a.n_elems, dt).copy(), a.cols, piz=a.dir == 1)
            _view(a.data, a.n_elems, dt)[:] = out
            a.transposed = 1
        return 0

    def gzb_oq_mux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            sl = _view(a.seq_len, nl, np.uint32).copy() if a.seq_len else None
            r = orc.oq_mux(_view(a.txt, a.txt_len).copy(), _view(a.qual_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(),
                           _view(a.oq_off, nl, np.uint64).copy(), sl)
            if r is None:
                a.status = -4; self.err = "OQ: a QUAL character outside '!'..'~'"; return -4
            ch, cnt, mono = r
            if ch.size:
                _view(a.channels, ch.size)[:] = ch
            for q in range(94):
                a.count[q] = int(cnt[q]); a.monochars[q] = int(mono[q])
            a.status = 0
        return 0

    def gzb_oq_demux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            cnt = np.array(a.count[:], np.uint32); mono = np.array(a.monochars[:], np.uint8)
            out = orc.oq_demux(_view(a.txt, a.txt_len).copy(), _view(a.qual_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(),
                               _view(a.out_off, nl, np.uint64).copy(), a.out_cap, a.key_bias, _view(a.channels, int(cnt.sum())).copy(), cnt, mono)
            if out is None:
                a.status = -4; self.err = "OQ: a channel is out of data"; return -4
            if out.size:
                _view(a.out, out.size)[:] = out
            a.status = 0
        return 0

    def gzb_pacb_mux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            np0 = _view(a.np0, nl).copy() if a.np0 and nl else None
            ch, cnt = orc.pacb_mux(_view(a.txt, a.txt_len).copy(), _view(a.qual_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(),
                                   _view(a.seq_off, nl, np.uint64).copy(), np0, a.max_np)
            if ch.size:
                _view(a.channels, ch.size)[:] = ch
            for c in range(84):
                a.count[c] = int(cnt[c])
            a.status = 0
        return 0

    def gzb_pacb_demux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            np0 = _view(a.np0, nl).copy() if a.np0 and nl else None
            cnt = np.array(a.count[:], np.uint32)
            ch = _view(a.channels, int(cnt.sum())).copy()
            if (ch == 32).any():
                a.status = -5; self.err = "SMUX / PACB: a read without quality"; return -5
            out = orc.pacb_demux(_view(a.txt, a.txt_len).copy(), _view(a.seq_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(), np0, a.max_np,
                                 _view(a.out_off, nl, np.uint64).copy(), a.out_cap, ch, cnt)
            if out is None:
                a.status = -4; self.err = "OQ / SMUX: a channel is out of data"; return -4
            if out.size:
                _view(a.out, out.size)[:] = out
            a.status = 0
        return 0

    def gzb_tmpl_mux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            ch, cnt = orc.tmpl_mux(_view(a.txt, a.txt_len).copy(), _view(a.qual_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(), _view(a.tmpl, a.tmpl_len).copy())
            if ch.size:
                _view(a.channels, ch.size)[:] = ch
            for q in range(95):
                a.count[q] = int(cnt[q])
            a.status = 0
        return 0

    def gzb_tmpl_demux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            cnt = np.array(a.count[:], np.uint32)
            out = orc.tmpl_demux(_view(a.qual_len, nl, np.uint32).copy(), _view(a.out_off, nl, np.uint64).copy(), a.out_cap, _view(a.tmpl, a.tmpl_len).copy(),
                                 _view(a.channels, int(cnt.sum())).copy(), cnt)
            if out is None:
                a.status = -4; self.err = "OQ / SMUX: a channel is out of data"; return -4
            if out.size:
                _view(a.out, out.size)[:] = out
            a.status = 0
        return 0

    def gzb_smux_mux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            rev = _view(a.is_rev, nl).copy() if a.is_rev and nl else None
            ch, cnt, par = orc.smux_mux(_view(a.txt, a.txt_len).copy(), _view(a.qual_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(),
                                        _view(a.seq_off, nl, np.uint64).copy(), _view(a.seq_len, nl, np.uint32).copy(), rev)
            if ch.size:
                _view(a.channels, ch.size)[:] = ch
            for b in range(5):
                a.count[b] = int(cnt[b])
            a.n_param = par; a.status = 0
        return 0

    def gzb_smux_demux(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            rev = _view(a.is_rev, nl).copy() if a.is_rev and nl else None
            cnt = np.array(a.count[:], np.uint32)
            ch = _view(a.channels, int(cnt.sum())).copy()
            if (ch == 32).any():
                a.status = -5; self.err = "SMUX: a read without quality"; return -5
            out = orc.smux_demux(_view(a.txt, a.txt_len).copy(), _view(a.seq_off, nl, np.uint64).copy(), _view(a.qual_len, nl, np.uint32).copy(), rev,
                                 _view(a.out_off, nl, np.uint64).copy(), a.out_cap, ch, cnt, a.n_param)
            if out is None:
                a.status = -4; self.err = "OQ / SMUX: a channel is out of data"; return -4
            if out.size:
                _view(a.out, out.size)[:] = out
            a.status = 0
        return 0

    def gzb_homp_condense(self, h, vbs, n, mode, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            out, newlen = orc.hp_condense(mode, _view(a.txt, a.txt_len).copy(), _view(a.str_off, nl, np.uint64).copy(), _view(a.str_len, nl, np.uint32).copy(),
                                          _view(a.seq_off, nl, np.uint64).copy())
            if out.size:
                _view(a.local, out.size)[:] = out
            if a.new_len and nl:
                _view(a.new_len, nl, np.uint32)[:] = newlen
            a.local_len = out.size; a.status = 0
        return 0

    def gzb_homp_expand(self, h, vbs, n, mode, flags):
        for i in range(n):
            a = vbs[i]
            nl = a.n_lines
            r = orc.hp_expand(mode, _view(a.local, a.local_len).copy(), _view(a.txt, a.txt_len).copy(), _view(a.seq_off, nl, np.uint64).copy(), _view(a.str_len, nl, np.uint32).copy())
            if r is None:
                a.status = -4; self.err = "HOMP / T0: the stream does not match the lines"; return -4
            out, miss = r
            if out.size:
                _view(a.out, out.size)[:] = out
            if a.missing and nl:
                _view(a.missing, nl)[:] = miss
            a.status = 0
        return 0

    def gzb_normq_gather(self, h, vbs, n, flags):
        for i in range(n):
            a = vbs[i]
            ln = _view(a.line_len, a.n_lines, np.uint32).copy() if a.n_lines else np.zeros(0, np.uint32)
            off = _view(a.line_off, a.n_lines, np.uint64).copy() if a.n_lines else np.zeros(0, np.uint64)
            rev = _view(a.is_rev, a.n_lines).copy() if a.is_rev and a.n_lines else None
            out = orc.normq_encode(_view(a.txt, a.txt_len) if a.txt_len else np.zeros(1, np.uint8), off, ln, rev)
            if out.size:
                _view(a.local, out.size)[:] = out
            a.local_len = out.size; a.status = 0
        return 0

    def gzb_normq_reconstruct(self, h, vbs, n, flags):
        rc = 0
        for i in range(n):
            a = vbs[i]
            ln = _view(a.line_len, a.n_lines, np.uint32).copy() if a.n_lines else np.zeros(0, np.uint32)
            rev = _view(a.is_rev, a.n_lines).copy() if a.is_rev and a.n_lines else None
            r = orc.normq_decode(_view(a.local, a.local_len).copy() if a.local_len else np.zeros(0, np.uint8), ln, rev)
            if r is None:
                a.status = -4; rc = -4; self.err = "NORMQ: the stream does not match the lines"; continue
            out, miss = r
            if out.size:
                _view(a.out, out.size)[:] = out
            if a.missing and miss.size:
                _view(a.missing, miss.size)[:] = miss
            a.status = 0
        return rc

    def gzb_stage_upload(self, h, dst, src, n):
        if n:
            _view(dst, n)[:] = _view(src, n)
        return 0

    gzb_stage_fetch = gzb_stage_upload

    def gzb_stage_wait(self, h, which):
        return 0

    def gzb_assign_codecs(self, h, items, n, flags):
        names = ("RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw")
        ids = (6, 7, 8, 9, 16, 17, 18, 19)
        for i in range(n):
            it = items[i]
            it.sample_len = min(it.len, 99999); it.best = 0
            if it.sample_len < 50:
                continue
            d = _view(it.data, it.sample_len).copy()
            best, best_size = 1, it.sample_len
            for k, nm in enumerate(names):
                it.size[k] = orc.compress("port", "rans" if nm.startswith("RAN") else "arith", d, orc.ORDER[nm]).size
                if it.size[k] + 28 < best_size:
                    best, best_size = ids[k], it.size[k] + 28
            it.best = best
        return 0

    def gzb_adler32_batch(self, h, items, n, flags):
        import zlib
        for i in range(n):
            it = items[i]
            it.adler = zlib.adler32(_view(it.data, it.len).tobytes() if it.len else b"", 1) & 0xffffffff
        return 0

    def gzb_uncompress_sections(self, h, secs, n, flags):
        self.calls.append(("uncompress", n, flags))
        for i in range(n):
            s = secs[i]
            name = NAME[s.codec]
            try:
                d = orc.uncompress("port", "rans" if name.startswith("RAN") else "arith", _view(s.in_, s.in_len).copy(), s.out_cap)
            except AssertionError as e:
                self.err = f"section {i}: {e}"
                s.status = -1
                return -1
            _view(s.out, d.size)[:] = d
            s.out_len = d.size; s.status = 0
        return 0


class MockEngine:
    """Engine look-alike on top of MockLib (what FastqCodecPath needs of genozip_b200.lib.Engine)"""
    torch_device = "cpu"
    _lib = None

    def __init__(self, device=0):
        if MockEngine._lib is None:
            MockEngine._lib = MockLib()
        self.L, self.h, self.device, self.launches = MockEngine._lib, None, device, 0

    def close(self):
        pass

    def _err(self):
        return "mock"

    def compress_raw(self, secs, n, flags=0):
        assert self.L.gzb_compress_sections(self.h, secs, n, flags) == 0

    def uncompress_raw(self, secs, n, flags=0):
        assert self.L.gzb_uncompress_sections(self.h, secs, n, flags) == 0

    def compress(self, items):
        return [orc.compress("port", "rans" if c.startswith("RAN") else "arith", np.ascontiguousarray(d, np.uint8), orc.ORDER[c]) for c, d in items]

    def assign_codecs_ptrs(self, ptr_len, flags):
        from genozip_b200.lib import Engine
        return Engine.assign_codecs_ptrs(self, ptr_len, flags)      # the real marshalling over MockLib.gzb_assign_codecs


def install():
    """put the real Engine marshalling code on top of MockLib in place of genozip_b200.Engine (see the module docstring)"""
    import genozip_b200, genozip_b200.lib as lib

    class DryEngine(lib.Engine):
        def __init__(self, device=0):
            if MockEngine._lib is None:
                MockEngine._lib = MockLib()
            self.L, self.h, self.device = MockEngine._lib, None, device

        def close(self):
            pass

    genozip_b200.Engine = lib.Engine = DryEngine
    return DryEngine
