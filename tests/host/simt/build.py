"""Builds tests/host/_build/libgzb200_simt.so: the product's .cu sources (genozip_b200/csrc), UNCHANGED except for the three
mechanical rewrites below, compiled by g++ against tests/host/simt/cuda_runtime.h and linked with the SIMT emulator — the
same C-ABI as libgzb200.so, every kernel executed on the host, lane by lane.  TEST INFRASTRUCTURE ONLY (see cuda_runtime.h).

Rewrites (done on a temporary copy, nothing is committed):
  kern<<<grid, block, smem, stream>>>(args)            ->  SIMT_LAUNCH (kern, (grid), (block), smem, stream, args)
  extern __shared__ ... T name[];                      ->  T *name = (T *)simt::dyn_smem ();
  asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)) ->  r = simt::rcp_approx (x)
  __noinline__                                          ->  __attribute__((noinline))
"""
import glob, os, re, shutil, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, "genozip_b200", "csrc")
OUTDIR = os.path.join(os.path.dirname(HERE), "_build")
OUT = os.path.join(OUTDIR, "libgzb200_simt.so")
CXX = os.environ.get("CXX", "g++")
FLAGS = ["-std=c++17", "-O1", "-g1", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-DGZB_SIMT_EMULATION=1", "-x", "c++"]


def _split_args(s):
    """split at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _matching(s, i, open_ch="(", close_ch=")"):
    """index just past the bracket that closes the one at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced")


def rewrite(src, allow_asm=False):
    # kernel launches
    out, pos = "", 0
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9]*(?:\s*<[^<>;(){}]*>)?)\s*<<<", src):
        if m.start() < pos:
            continue
        cfg_end = src.index(">>>", m.end())
        cfg = _split_args(src[m.end():cfg_end])
        while len(cfg) < 4:
            cfg.append("0")
        k = cfg_end + 3
        while src[k].isspace():
            k += 1
        assert src[k] == "(", src[m.start():k + 20]
        end = _matching(src, k)
        args = src[k + 1:end - 1].strip()
        out += src[pos:m.start()] + f"SIMT_LAUNCH (({m.group(1)}), ({cfg[0]}), ({cfg[1]}), {cfg[2]}, {cfg[3]}" + (", " + args if args else "") + ")"
        pos = end
    src = out + src[pos:]
    # dynamic shared memory
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\s*\(\s*\d+\s*\)\s*)?([A-Za-z_0-9]+)\s+([A-Za-z_0-9]+)\s*\[\s*\]\s*;",
                 r"\1 *\2 = (\1 *)simt::dyn_smem ();", src)
    # (libstdc++ spells an attribute __noinline__, so this one cannot be a macro)
    src = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", src)
    # the one PTX instruction
    src = re.sub(r'asm\s*\(\s*"rcp\.approx\.ftz\.f32 %0, %1;"\s*:\s*"=f"\s*\((\w+)\)\s*:\s*"f"\s*\((.*?)\)\s*\)\s*;', r"\1 = simt::rcp_approx (\2);", src)
    # (gzb_tma.cuh keeps its PTX — bulk copies, mbarriers — behind #ifndef GZB_SIMT_EMULATION and brings its own host stand-ins)
    assert "<<<" not in src and (allow_asm or "asm" not in re.sub(r"//.*", "", src).replace("rcp_approx", "")), "an untranslated launch or asm statement is left"
    return src


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(HERE, "*")) + [os.path.join(ROOT, "include", "gzb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    gen = os.path.join(OUTDIR, "gen", "genozip_b200", "csrc")              # (the sources include "../../include/gzb200.h")
    shutil.rmtree(os.path.join(OUTDIR, "gen"), ignore_errors=True)          # (no stale sources of another checkout)
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(OUTDIR, "gen", "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "gzb200.h"), os.path.join(OUTDIR, "gen", "include", "gzb200.h"))
    for f in glob.glob(os.path.join(CSRC, "*")):
        txt = open(f).read()
        open(os.path.join(gen, os.path.basename(f)), "w").write(rewrite(txt, allow_asm=os.path.basename(f) == "gzb_tma.cuh") if f.endswith((".cu", ".cuh")) else txt)
    objs, procs = [], []
    for f in sorted(glob.glob(os.path.join(gen, "*.cu"))) + [os.path.join(HERE, "simt.cpp")]:
        o = os.path.join(OUTDIR, os.path.basename(f) + ".o")
        objs.append(o)
        cmd = [CXX] + FLAGS + ["-I", HERE, "-I", gen, "-I", os.path.join(ROOT, "include"), "-c", f, "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for f, p in procs:
        log = p.communicate()[0]
        if p.returncode:
            bad = True
            sys.stderr.write(f"--- {f}\n{log[:6000]}\n")
        elif verbose and log:
            print(log)
    if bad:
        raise RuntimeError("g++ failed on the SIMT build")
    r = subprocess.run([CXX, "-shared", "-o", OUT] + objs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libgzb200_simt.so failed")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
