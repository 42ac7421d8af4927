"""Builds libgzb200.so of several branches side by side, for an A/B run inside ONE gpurun call:

    python tools/ab_build.py main wip/arith-enc-split wip/domq-line-kernels wip/arith-dec-run4
      -> ab_libs/<branch with / replaced by _>/libgzb200.so      (git-ignored, travels with the gpurun snapshot)
    /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/ab_bench.sh'
      -> gpurun_out/ab_<name>.{json,log}

Only for branches that change the CUDA library alone: the Python driver of the checked-out tree is used for all of them
(GZB200_LIB selects the library build, genozip_b200/lib.py).  Each branch is built in a temporary git worktree."""
import os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(branches):
    out_root = os.path.join(ROOT, "ab_libs")
    for b in branches:
        # "branch+FLAG": the branch built with an extra nvcc flag, e.g. main+-DGZB_READS_DONE_SYNCWARP (the lock-step markers as real __syncwarp())
        b, _, flag = b.partition("+")
        name = b.replace("/", "_") + (("_" + flag.lstrip("-D").lower()) if flag else "")
        env = dict(os.environ, GZB_NVCC_FLAGS=flag) if flag else dict(os.environ)
        wt = tempfile.mkdtemp(prefix="gzb_ab_")
        try:
            subprocess.run(["git", "-C", ROOT, "worktree", "add", "--detach", "--force", wt, b], check=True, capture_output=True)
            subprocess.run([sys.executable, os.path.join(wt, "genozip_b200", "build.py")], check=True, capture_output=True, env=env)
            os.makedirs(os.path.join(out_root, name), exist_ok=True)
            shutil.copy(os.path.join(wt, "genozip_b200", "libgzb200.so"), os.path.join(out_root, name, "libgzb200.so"))
            rev = subprocess.run(["git", "-C", wt, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
            open(os.path.join(out_root, name, "REV"), "w").write(f"{b} {rev} {flag}\n")
            print(f"{b} ({rev}) {flag} -> ab_libs/{name}/libgzb200.so")
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], capture_output=True)
            shutil.rmtree(wt, ignore_errors=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["main", "main+-DGZB_READS_DONE_SYNCWARP", "wip/arith-enc-split", "wip/domq-line-kernels", "wip/arith-dec-run4"])
