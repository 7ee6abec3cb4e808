"""nvcc build of ``librightlane.so`` for sm_100a, in-tree (the .so travels to the GPU box with the repo)."""
import glob
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "librightlane.so")
BUILD = os.path.join(_HERE, "build")

INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
NVCC_FLAGS = ["-I", INCLUDE, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(INCLUDE, "*.h")))
    os.makedirs(BUILD, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s}\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
        else:
            for line in out.splitlines():
                if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                    sys.stderr.write(f"[spill] {os.path.basename(s)}: {line.strip()}\n")
    if failed:
        raise RuntimeError("nvcc build failed")
    if procs or not os.path.exists(OUT):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
