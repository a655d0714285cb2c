"""Builds libvapb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m voiceactivityprojection_b200.build [--force]
    python -m voiceactivityprojection_b200.build --variant b -D SOME_KNOB=1     # libvapb_b.so for VAPB_LIB A/B runs

nvcc cross-compiles without a GPU. Objects are cached under csrc/build/ keyed on
the source mtime; the .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvapb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant / defines: a second build for same-box A/B runs (`VAPB_LIB=.../libvapb_<variant>.so`): every source is
    compiled with the extra -D flags into csrc/build_<variant>/ and linked as libvapb_<variant>.so."""
    if variant:
        return _build(force, verbose, os.path.join(CSRC, "build_" + variant),
                      os.path.join(HERE, f"libvapb_{variant}.so"), [f"-D{d}" for d in defines])
    return _build(force, verbose, os.path.join(CSRC, "build"), OUT, [])


def _build(force: bool, verbose: bool, bdir: str, out: str, extra) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
                  + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC, *FLAGS, *extra, "-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for s, r in ex.map(compile_one, jobs):
            log = os.path.join(bdir, os.path.basename(s) + ".log")
            with open(log, "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {s}")
            if verbose:
                sys.stderr.write(r.stderr)
    if jobs or force or _stale(out, objs):
        r = subprocess.run([NVCC, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--variant", default="", help="build libvapb_<variant>.so next to the main library")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="extra preprocessor define (with --variant)")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.v, variant=a.variant, defines=a.defines))
