"""Build libp3d.so (sm_100a only) in-tree with nvcc.

    python 3d-pose-baseline_b200/build.py [--force]

Cross-compiles without a GPU.  The shared library lands next to the Python host layer
(p3d/libp3d.so) so that it travels with the tree; ptxas resource usage goes to build/ptxas.log.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "p3d", "libp3d.so")
SOURCES = ["api.cu", "mlp_prep.cu", "mlp_tc.cu", "tc_gemm.cu", "mlp_layered.cu", "mlp_simt.cu", "mlp_mid.cu", "geometry.cu", "procrustes.cu", "train.cu", "p2p.cu", "realtime.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(os.path.dirname(HERE), "include", "p3d.h"))
    return hdr


def _compile(src):
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    deps = _deps()
    todo = [s for s in SOURCES
            if force or _newer([os.path.join(CSRC, s)] + deps, os.path.join(BUILD, os.path.splitext(s)[0] + ".o"))]
    logs = []
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, obj, rc, log in ex.map(_compile, todo):
                logs.append(f"==== {src}\n{log}")
                if rc != 0:
                    sys.stderr.write(log)
                    raise RuntimeError(f"nvcc failed on {src}")
        with open(os.path.join(BUILD, "ptxas.log"), "a" if not force else "w") as f:
            f.write("\n".join(logs))
    objs = [os.path.join(BUILD, os.path.splitext(s)[0] + ".o") for s in SOURCES]
    if todo or _newer(objs, OUT):
        cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print(f"built {OUT} ({os.path.getsize(OUT)} bytes; recompiled: {', '.join(todo) or 'nothing'})")
    elif verbose:
        print(f"{OUT} is up to date")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
