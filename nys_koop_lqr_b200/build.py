"""In-tree build of libnkb200.so (hand-written sm_100a CUDA + the C ABI of include/nk_b200.h).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import pathlib
import subprocess

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libnkb200.so"
SOURCES = ["nk_api.cu", "nk_gram.cu", "nk_dense.cu", "nk_rollout.cu", "nk_cv.cu", "nk_pgemm.cu", "nk_tgemm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + os.environ.get("NK_EXTRA_NVCC_FLAGS", "").split()


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "nk_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    stamp_file = HERE / ".libnkb200.stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    objs = []
    procs = []
    bdir = HERE / "build"
    bdir.mkdir(exist_ok=True)
    for s in srcs:
        o = bdir / (pathlib.Path(s).stem + ".o")
        objs.append(str(o))
        cmd = [NVCC, *FLAGS, "-c", s, "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", str(LIB), *objs, "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True)
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
