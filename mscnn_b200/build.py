"""Build the in-tree native libraries with nvcc / g++ (no JIT cache: the built .so files travel
with the repo snapshot to the GPU box).

  libmscnn_b200.so  -- CUDA kernels + C ABI + Caffe-API mirror   (mscnn_b200/csrc/*.cu, *.cpp)

Usage:  python -m mscnn_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "mscnn_b200" / "csrc"
LIB = ROOT / "mscnn_b200" / "libmscnn_b200.so"
OBJDIR = ROOT / "build" / "obj"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-fvisibility=hidden,-fno-gnu-unique",
    "--expt-relaxed-constexpr",
    # decode / IoU arithmetic must round exactly like the reference's scalar CPU code:
    # no FMA contraction anywhere unless a kernel asks for it explicitly with __fmaf_rn.
    "-fmad=false",
    "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(CSRC / "caffe_api"), "-I", str(CSRC / "proto_shared"),
]


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _headers() -> list[Path]:
    return sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp"))
                  + list((CSRC / "caffe_api").glob("**/*.hpp")) + list((CSRC / "proto_shared").glob("**/*.h*"))
                  + list((ROOT / "include").glob("*.h")))


def _digest(paths: list[Path], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    hdr_digest = _digest(_headers(), " ".join(NVCC_FLAGS))
    objs = []
    rebuilt = False
    for src in _sources():
        tag = _digest([src], hdr_digest)
        obj = OBJDIR / f"{src.stem}.{tag}.o"
        objs.append(obj)
        if obj.exists() and not force:
            continue
        for old in OBJDIR.glob(f"{src.stem}.*.o"):
            old.unlink()
        cmd = [NVCC, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}")
        rebuilt = True
    if rebuilt or force or not LIB.exists():
        # static cudart (nvcc default): the library carries its own runtime and shares the
        # primary context with torch, so torch device pointers / streams are usable as-is.
        cmd = [NVCC, "-shared", "-Xlinker", "-Bsymbolic", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
