"""In-tree build of libbfcnn_b200.so (nvcc, sm_100a only).

`python -m blind_image_denoising_b200.build` or `__graft_entry__.build()`.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libbfcnn_b200.so"
SOURCES = ["host_pack.cu", "api.cu", "conv_f32.cu", "conv_x3.cu", "conv_t5.cu", "base_conv.cu", "base_conv_t5.cu", "fused_stream.cu", "fused_stream_x3.cu", "train.cu", "generic.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "bfcnn_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    logs = {}

    def compile_one(src: str):
        obj = objdir / (src.replace(".cu", ".o"))
        extra = os.environ.get("BFCNN_NVCC_EXTRA", "").split()   # e.g. -DBFCNN_STREAM_TRACE_BUILD (kernel timeline)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(OUT), *map(str, objs), "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (objdir / "ptxas.log").write_text("\n".join(f"==== {k}\n{v}" for k, v in logs.items()))
    if verbose:
        print((objdir / "ptxas.log").read_text())
    return OUT


if __name__ == "__main__":
    p = build_native(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
