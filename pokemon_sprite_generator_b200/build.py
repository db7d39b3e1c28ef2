"""Build the C-ABI CUDA library in-tree: csrc/*.cu -> csrc/libpsg_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU container; the resulting .so travels to the
GPU box with the repo snapshot.  No JIT cache, no torch extension machinery: plain `nvcc -shared`.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_NAME = "libpsg_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]
# tuning builds only (e.g. PSG_EXTRA_NVCC_FLAGS=-DUATTN_PROF): part of the object digests, so a change of it recompiles
NVCC_FLAGS += os.environ.get("PSG_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (looked at $NVCC, PATH, /usr/local/cuda/bin)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def lib_path() -> Path:
    return CSRC / LIB_NAME


def build(force: bool = False, verbose: bool = True) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    obj_dir = CSRC / "build"
    obj_dir.mkdir(exist_ok=True)
    hdr_digest = _digest(headers)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = obj_dir / (src.stem + ".o")
        stamp = obj_dir / (src.stem + ".sha")
        dig = _digest([src]) + hdr_digest
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(dig)
        return obj, True

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    rebuilt = any(r for _, r in results)
    out = lib_path()
    if rebuilt or not out.exists() or force:
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", str(out), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[psg_b200] built {out} from {len(objs)} objects", file=sys.stderr)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
