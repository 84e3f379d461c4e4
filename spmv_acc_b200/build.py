"""In-tree nvcc build of the CUDA libraries (sm_100a only).

The shared objects are written to ``spmv_acc_b200/lib/`` so that they travel with the repository snapshot to the
GPU box; nothing is JIT-compiled at run time.

    python -m spmv_acc_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
INCLUDE = ROOT / "include"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-shared", f"-I{INCLUDE}", f"-I{CSRC}"]

TARGETS = {
    # library -> (sources, extra link flags)
    "libspmv_b200.so": (["analysis.cu", "kernels.cu", "convert.cu", "capi.cu"], []),
    "libspmv_b200_gen.so": (["gen.cu"], []),
    "libspmv_b200_ctx.so": (["context_baselines.cu"], ["-lcusparse"]),
}


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA libraries of spmv_acc_b200 cannot be built")


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build_target(name: str, force: bool = False, verbose: bool = False) -> Path:
    sources, link = TARGETS[name]
    srcs = [CSRC / s for s in sources]
    missing = [s for s in srcs if not s.exists()]
    if missing:
        raise FileNotFoundError(f"missing CUDA sources for {name}: {missing}")
    out = LIB / name
    deps = srcs + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    if not force and not _stale(out, deps):
        return out
    LIB.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc_path(), *ARCH, *COMMON]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(out), *map(str, srcs), *link]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


def build_all(force: bool = False, verbose: bool = False) -> list[Path]:
    return [build_target(n, force=force, verbose=verbose) for n in TARGETS if all((CSRC / s).exists() for s in TARGETS[n][0])]


if __name__ == "__main__":
    outs = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    for o in outs:
        print(o)
