"""Build libb200unet.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "lib"
LIB_PATH = LIB_DIR / "libb200unet.so"
SOURCES = ["api.cu", "conv_simt.cu", "conv_small.cu", "conv_tc.cu", "conv_gemm.cu", "wgrad_tc.cu", "norm.cu", "resample.cu", "loss.cu", "optim.cu", "pipeline.cu", "metrics.cu", "head_mid.cu", "peer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b200_unet.h"]
    jobs = []
    for src in SOURCES:
        obj = obj_dir / (src + ".o")
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            outs = list(ex.map(run, jobs))
        if verbose:
            print("\n".join(outs))
    objs = [str(obj_dir / (s + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        run([_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
