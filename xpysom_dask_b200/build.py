"""Build libsom_b200.so in-tree with nvcc for sm_100a.

    python -m xpysom_dask_b200.build [--force] [--experiments]

--experiments compiles the tuning knobs in (-DSOM_B200_EXPERIMENTS: SOM_B200_* environment variables, the
in-kernel timeline probe); the production library never reads the environment.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsom_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(HERE, "..", "include", "som_b200.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=True, experiments=False):
    if not force and not experiments and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + (["-DSOM_B200_EXPERIMENTS"] if experiments else []) + ["-o", OUT, os.path.join(CSRC, "som_api.cu")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, experiments="--experiments" in sys.argv)
    print("built", OUT)
