"""Builds ``lib/librnnt_b200.so`` for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "librnnt_b200.so")
SOURCES = ["capi.cu", "joint.cu", "persist.cu", "lattice.cu", "decode.cu"]
HEADERS = ["ptx.cuh", "common.cuh", "launch.h", os.path.join("..", "..", "include", "rnnt_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    prof = ["-DRNNT_PROFILE"] if os.environ.get("RNNT_PROFILE") == "1" else []
    prof += os.environ.get("RNNT_EXTRA_NVCC_FLAGS", "").split()     # bring-up experiments, e.g. -DRNNT_EXP_STAGES=6
    out = os.environ.get("RNNT_LIB_OUT", OUT)
    cmd = [nvcc] + NVCC_FLAGS + prof + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES + ["-lcudart"]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librnnt_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
