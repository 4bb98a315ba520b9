"""Build libsdrd_b200.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc."""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsdrd_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def sources():
    return [os.path.join(CSRC, "sdrd_capi.cu")]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sdrd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile into a temporary file and rename it over libsdrd_b200.so under a file lock, so that several
    processes (ranks) starting with a stale library neither race on the output nor load a half-written one."""
    if not force and not _stale():
        return OUT
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():  # another process built it while this one waited
            return OUT
        fd, tmp = tempfile.mkstemp(prefix=".libsdrd_b200.", suffix=".so.tmp", dir=HERE)
        os.close(fd)
        try:
            cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            os.chmod(tmp, 0o755)
            os.replace(tmp, OUT)
        finally:
            if os.path.exists(tmp):
                os.unlink(tmp)
    if verbose:
        print(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
