"""Build the C-ABI CUDA library in-tree: ``python -m probpose_pytorch_b200.build``.

nvcc cross-compiles for sm_100a without a GPU.  The resulting
``csrc/libprobpose_b200.so`` has no dependency on torch or Python (cudart is
linked statically) and is git-ignored but travels to the GPU box with the tree.

The library carries a hash of the sources it was compiled from
(``pp_source_hash()``); ``_lib.lib()`` compares it with the tree on every import,
so an edited ``.cu`` / ``.cuh`` can never run against a stale binary.  Builds are
serialised with a file lock (several ranks of a ``torchrun`` may import at once).
"""

from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
ROOT = CSRC.parents[1]
LIB = CSRC / "libprobpose_b200.so"
SOURCES = ["pp_api.cu", "pp_encode.cu", "pp_decode.cu", "pp_loss.cu", "pp_targets.cu", "pp_sparsemax.cu", "pp_metrics.cu", "pp_records.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: cannot build libprobpose_b200.so")
    return cand


def _deps(experiments: bool = False) -> list[Path]:
    deps = [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "probpose_b200.h"]
    if experiments:
        deps += sorted((ROOT / "tools" / "experiments").glob("*.cuh"))
    return deps


def source_hash(experiments: bool = False) -> str:
    """sha256 over the CUDA sources, the header and the compiler flags (16 hex digits)."""
    h = hashlib.sha256()
    for d in _deps(experiments):
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16] + ("+x" if experiments else "")


def built_hash() -> str | None:
    """The source hash embedded in the existing library, or None (no library / no tag).  Read from the file's bytes:
    loading the library here would pin that copy in the process (dlopen returns an already loaded library of the same
    name even after the file has been replaced)."""
    if not LIB.exists():
        return None
    data = LIB.read_bytes()
    i = data.find(b"pp_source_hash=")
    if i < 0:
        return None
    j = data.find(b";", i)
    return data[i + 15:j].decode(errors="replace") if 0 < j - i < 64 else None


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link ``libprobpose_b200.so`` (no-op when the
    existing library was built from the current sources)."""
    want = source_hash(experiments)
    if not force and built_hash() == want:
        return LIB
    nvcc = _nvcc()
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    with open(objdir / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and built_hash() == want:   # another process built it while we waited
            return LIB
        flags = list(NVCC_FLAGS) + (["-DPP_EXPERIMENTS"] if experiments else [])
        procs = []
        for src in SOURCES:
            obj = objdir / (Path(src).stem + ".o")
            cmd = [nvcc, *flags, "-c", str(CSRC / src), "-o", str(obj)]
            if src == "pp_api.cu":
                cmd.insert(1, f'-DPP_SOURCE_HASH="{want}"')
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs = []
        for src, obj, pr in procs:
            out, _ = pr.communicate()
            if verbose or pr.returncode:
                sys.stderr.write(out)
            if pr.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(str(obj))
        tmp = LIB.with_suffix(".so.tmp")
        link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                "-Xcompiler", "-fPIC", "-o", str(tmp), *objs]
        res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout)
            raise RuntimeError("link of libprobpose_b200.so failed")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
