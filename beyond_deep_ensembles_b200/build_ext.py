"""Build the C-ABI shared library libbde_b200.so for sm_100a with nvcc (in-tree).

`python -m beyond_deep_ensembles_b200.build_ext` compiles every .cu under csrc/ in parallel
(-gencode arch=compute_100a,code=sm_100a -lineinfo) and links them into
beyond_deep_ensembles_b200/lib/libbde_b200.so.  Objects are cached by a hash of the source,
the headers and the flags, so rebuilding after an edit only recompiles what changed.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libbde_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fno-gnu-unique",   # template statics stay private to this .so (two builds can share a process)
    "-Xptxas", "-v",
    f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built")
    return exe


def _digest(src: Path, headers: list[Path]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [src, *headers]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def _compile(src: Path, headers: list[Path], verbose: bool) -> Path:
    obj = OBJDIR / f"{src.stem}.{_digest(src, headers)}.o"
    if obj.exists():
        return obj
    for old in OBJDIR.glob(f"{src.stem}.*.o"):
        old.unlink()
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (OBJDIR / f"{src.stem}.ptxas.log").write_text(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr[-4000:]}")
    if verbose:
        print(f"[build_ext] compiled {src.name}")
    return obj


def build(verbose: bool = True, force: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    if force:
        for old in OBJDIR.glob("*.o"):
            old.unlink()
        (LIBDIR / "link.stamp").unlink(missing_ok=True)   # fresh objects are linked again even if their names did not change
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, headers, verbose), sources))
    stamp = LIBDIR / "link.stamp"
    want = " ".join(o.name for o in objs)
    if LIB.exists() and stamp.exists() and stamp.read_text() == want:
        return LIB
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stderr[-4000:]}")
    stamp.write_text(want)
    if verbose:
        print(f"[build_ext] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
