"""Stage the UNMODIFIED reference (its optimizer path only) under oracle/_ref/ so that it travels to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE — never imported by the product (beyond_deep_ensembles_b200/).  Users:
`bench.py --impl reference` (the reference's own `SVGDOptimizer.step` on the box's host cores), bench.py's
`cpu_baseline` / `eager_cuda` legs, `tests/test_reference_on_gpu.py` (install() against the reference's factories on
CUDA) and `tests/perf_whole_step.py`.

The reference is pure Python: "building" it means copying the files of its `BayesianOptimizer` path as they lie
under /root/reference — `src/algos/**`, `src/architectures/*.py` and the four `experiments/<task>/models.py`
factories — byte for byte, directory layout kept, into oracle/_ref/ (git-ignored like a built .so, NOT
gpurun-ignored).  Nothing else of the reference (training scripts, data loaders, evaluation, notebooks) is staged.
A MANIFEST.json with the sha256 of every staged file is written beside them; `verify()` re-checks it.

    python oracle/install_ref.py            # needs /root/reference (build container only)
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent
DEST = ROOT / "oracle" / "_ref"

PATTERNS = [
    "src/algos/*.py",
    "src/algos/kernel/*.py",
    "src/architectures/*.py",
    "experiments/__init__.py",
    "experiments/uci/__init__.py", "experiments/uci/models.py",
    "experiments/cifar/models.py",
    "experiments/iwildcam/__init__.py", "experiments/iwildcam/models.py",
    "experiments/civilcomments/__init__.py", "experiments/civilcomments/models.py",
    "LICENSE",
]


def _sha(path: Path) -> str:
    return hashlib.sha256(path.read_bytes()).hexdigest()


def install(verbose: bool = True) -> Path | None:
    """Copy the reference's optimizer path into oracle/_ref/.  Returns the destination, or None when
    /root/reference is absent (GPU box: the staged copy that travelled with the snapshot is used as is)."""
    if not REF.is_dir():
        return DEST if (DEST / "MANIFEST.json").exists() else None
    if DEST.exists():
        shutil.rmtree(DEST)
    manifest = {}
    for pat in PATTERNS:
        for src in sorted(REF.glob(pat)):
            rel = src.relative_to(REF)
            dst = DEST / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[str(rel)] = _sha(dst)
    (DEST / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "files": manifest}, indent=1, sort_keys=True))
    if verbose:
        print(f"[install_ref] staged {len(manifest)} reference files under {DEST}")
    return DEST


def verify() -> int:
    """Number of staged files whose bytes still match the manifest (raises if one differs or is missing)."""
    man = json.loads((DEST / "MANIFEST.json").read_text())["files"]
    for rel, digest in man.items():
        if _sha(DEST / rel) != digest:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the staged reference")
    return len(man)


def available() -> bool:
    return (DEST / "MANIFEST.json").exists()


def add_to_path() -> str:
    """Put oracle/_ref first on sys.path (so `import src.algos...` / `import experiments...` resolve to the staged
    reference) and return the path.  Raises if the reference was never staged."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/install_ref.py` in the build container")
    p = str(DEST)
    if p not in sys.path:
        sys.path.insert(0, p)
    return p


if __name__ == "__main__":
    d = install()
    if d is None:
        print("[install_ref] /root/reference not present and nothing staged", file=sys.stderr)
        sys.exit(1)
    print(f"[install_ref] verified {verify()} files")
