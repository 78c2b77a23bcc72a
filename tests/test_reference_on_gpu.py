"""install() against the reference's OWN experiment factories on a CUDA device:
`experiments.uci.models.get_model` (MLP, D = 501) and `experiments.cifar.models.get_model` (ResNet-20-FRN, D = 273,610,
96 tensors, n = 20 particles) build their ensembles twice — with the unmodified reference classes (eager PyTorch on the
same GPU) and, after `bde.install()`, with this package's classes on the C-ABI library — and two optimizer steps of
every algorithm must agree on identical data and identical injected noise (tests/ref_live.py, one subprocess per
case because install() rebinds names inside the reference's modules).

The reference comes from oracle/_ref (staged by oracle/install_ref.py, travels with the snapshot); /root/reference
is never read here.  Tolerance: fp32 rtol 1e-5 / atol 1e-6 for one update, 3e-5 / 3e-6 where steps compound."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest
import torch

from oracle import install_ref

ROOT = Path(__file__).resolve().parent.parent

CASES = [("uci", a) for a in ("svgd", "swag", "ivon", "bbb", "rank1")] + \
        [("cifar", a) for a in ("svgd", "swag", "ivon", "bbb", "rank1")] + \
        [("civilcomments", a) for a in ("bbb", "rank1", "ivon")]     # DistilBERT (random init) + BBB / Rank-1 head, full iVON


def _staged():
    if not install_ref.available():
        install_ref.install(verbose=False)   # build container: stage it from /root/reference
    return install_ref.available()


def _live(task, algo, device):
    import os
    res = subprocess.run([sys.executable, str(ROOT / "tests" / "ref_live.py"), task, algo], capture_output=True, text=True,
                         timeout=1500, env={**os.environ, "REF_LIVE_DEVICE": device})
    assert res.returncode == 0 and "REF_LIVE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-6000:]


@pytest.mark.gpu
@pytest.mark.parametrize("task,algo", CASES)
def test_installed_path_matches_the_reference_on_cuda(task, algo):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not _staged():   # __graft_entry__.build() stages it; a box that never saw the reference cannot run the comparison
        pytest.skip("oracle/_ref is missing: run `python oracle/install_ref.py` (or __graft_entry__.build()) where /root/reference exists")
    _live(task, algo, "cuda:0")


@pytest.mark.parametrize("algo", ["svgd", "ivon", "rank1"])
def test_installed_host_logic_matches_the_reference_on_cpu(algo):
    """The same comparison on the host: live reference on CPU against this package's classes over the oracle-backed
    double of the C-ABI (tests/fake_abi.py) — pins the host logic (closure protocol, noise order, particle
    aliasing) against the reference itself, not only against recorded fixtures."""
    if not _staged():
        pytest.skip("reference neither staged under oracle/_ref nor present at /root/reference")
    _live("uci", algo, "cpu")


def test_staged_reference_is_intact():
    if not _staged():
        pytest.skip("reference neither staged under oracle/_ref nor present at /root/reference")
    """Every staged file still has the bytes recorded when it was copied from the reference checkout."""
    assert install_ref.verify() >= 30
