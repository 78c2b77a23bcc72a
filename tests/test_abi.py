"""The C-ABI library loads and exports every symbol include/bde_b200.h declares (no compute)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "bde_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bde_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from beyond_deep_ensembles_b200 import build_ext
    return build_ext.build(verbose=False)


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("bde_svgd_pairdist", "bde_svgd_bandwidth", "bde_svgd_apply", "bde_swag_update", "bde_swag_sample",
                 "bde_ivon_sample", "bde_ivon_update", "bde_gauss_sample_fwd", "bde_gauss_sample_bwd",
                 "bde_kl_gauss_value_and_grad", "bde_l2_value_and_grad", "bde_svgd_step_host", "bde_swag_sample_batch",
                 "bde_svgd_apply_sgd", "bde_svgd_apply_adam", "bde_svgd_train_step_sgd", "bde_svgd_train_step_adam",
                 "bde_prior_terms_value_and_grad", "bde_multi_tensor_copy", "bde_peer_attach"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in bde_b200.h but not exported"
    lib.bde_version.restype = ctypes.c_int
    assert lib.bde_version() == 100
    lib.bde_error_string.restype = ctypes.c_char_p
    assert lib.bde_error_string(-3) == b"bde: workspace missing or too small"


def test_tuning_keys_documented_in_the_header_are_accepted(built_lib):
    """bde_tune is host-only code: every key the header's comment names is accepted (value 0 = automatic choice),
    anything else and negative values are BDE_ERR_INVALID_ARG, and no key is left without a word in the header."""
    lib = ctypes.CDLL(str(built_lib))
    lib.bde_tune.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.bde_tune.restype = ctypes.c_int
    comment = HEADER.read_text().split("int bde_tune(")[0].rsplit("/*", 1)[1]
    documented = set(re.findall(r'"([a-z_0-9]+)"', comment))
    assert {"pairdist_ctas_per_sm", "apply_variant", "swag_batch", "batch_prefetch", "batch_splits"} <= documented
    for key in documented:
        assert lib.bde_tune(key.encode(), 0) == 0, key
    assert lib.bde_tune(b"no_such_knob", 0) != 0
    assert lib.bde_tune(b"swag_batch", -1) != 0
    src = (ROOT / "beyond_deep_ensembles_b200" / "csrc" / "util.cu").read_text()
    implemented = set(re.findall(r'k == "([a-z_0-9]+)"', src))
    internal = {"gram_pairing", "gram_fold", "gram_l2_promotion", "ring_kb", "gram_guard_x1000"}   # csrc/common.cuh:Tuning only
    assert implemented - internal == documented


def test_python_binding_covers_the_header(built_lib):
    from beyond_deep_ensembles_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"bde_error_string"}
    assert bound == set(declared_symbols())


def test_fake_abi_mirrors_the_header():
    from fake_abi import FakeLib
    for name in declared_symbols():
        assert callable(getattr(FakeLib, name, None)), f"test double lacks {name}"


def test_sass_uses_packed_fp32(built_lib):
    """The SVGD kernels must compile to Blackwell's packed FADD2/FFMA2 (not scalar FFMA)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN3bde20svgd_pairdist_kernelILi10EEEvPKfllPdiPviNS_15BandwidthParamsEi",
                           str(built_lib)], capture_output=True, text=True).stdout
    assert sass.count("FFMA2") >= 90 and sass.count("FADD2") >= 90, "pairdist<10> lost its packed-fp32 inner loop"
    assert "LDG.E.128" in sass or "LDG.E.ENL2.128" in sass or ".128" in sass
    # centred-Gram K1 (n = 20): 190 entries x 2 FFMA2 per column quad, fed by one tensor-map TMA load per tile
    nm = subprocess.run(["nm", "-D", str(built_lib)], capture_output=True, text=True).stdout
    gram = [ln.split()[-1] for ln in nm.splitlines() if "svgd_pairgram_kernelILi20ELb1" in ln]
    assert gram, "pairgram<20> is not in the library"
    sass = subprocess.run([cuobjdump, "-sass", "-fun", gram[0], str(built_lib)], capture_output=True, text=True).stdout
    assert sass.count("FFMA2") >= 380 and "UTMALDG.2D" in sass, "pairgram<20> lost its packed inner loop / TMA tensor load"


def test_no_fallback_when_library_missing(monkeypatch, tmp_path):
    from beyond_deep_ensembles_b200 import _lib
    monkeypatch.setattr(_lib, "_handle", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.BdeError):
        _lib.get()


def test_cpu_tensors_are_rejected(built_lib):
    import torch
    from beyond_deep_ensembles_b200 import _lib, ops
    with pytest.raises(_lib.BdeError):
        ops.require_cuda(torch.zeros(4))
