"""ctypes binding of the C-ABI library (include/bde_b200.h).

There is exactly one implementation behind this module: lib/libbde_b200.so built for sm_100a.
If it is missing, or a kernel returns an error, this module raises — there is no CPU or
PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

# BDE_B200_LIB points at another BUILD of the same library (same-box A/B runs of two kernel versions); default in-tree
LIB_PATH = Path(os.environ.get("BDE_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libbde_b200.so")

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_d = C.c_double
_sz = C.c_size_t

# name -> argtypes, mirroring include/bde_b200.h declaration by declaration
SIGNATURES = {
    "bde_version": [],
    "bde_device_sm_count": [_p],
    "bde_tune": [C.c_char_p, _i],
    "bde_svgd_workspace_bytes": [_i, C.POINTER(_sz)],
    "bde_value_workspace_bytes": [C.POINTER(_sz)],
    "bde_peer_buffer_bytes": [C.POINTER(_sz)],
    "bde_peer_alloc": [C.POINTER(_p), C.c_char_p],
    "bde_peer_open": [C.c_char_p, C.POINTER(_p)],
    "bde_peer_close": [_p],
    "bde_peer_free": [_p],
    "bde_peer_attach": [_p, _sz, _i, _i, C.POINTER(_u64), _d, _p, _p],
    "bde_peer_detach": [_p, _sz, _p],
    "bde_peer_status": [_p, C.POINTER(_u64), C.POINTER(_u64)],
    "bde_peer_wait_stats": [_p, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), _i],
    "bde_svgd_pairdist": [_p, _i, _i64, _i64, _p, _i, _p, _sz, _p],
    "bde_svgd_bandwidth": [_p, _i, _d, _d, _d, _d, _p, _p, _p, _p, _p],
    "bde_svgd_apply": [_p, _p, _p, _p, _p, _i, _i64, _i64, _p],
    "bde_svgd_apply_sgd": [_p, _p, _p, _p, _i, _i64, _i64, _p, _i, _d, _d, _d, _d, _i, _p, _p],
    "bde_svgd_apply_adam": [_p, _p, _p, _p, _i, _i64, _i64, _p, _p, _i64, _d, _d, _d, _d, _d, _i, _p, _p],
    "bde_svgd_train_step_sgd": [_p, _p, _p, _p, _i, _i64, _i64, _p, _i, _d, _d, _d, _d, _i, _p,
                                _p, _i, _d, _d, _d, _d, _p, _p, _p, _p, _p, _sz, _p],
    "bde_svgd_train_step_adam": [_p, _p, _p, _p, _i, _i64, _i64, _p, _p, _i64, _d, _d, _d, _d, _d, _i, _p,
                                 _p, _i, _d, _d, _d, _d, _p, _p, _p, _p, _p, _sz, _p],
    "bde_svgd_pairdist_bandwidth": [_p, _i, _i64, _i64, _d, _d, _d, _d, _p, _p, _p, _p, _p, _p, _sz, _p],
    "bde_svgd_host_pairdist": [_p, _i, _i64, _i64, _i64, _p, _p, _p, _sz],
    "bde_svgd_host_apply": [_p, _p, _i, _i64, _i64, _d, _d, _d, _d, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "bde_svgd_chain_next": [_p],
    "bde_svgd_step": [_p, _p, _p, _i, _i64, _i64, _d, _d, _d, _d, _p, _p, _p, _p, _p, _p, _sz, _p],
    "bde_svgd_step_host": [_p, _p, _p, _i, _i64, _i64, _d, _d, _d, _d, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz,
                           _p, _p],
    "bde_swag_update": [_p, _p, _p, _p, _i64, _i64, _p],
    "bde_swag_sample": [_p, _p, _p, _i, _i, _i64, _i64, _p, _p, _u64, _u64, _i64, _p, _p],
    "bde_swag_sample_batch": [_p, _p, _p, _i, _i, _i64, _i64, _i, _p, _p, _i64, _u64, _u64, _i64, _p, _i64, _p],
    "bde_ivon_sample": [_p, _p, _p, _p, _i64, _d, _i, _i, _p, _u64, _u64, _i64, _p],
    "bde_ivon_sample_batch": [_p, _p, _p, _p, _i64, _i64, _i, _d, _i, _i, _p, _i64, _u64, _u64, _u64, _i64, _p],
    "bde_ivon_accumulate": [_p, _p, _i64, _i, _p],
    "bde_ivon_update": [_p, _p, _p, _p, _p, _i64, _i, _i64, _d, _d, _d, _d, _d, _d, _d, _p],
    "bde_gauss_sample_fwd": [_p, _p, _p, _i64, _p, _u64, _u64, _i64, _p],
    "bde_gauss_sample_bwd": [_p, _p, _p, _i64, _p, _u64, _u64, _i64, _p],
    "bde_kl_gauss_value_and_grad": [_p, _p, _i64, _d, _d, _p, _p, _p, _d, _p, _i, _p, _sz, _p],
    "bde_kl_mixture_value_and_grad": [_p, _i64, _d, _d, _d, _p, _p, _d, _p, _i, _p, _sz, _p],
    "bde_l2_value_and_grad": [_p, _i64, _d, _p, _p, _d, _p, _i, _p, _sz, _p],
    "bde_prior_terms_value_and_grad": [_i, _p, _p, _p, _p, _p, _p, _p, _d, _d, _d, _p, _d, _p, _i, _p, _sz, _p],
    "bde_bbb_linear_workspace_bytes": [_i, _i, _i, C.POINTER(_sz)],
    "bde_bbb_linear_fwd": [_p, _i64, _i, _i, _i, _p, _p, _p, _p, _p, _u64, _u64, _d, _p, _p, _p, _p, _sz, _p],
    "bde_rank1_linear_fwd": [_p, _i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _u64, _u64, _u64, _p, _p, _p, _p, _p, _p, _p, _sz, _p],
    "bde_philox_normal": [_p, _i64, _u64, _u64, _i64, _p],
    "bde_multi_tensor_copy": [_p, _p, _p, _p, _i, _i, _p],
    "bde_multi_tensor_unscale_copy": [_p, _p, _p, _p, _i, _i, _p, _p, _p],
}

_handle = None
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


class BdeError(RuntimeError):
    pass


def get():
    """Load (once) and return the ctypes handle.  Raises if the library is not built."""
    global _handle
    if _handle is None:
        if not LIB_PATH.exists():
            raise BdeError(
                f"{LIB_PATH} is missing: build it with `python -m beyond_deep_ensembles_b200.build_ext` "
                "(there is no fallback implementation)")
        h = C.CDLL(str(LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        h.bde_error_string.argtypes = [C.c_int]
        h.bde_error_string.restype = C.c_char_p
        _handle = h
    return _handle


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = get().bde_error_string(rc)
        raise BdeError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")


def ptr(t: torch.Tensor | None) -> int | None:
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_call_device = None  # device of the stream handed out last: call() launches there (see call)


def stream_ptr(device: torch.device | None = None) -> int:
    """Current stream of `device` as a raw handle; also notes the device so that call() can make it current."""
    global _call_device
    if device is not None and device.type == "cuda":
        _call_device = device
        return torch.cuda.current_stream(device).cuda_stream
    return 0


def require_f32(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise TypeError(f"expected float32 tensor, got {t.dtype}")


def call(name: str, *args) -> None:
    """Invoke one C-ABI entry point and raise on a non-zero return code."""
    global launch_count, _call_device
    fn = getattr(get(), name)
    dev, _call_device = _call_device, None
    # The library launches on the calling thread's CURRENT device (and queries its SM count / occupancy), while the
    # stream argument was taken from the tensors' device: make the two agree (model on cuda:1, current device cuda:0).
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = fn(*args)
    else:
        rc = fn(*args)
    launch_count += 1
    check(rc, name)
