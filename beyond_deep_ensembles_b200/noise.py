"""Noise source of the sampling kernels.

Production: Philox4x32-10 inside the kernels, keyed by (seed, stream_id) and counted by the
global element index; `seed` follows torch.initial_seed() (so torch.manual_seed controls it)
and every sampling call takes a fresh stream_id.  All ranks of a D-sharded job must be seeded
identically.

Parity tests: the reference draws its noise at three points (src.algos.util.normal_like,
ivorn.normal_like, LowRankMultivariateNormal._standard_normal); `inject(fn)` installs a callable
fn(kind, numel) -> flat fp32 tensor | None that supplies the noise for those same draws:
kind is "ivon", "swag_k", "swag_d" or "gauss".
"""
from __future__ import annotations

import contextlib
from typing import Callable, Optional

import torch

_injector: Optional[Callable] = None
_seed: Optional[int] = None
_next_stream = 0


def seed() -> int:
    return torch.initial_seed() & 0xFFFFFFFFFFFFFFFF if _seed is None else _seed


def set_seed(value: Optional[int]) -> None:
    """Pin the Philox key (None = follow torch.initial_seed()) and restart the stream counter."""
    global _seed, _next_stream
    _seed = None if value is None else int(value) & 0xFFFFFFFFFFFFFFFF
    _next_stream = 0


def next_stream_id() -> int:
    global _next_stream
    _next_stream += 1
    return _next_stream


def reserve_stream_ids(count: int) -> int:
    """`count` consecutive stream ids (what `count` calls of next_stream_id() would return); returns the first."""
    global _next_stream
    first = _next_stream + 1
    _next_stream += int(count)
    return first


def stream_position() -> int:
    """How many Philox stream ids this process has handed out (checkpointed by the optimizers' state_dict)."""
    return _next_stream


def restore_stream_position(position) -> None:
    """Resume after a checkpoint: never hand out a stream id again that the saved run had already used (a resumed
    run with the same seed would otherwise replay the same noise)."""
    global _next_stream
    if position is not None:
        _next_stream = max(_next_stream, int(position))


def draw(kind: str, numel: int, device) -> Optional[torch.Tensor]:
    """Injected noise for this draw, or None to let the kernel use Philox."""
    if _injector is None:
        return None
    z = _injector(kind, numel)
    if z is None:
        return None
    z = torch.as_tensor(z, dtype=torch.float32).reshape(-1)
    if z.numel() != numel:
        raise ValueError(f"injected noise for {kind!r} has {z.numel()} elements, expected {numel}")
    return z.to(device).contiguous()


@contextlib.contextmanager
def inject(fn: Callable):
    global _injector
    prev = _injector
    _injector = fn
    try:
        yield
    finally:
        _injector = prev
