"""Property tests (hypothesis) of the host-side logic that every kernel call depends on: arena layouts, D-shard
bounds, the copy tables of the gradient gather, the MultiX sample split and the Philox stream-id bookkeeping.
CPU only; the C-ABI is the oracle-backed double of tests/fake_abi.py where a call is needed."""
from __future__ import annotations

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import fake_abi
from beyond_deep_ensembles_b200 import noise, ops
from beyond_deep_ensembles_b200.ensemble import split_samples
from beyond_deep_ensembles_b200.layout import ALIGN, ParamLayout, shard_bounds

shapes = st.lists(st.lists(st.integers(1, 7), min_size=1, max_size=3).map(tuple), min_size=1, max_size=9)


@given(D=st.integers(1, 5_000_000), world=st.integers(1, 9))
@settings(max_examples=200, deadline=None)
def test_shard_bounds_partition_the_columns(D, world):
    spans = [shard_bounds(D, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == D
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert a <= b == c <= d
    widths = [b - a for a, b in spans]
    assert sum(widths) == D
    # whole ALIGN-blocks are dealt out evenly: full shards differ by at most one block
    full = [w for w in widths[:-1] if w > 0]
    assert not full or max(full) - min(full) <= ALIGN
    assert all(a % ALIGN == 0 or a == D for a, _ in spans)


@given(shapes=shapes, rows=st.integers(1, 3))
@settings(max_examples=60, deadline=None)
def test_layout_views_alias_rows_and_logical_order_roundtrips(shapes, rows):
    params = [torch.randn(s) for s in shapes]
    L = ParamLayout(params)
    assert L.logical_size == sum(p.numel() for p in params)
    assert L.size % ALIGN == 0 and all(o % ALIGN == 0 for o in L.offsets)
    assert all(o1 >= o0 + n0 for o0, n0, o1 in zip(L.offsets, L.numels, L.offsets[1:]))
    arena = L.new_arena(rows, "cpu")
    for r in range(rows):
        for k, v in enumerate(L.views(arena[r])):
            assert v.shape == params[k].shape and v.data_ptr() == arena[r, L.offsets[k]:].data_ptr()
            v.copy_(params[k] + r)
    logical = L.to_logical(arena)                     # the reference's parameters_to_vector order, no padding
    want = torch.stack([torch.cat([(p + r).reshape(-1) for p in params]) for r in range(rows)])
    assert torch.equal(logical, want)
    back = L.from_logical(logical)
    assert torch.equal(back, arena)                   # padding stays zero


@given(shapes=shapes, mode=st.sampled_from([0, 1]), scale=st.sampled_from([None, 1.0, 512.0]))
@settings(max_examples=40, deadline=None)
def test_gather_then_scatter_roundtrip_through_the_copy_table(shapes, mode, scale):
    class MP:   # minimal monkeypatch stand-in: this test restores the attributes itself
        def __init__(self):
            self.undo = []

        def setattr(self, obj, name, val):
            self.undo.append((obj, name, getattr(obj, name)))
            setattr(obj, name, val)

    mp = MP()
    try:
        fake_abi.install(mp)
        tensors = [torch.randn(s) for s in shapes]
        L = ParamLayout(tensors)
        row = torch.full((L.size,), 0.5)
        kw = {}
        if scale is not None:
            kw = dict(inv_scale=torch.tensor(1.0 / scale), found_inf=torch.zeros(()))
        ops.multi_tensor_copy(row, tensors, L.offsets, mode, table=L.copy_table, **kw)
        factor = 1.0 if scale is None else np.float32(1.0 / scale)
        for t, v in zip(tensors, L.views(row)):
            want = t * factor + (0.5 if mode == 1 else 0.0)
            assert torch.equal(v, want)
        if scale is not None:
            assert kw["found_inf"].item() == 0.0
        out = [torch.empty_like(t) for t in tensors]
        ops.multi_tensor_copy(row, out, L.offsets, 2, table=L.copy_table)
        for o, v in zip(out, L.views(row)):
            assert torch.equal(o, v)
        try:
            ops.multi_tensor_copy(row, tensors[:-1] + [torch.randn(tensors[-1].numel() + 1)], L.offsets, 0, table=L.copy_table)
            raise AssertionError("a tensor of the wrong size must be rejected")
        except ValueError:
            pass
    finally:
        for obj, name, val in reversed(mp.undo):
            setattr(obj, name, val)


@given(samples=st.integers(0, 200), members=st.integers(1, 12))
def test_split_samples_is_the_reference_rule(samples, members):
    counts = split_samples(samples, members)
    per = samples // members
    assert len(counts) == members and sum(counts) == samples
    assert counts[1:] == [per] * (members - 1) and counts[0] == samples - (members - 1) * per   # ensemble.py:37-39


@given(chunks=st.lists(st.integers(1, 9), min_size=1, max_size=6))
def test_reserved_stream_ids_are_the_ids_single_calls_would_take(chunks):
    noise.set_seed(3)
    singles = [noise.next_stream_id() for _ in range(sum(chunks))]
    noise.set_seed(3)
    got = []
    for c in chunks:
        first = noise.reserve_stream_ids(c)
        got.extend(range(first, first + c))
    noise.set_seed(None)
    assert got == singles and len(set(singles)) == len(singles)


@given(shapes=shapes, seed=st.integers(0, 2**16))
@settings(max_examples=40, deadline=None)
def test_column_sharded_model_closures_on_any_parameter_list(shapes, seed):
    """ColumnShardedModel over arbitrary tensor lists (one-rank group, no collective): the flat parameter holds the
    layout's padded vector, a write to it reaches the tensors at the next forward, the backward hands exactly the
    tensors' gradients (padding zero) to param.grad, and accumulates like autograd."""
    from beyond_deep_ensembles_b200 import ColumnShardedModel
    g = torch.Generator().manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    coeff = [torch.randn(s, generator=g) for s in shapes]
    start = [p.detach().clone() for p in params]
    sm = ColumnShardedModel(params)
    L = sm.layout
    assert sm.world == 1 and sm.shard == L.size == sm.param.numel() and (sm.lo, sm.hi) == (0, L.size)
    assert torch.equal(L.to_logical(sm.param.detach()), torch.cat([p.reshape(-1) for p in start]))
    assert float(sm.param.detach().abs().sum()) == float(L.from_logical(L.to_logical(sm.param.detach())).abs().sum())

    def fwd():
        return sum((c * p * p).sum() for c, p in zip(coeff, params))

    def bwd(loss):
        loss.backward()
    f, b = sm.closures(fwd, bwd)
    with torch.no_grad():
        sm.param.mul_(0.5)
    loss = f()
    for p, s0 in zip(params, start):
        assert torch.equal(p.detach(), 0.5 * s0)
    b(loss)
    want = L.from_logical(torch.cat([(2 * c * 0.5 * s0).reshape(-1) for c, s0 in zip(coeff, start)]))
    assert torch.allclose(sm.param.grad, want, rtol=1e-6, atol=1e-7)
    b(f())
    assert torch.allclose(sm.param.grad, 2 * want, rtol=1e-6, atol=1e-7)
    assert sm.collectives == 0
