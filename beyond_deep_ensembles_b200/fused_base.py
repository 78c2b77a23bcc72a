"""Running the caller's torch.optim base optimizer INSIDE the SVGD apply kernel.

Reference behaviour being reproduced (src/algos/svgd.py:92-103): after the posterior update the
ONE base optimizer the caller passed in — its state (momentum buffer / Adam moments / step count)
shared by all particles — takes one `step()` per particle, in particle order, on
`param.data = particle_i`, `param.grad = new_gradient_i`.  That is n optimizer passes over D plus
2·n·(#tensors) Python rebinding operations per SVGD step.

`FusedBasePlan` recognises the stock optimizers the reference's experiments use
(torch.optim.SGD with momentum / Nesterov / weight decay — experiments/cifar/models.py:82;
torch.optim.Adam / AdamW — experiments/uci/models.py, experiments/civilcomments/models.py:106) and
runs those n sequential steps in registers inside K2 (`bde_svgd_apply_sgd` / `bde_svgd_apply_adam`):
X is updated in place and the [n, D] gradient matrix never reaches HBM.  The caller's optimizer
object stays the owner of hyper-parameters and state: `param_groups[*]["lr"]` is read every step
(LR schedulers keep working), `state[param]["momentum_buffer" | "exp_avg" | "exp_avg_sq"]` are views
into flat arenas that the kernel updates, `state[param]["step"]` advances by n per SVGD step, so
`base_optimizer.state_dict()` / `load_state_dict()` round-trip exactly as with the reference.

Anything else (another optimizer class or a subclass, amsgrad / maximize / capturable /
differentiable, tensor learning rates, step hooks, an enabled GradScaler, parameter groups that are
not contiguous in parameter order) is not fused: the SVGD optimizer then calls `base.step()` once
per particle exactly like the reference.  That is the reference's own code path on CUDA tensors,
not a CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from .layout import ParamLayout

_SGD, _ADAM, _ADAMW = "sgd", "adam", "adamw"


def _has_hooks(base) -> bool:
    for name in ("_optimizer_step_pre_hooks", "_optimizer_step_post_hooks"):
        if len(getattr(base, name, {}) or {}) > 0:
            return True
    try:
        from torch.optim import optimizer as _o
        if len(getattr(_o, "_global_optimizer_pre_hooks", {})) or len(getattr(_o, "_global_optimizer_post_hooks", {})):
            return True
    except Exception:  # noqa: BLE001
        pass
    return False


def _plain_number(v) -> bool:
    return isinstance(v, (int, float)) and not isinstance(v, bool)


class FusedBasePlan:
    """Column segments of the flat particle arena, one per base-optimizer param group."""

    def __init__(self, base, kind: str, segments, plist, layout: ParamLayout, device):
        self.base, self.kind, self.segments = base, kind, segments
        self.plist, self.layout = plist, layout
        self.state0 = torch.zeros(layout.size, dtype=torch.float32, device=device)
        self.state1 = torch.zeros(layout.size, dtype=torch.float32, device=device) if kind != _SGD else None
        self.views0 = layout.views(self.state0)
        self.views1 = layout.views(self.state1) if self.state1 is not None else None
        self._uses_state0 = kind != _SGD or any(g["momentum"] != 0 for _, _, g in segments)

    # ------------------------------------------------------------------ recognition
    @staticmethod
    def build(base, plist, layout: ParamLayout, device):
        """A plan, or None when `base` is not one of the recognised stock optimizers."""
        t = type(base)
        if t is torch.optim.SGD:
            kind = _SGD
        elif t is torch.optim.Adam:
            kind = _ADAM
        elif t is torch.optim.AdamW:
            kind = _ADAMW
        else:
            return None
        if _has_hooks(base):
            return None
        index = {id(p): k for k, p in enumerate(plist)}
        seen, segments = set(), []
        for group in base.param_groups:
            if not FusedBasePlan._group_ok(kind, group):
                return None
            ks = sorted(index.get(id(p), -1) for p in group["params"])
            if not ks or ks[0] < 0 or ks != list(range(ks[0], ks[-1] + 1)) or seen.intersection(ks):
                return None  # foreign parameter, or a group that is not a contiguous run of our tensors
            seen.update(ks)
            c0 = layout.offsets[ks[0]]
            c1 = layout.offsets[ks[-1] + 1] if ks[-1] + 1 < len(plist) else layout.size
            segments.append((c0, c1, group))
        if len(seen) != len(plist):
            return None
        return FusedBasePlan(base, kind, segments, plist, layout, device)

    @staticmethod
    def _group_ok(kind, g) -> bool:
        if g.get("maximize") or g.get("differentiable") or g.get("fused"):
            return False
        if not _plain_number(g["lr"]) or not _plain_number(g["weight_decay"]):
            return False
        if kind == _SGD:
            return all(_plain_number(g[k]) for k in ("momentum", "dampening"))
        if g.get("amsgrad") or g.get("capturable"):
            return False
        return all(_plain_number(b) for b in g["betas"]) and _plain_number(g["eps"])

    def still_valid(self) -> bool:
        """Hyper-parameters can change between steps (schedulers); re-check the cheap invariants."""
        if _has_hooks(self.base) or len(self.base.param_groups) != len(self.segments):
            return False
        return all(g is seg[2] and self._group_ok(self.kind, g) for g, seg in zip(self.base.param_groups, self.segments))

    # ------------------------------------------------------------------ state aliasing
    def bind_state(self):
        """Make base.state[param] alias the flat arenas.  Returns None if the optimizer's state is
        inconsistent (then the caller takes the unfused path), else (initialized, step0)."""
        state = self.base.state
        if self.kind == _SGD:
            if not self._uses_state0:
                return False, 0
            bufs = [state[p].get("momentum_buffer") if p in state else None for p in self.plist]
            have = sum(b is not None for b in bufs)
            if have == 0:
                return False, 0
            if have != len(bufs):
                return None
            for p, b, v in zip(self.plist, bufs, self.views0):
                if b.data_ptr() != v.data_ptr():
                    v.copy_(b)
                    state[p]["momentum_buffer"] = v
            return True, 0
        entries = [state[p] if p in state else None for p in self.plist]
        have = sum(bool(e) for e in entries)
        if have == 0:
            self.state0.zero_()
            self.state1.zero_()
            for p, v0, v1 in zip(self.plist, self.views0, self.views1):
                state[p]["step"] = torch.tensor(0.0, dtype=torch.float32)
                state[p]["exp_avg"] = v0
                state[p]["exp_avg_sq"] = v1
            return True, 0
        if have != len(entries):
            return None
        steps = set()
        for p, e, v0, v1 in zip(self.plist, entries, self.views0, self.views1):
            if e["exp_avg"].data_ptr() != v0.data_ptr():
                v0.copy_(e["exp_avg"])
                e["exp_avg"] = v0
            if e["exp_avg_sq"].data_ptr() != v1.data_ptr():
                v1.copy_(e["exp_avg_sq"])
                e["exp_avg_sq"] = v1
            st = e["step"]
            if torch.is_tensor(st) and st.is_cuda:
                return None
            steps.add(float(st))
        if len(steps) != 1:
            return None
        return True, int(steps.pop())

    # ------------------------------------------------------------------ launch
    def launch(self, X: torch.Tensor, G: torch.Tensor, sc, out_last: torch.Tensor, initialized: bool, step0: int,
               next_kernel: "ops.NextKernel | None" = None) -> bool:
        """Run the fused K2 + base-optimizer launch(es).  With `next_kernel` and a single column segment
        the training-step form is used; returns True when sc.dist then holds the pair distances of the
        updated particles."""
        n = X.shape[0]
        nk = next_kernel if (next_kernel is not None and len(self.segments) == 1) else None
        for c0, c1, g in self.segments:
            Xs, Gs, ol = X[:, c0:c1], G[:, c0:c1], out_last[c0:c1]
            if self.kind == _SGD:
                ops.svgd_apply_sgd(Xs, Gs, sc, self.state0[c0:c1] if g["momentum"] != 0 else None,
                                   buf_initialized=initialized, lr=g["lr"], momentum=g["momentum"],
                                   dampening=g["dampening"], weight_decay=g["weight_decay"], nesterov=g["nesterov"],
                                   out_last=ol, next_kernel=nk)
            else:
                decoupled = self.kind == _ADAMW or bool(g.get("decoupled_weight_decay"))
                ops.svgd_apply_adam(Xs, Gs, sc, self.state0[c0:c1], self.state1[c0:c1], step0=step0, lr=g["lr"],
                                    beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"],
                                    weight_decay=g["weight_decay"], decoupled_weight_decay=decoupled, out_last=ol,
                                    next_kernel=nk)
        state = self.base.state
        if self.kind == _SGD:
            if self._uses_state0 and not initialized:
                for p, v in zip(self.plist, self.views0):
                    state[p]["momentum_buffer"] = v
        else:
            torch._foreach_add_([state[p]["step"] for p in self.plist], float(n))
        # lr_scheduler's "step() before optimizer.step()" check looks at this flag
        self.base._opt_called = True
        return nk is not None
