"""BayesianOptimizer base API and the last-layer wrapper.

Mirrors the reference's public surface (src/algos/algo.py:5-133): same class names, method
names, argument meaning and error behaviour, so training loops written against the reference
(`loss = optimizer.step(forward_closure, backward_closure, grad_scaler=scaler)`,
`optimizer.complete_epoch()`, `optimizer.sample_parameters()`, `get_base_optimizer()`,
`init_grad_scaler()`) run unchanged.  Nothing in this file touches parameter data.
"""
from __future__ import annotations

from typing import Any, Dict

from torch.optim import Optimizer


def _scaler_active(grad_scaler) -> bool:
    return grad_scaler is not None and grad_scaler.is_enabled()


class BayesianOptimizer(Optimizer):
    """An optimizer over a distribution of parameters (reference: algo.py:5-81).

    step() owns the forward/backward passes: `forward_closure()` returns the loss without
    calling backward or clearing gradients, `backward_closure(loss)` runs one backward pass
    (calling grad_scaler.scale(loss) itself when AMP is used).
    """

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        # tells torch's GradScaler that step() deals with the loss scale itself (algo.py:17)
        self._step_supports_amp_scaling = True

    # ---- what subclasses implement (algo.py:19-56) ----
    def step(self, forward_closure, backward_closure):
        raise NotImplementedError()

    def sample_parameters(self):
        raise NotImplementedError()

    def complete_epoch(self):
        """End-of-epoch bookkeeping; nothing by default."""

    def get_base_optimizer(self):
        """The optimizer that makes the actual parameter updates (the one LR schedulers attach to)."""

    # ---- GradScaler plumbing (algo.py:44-49, 65-81) ----
    def init_grad_scaler(self, grad_scaler):
        # GradScalers initialise lazily on the first step/unscale; the optimizers poke at the
        # per-optimizer state before that, so force the initialisation here.
        if _scaler_active(grad_scaler) and grad_scaler._scale is None:
            grad_scaler._lazy_init_scale_growth_tracker(self._params_device())

    def _prepare_and_check_grads(self, grad_scaler, optimizer=None):
        """Unscale the gradients held by `optimizer` (default: self); True when they may be used."""
        if not _scaler_active(grad_scaler):
            return True
        grad_scaler.unscale_(optimizer if optimizer is not None else self)
        # Kept literally from algo.py:73: this reads the OPTIMIZER's state dict (a defaultdict,
        # so the key springs into existence as {}), not the scaler's found-inf record, hence it
        # is always True — the reference's behaviour, which callers have been trained against.
        found = self.state["found_inf_per_device"]
        return sum(flag.item() for flag in found.values()) == 0

    def _set_grad_scaler_state(self, grad_scaler, stage, optimizer=None):
        if _scaler_active(grad_scaler):
            key = id(optimizer if optimizer is not None else self)
            grad_scaler._per_optimizer_states[key]["stage"] = stage

    # ---- parameter access (algo.py:57-63) ----
    def _params(self):
        for group in self.param_groups:
            yield from group["params"]

    def _params_device(self):
        return next(self._params()).device


class LastLayerBayesianOptimizer(BayesianOptimizer):
    """Bayesian last layer + deterministic body (reference: algo.py:83-133).

    The deterministic gradients are zeroed once and ACCUMULATE over all forward/backward
    passes the Bayesian optimizer makes inside its step before the deterministic optimizer
    steps — exactly the reference's order (algo.py:100-103).  GradScalers are refused, as there.
    """

    _PARTS = ("ll_bayesian_optimizer", "deterministic_optimizer")

    def __init__(self, ll_bayesian_optimizer: BayesianOptimizer, deterministic_optimizer: Optimizer):
        # deliberately no super().__init__(): the reference does not call it either
        self.ll_bayesian_optimizer = ll_bayesian_optimizer
        self.deterministic_optimizer = deterministic_optimizer

    def step(self, forward_closure, backward_closure, grad_scaler=None):
        if _scaler_active(grad_scaler):
            raise ValueError("Doesn't support grad scaler")
        body, head = self.deterministic_optimizer, self.ll_bayesian_optimizer
        body.zero_grad()
        loss = head.step(forward_closure, backward_closure)   # >= 1 backward pass: fills the body's gradients too
        body.step()
        return loss

    def init_grad_scaler(self, grad_scaler):
        if grad_scaler.is_enabled():
            raise RuntimeError("Doesn't support grad scaler")

    def get_base_optimizer(self):
        raise RuntimeError("There is no defined base optimizer on the ll optimizer. Call get_base_optimizer directly "
                           "on the passed ll bayesian optimizer")

    def complete_epoch(self):
        self.ll_bayesian_optimizer.complete_epoch()

    def sample_parameters(self):
        self.ll_bayesian_optimizer.sample_parameters()

    def presample(self, count):
        """Forward DeepEnsemble.predict's batch announcement to the Bayesian part (no-op if it cannot batch)."""
        announce = getattr(self.ll_bayesian_optimizer, "presample", None)
        if announce is not None:
            announce(count)

    def state_dict(self) -> Dict[str, Any]:
        return {part: getattr(self, part).state_dict() for part in self._PARTS}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        for part in self._PARTS:
            getattr(self, part).load_state_dict(state_dict[part])

    def __repr__(self) -> str:
        bar = "=" * 34
        return (f"LL Bayesian Optimizer: \n\n{self.ll_bayesian_optimizer!r}\n{bar}\n"
                f"Deterministic Optimizer:\n\n{self.deterministic_optimizer!r}")
