"""BayesianOptimizer base API and the last-layer wrapper.

Mirrors the reference's public surface (src/algos/algo.py:5-133): same class names, method
names, argument meaning and error behaviour, so training loops written against the reference
(`loss = optimizer.step(forward_closure, backward_closure, grad_scaler=scaler)`,
`optimizer.complete_epoch()`, `optimizer.sample_parameters()`, `get_base_optimizer()`,
`init_grad_scaler()`) run unchanged.  Nothing in this file touches parameter data.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch.optim import Optimizer


class BayesianOptimizer(Optimizer):
    """An optimizer over a distribution of parameters (reference: algo.py:5-81).

    step() owns the forward/backward passes: `forward_closure()` returns the loss without
    calling backward or clearing gradients, `backward_closure(loss)` runs one backward pass
    (calling grad_scaler.scale(loss) itself when AMP is used).
    """

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._step_supports_amp_scaling = True

    def step(self, forward_closure, backward_closure):
        raise NotImplementedError()

    def complete_epoch(self):
        pass

    def sample_parameters(self):
        raise NotImplementedError()

    def init_grad_scaler(self, grad_scaler):
        # GradScalers initialise lazily on the first step/unscale; the optimizers poke at the
        # per-optimizer state before that (algo.py:44-49).
        if grad_scaler is not None and grad_scaler.is_enabled() and grad_scaler._scale is None:
            grad_scaler._lazy_init_scale_growth_tracker(self._params_device())

    def get_base_optimizer(self):
        pass

    def _params_device(self):
        return self.param_groups[0]["params"][0].device

    def _params(self):
        for group in self.param_groups:
            for param in group["params"]:
                yield param

    def _prepare_and_check_grads(self, grad_scaler, optimizer=None):
        if grad_scaler is None or not grad_scaler.is_enabled():
            return True
        opt = self if optimizer is None else optimizer
        grad_scaler.unscale_(opt)
        # Kept literally from algo.py:73: this reads the OPTIMIZER's state dict (a defaultdict,
        # so the key springs into existence as {}), not the scaler's found-inf record, hence it
        # is always True — the reference's behaviour, which callers have been trained against.
        return sum(v.item() for v in self.state["found_inf_per_device"].values()) == 0

    def _set_grad_scaler_state(self, grad_scaler, stage, optimizer=None):
        if grad_scaler is None or not grad_scaler.is_enabled():
            return
        opt = self if optimizer is None else optimizer
        grad_scaler._per_optimizer_states[id(opt)]["stage"] = stage


class LastLayerBayesianOptimizer(BayesianOptimizer):
    """Bayesian last layer + deterministic body (reference: algo.py:83-133).

    The deterministic gradients are zeroed once and ACCUMULATE over all forward/backward
    passes the Bayesian optimizer makes inside its step before the deterministic optimizer
    steps — exactly the reference's order (algo.py:100-103).
    """

    def __init__(self, ll_bayesian_optimizer: BayesianOptimizer, deterministic_optimizer: Optimizer):
        # deliberately no super().__init__(): the reference does not call it either
        self.ll_bayesian_optimizer = ll_bayesian_optimizer
        self.deterministic_optimizer = deterministic_optimizer

    def step(self, forward_closure, backward_closure, grad_scaler=None):
        if grad_scaler is not None and grad_scaler.is_enabled():
            raise ValueError("Doesn't support grad scaler")
        self.deterministic_optimizer.zero_grad()
        loss = self.ll_bayesian_optimizer.step(forward_closure, backward_closure)
        self.deterministic_optimizer.step()
        return loss

    def complete_epoch(self):
        self.ll_bayesian_optimizer.complete_epoch()

    def sample_parameters(self):
        self.ll_bayesian_optimizer.sample_parameters()

    def init_grad_scaler(self, grad_scaler):
        if grad_scaler.is_enabled():
            raise RuntimeError("Doesn't support grad scaler")

    def get_base_optimizer(self):
        raise RuntimeError("There is no defined base optimizer on the ll optimizer. Call get_base_optimizer directly "
                           "on the passed ll bayesian optimizer")

    def state_dict(self) -> Dict[str, Any]:
        return {
            "ll_bayesian_optimizer": self.ll_bayesian_optimizer.state_dict(),
            "deterministic_optimizer": self.deterministic_optimizer.state_dict(),
        }

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self.ll_bayesian_optimizer.load_state_dict(state_dict["ll_bayesian_optimizer"])
        self.deterministic_optimizer.load_state_dict(state_dict["deterministic_optimizer"])

    def __repr__(self) -> str:
        return ("LL Bayesian Optimizer: \n\n" + repr(self.ll_bayesian_optimizer) +
                "\n==================================\nDeterministic Optimizer:\n\n" + repr(self.deterministic_optimizer))
