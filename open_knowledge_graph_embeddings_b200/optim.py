"""Optimizer regime of the B200 path — drop-in for ``utils/optim.py``.

``OptimRegime`` keeps the reference's behaviour, including the quirk that shapes the effective
hyper-parameters: it bootstraps ``Adam(params, lr=0)`` (utils/optim.py:29) and, on the first
``update``, re-instantiates the configured optimizer FROM THE EXISTING param_groups (:143-145), so the
new optimizer inherits Adam's ``eps = 1e-8`` (and ``betas`` ...) instead of its own defaults.

``Adagrad`` and ``Adam`` below are ``torch.optim.Optimizer`` subclasses whose ``step`` is one
HBM-bound native kernel per parameter tensor (``okge_adagrad_dense`` / ``okge_adam_dense``), with the
update formula of ``torch.optim.Adagrad`` / ``torch.optim.Adam`` — dense over every row, like the
reference runs them.
"""
from __future__ import annotations

import logging
import re
from copy import deepcopy
from typing import Dict, List

import torch

from . import kernels as K
from .functional import mark_table_updated


def _take_deferred(p):
    """A factored table gradient left on the parameter by the backward pass (functional.DeferredTableGrad)."""
    d = getattr(p, "_okge_deferred", None)
    if d is not None:
        p._okge_deferred = None
    return d


class _DeferredAware(torch.optim.Optimizer):
    def zero_grad(self, set_to_none: bool = True):
        for group in self.param_groups:
            for p in group["params"]:
                d = getattr(p, "_okge_deferred", None)
                if d is not None:
                    if hasattr(d, "discard"):
                        d.discard()                  # e.g. release the slot map of an unconsumed compact gradient
                    p._okge_deferred = None
        super().zero_grad(set_to_none=set_to_none)


class Adagrad(_DeferredAware):
    """torch.optim.Adagrad semantics (defaults lr=1e-2, lr_decay=0, weight_decay=0,
    initial_accumulator_value=0, eps=1e-10); state: ``step`` and ``sum``."""

    def __init__(self, params, lr=1e-2, lr_decay=0, weight_decay=0, initial_accumulator_value=0, eps=1e-10):
        defaults = dict(lr=lr, lr_decay=lr_decay, eps=eps, weight_decay=weight_decay,
                        initial_accumulator_value=initial_accumulator_value)
        super().__init__(params, defaults)
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                st["step"] = 0
                st["sum"] = torch.full_like(p, float(group["initial_accumulator_value"]),
                                            memory_format=torch.preserve_format)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                deferred = _take_deferred(p)
                if deferred is not None and p.grad is None:
                    # fused path: dE contraction + this step in one pass, the gradient is never materialised
                    st = self.state[p]
                    st["step"] += 1
                    clr = group["lr"] / (1 + (st["step"] - 1) * group["lr_decay"])
                    deferred.adagrad_step(p, st["sum"], clr, group["eps"], group["weight_decay"])
                    continue
                if deferred is not None:
                    p.grad = p.grad + deferred.materialize()
                if p.grad is None:
                    continue
                st = self.state[p]
                st["step"] += 1
                clr = group["lr"] / (1 + (st["step"] - 1) * group["lr_decay"])
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                K.adagrad_dense(p.data, grad, st["sum"], clr, group["eps"], group["weight_decay"])
                mark_table_updated(p)
        return loss


class Adam(_DeferredAware):
    """torch.optim.Adam semantics without amsgrad; state: ``step``, ``exp_avg``, ``exp_avg_sq``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unused):
        if amsgrad:
            raise NotImplementedError("amsgrad is not part of the accelerated path")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                deferred = _take_deferred(p)
                if deferred is not None and deferred.__class__.__name__ == "RowsGrad":
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if deferred is not None:         # no fused Adam epilogue: build the dense gradient
                    dense = deferred.materialize()
                    p.grad = dense if p.grad is None else p.grad + dense
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                K.adam_dense(p.data, grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                             group["weight_decay"], st["step"])
                mark_table_updated(p)
        return loss


OPTIMIZERS: Dict[str, type] = {"Adagrad": Adagrad, "Adam": Adam}


def _optimizer_class(name: str):
    if name in OPTIMIZERS:
        return OPTIMIZERS[name]
    raise NotImplementedError(f"optimizer {name!r} is not on the accelerated path (available: {sorted(OPTIMIZERS)})")


class OptimRegime(object):
    """Reconfigures the optimizer according to a list of phases (utils/optim.py:14-217)."""

    def __init__(self, params, optimization_config, filter_weight_decay=False, lr_scheduler_config=None):
        self.optimizer = Adam(params, lr=0)                                   # utils/optim.py:29
        if not isinstance(optimization_config, list):
            optimization_config = dict(optimization_config)
            optimization_config['epoch'] = 0
            optimization_config = [optimization_config]
        if lr_scheduler_config is not None and not isinstance(lr_scheduler_config, list):
            lr_scheduler_config = dict(lr_scheduler_config)
            lr_scheduler_config['epoch'] = 0
            lr_scheduler_config = [lr_scheduler_config]
        if filter_weight_decay:
            optimization_config = [{k: v for k, v in step.items() if k != 'weight_decay'} for step in optimization_config]
        self.optimization_config = optimization_config
        self.lr_scheduler_config = lr_scheduler_config
        self.lr_scheduler = None
        self.setting = {}
        self.current_optimization_config_phase = None

    @staticmethod
    def setup_optimizer_regime(args, model) -> List["OptimRegime"]:
        """utils/optim.py:45-81: one regime per optimization_config entry, parameters selected by the
        optional ``match`` regex on their names."""
        optimizers = []
        ocs = args["optimization_config"]
        ocs = list(ocs) if isinstance(ocs, (list, tuple)) else [ocs]
        if "lr_scheduler_config" in args:
            lcs = args["lr_scheduler_config"]
            lcs = list(lcs) if isinstance(lcs, (list, tuple)) else [lcs]
        else:
            lcs = [None] * len(ocs)                                          # the reference zips with {} here
        for oc, lc in zip(ocs, lcs):
            params, names = [], []
            for name, p in model.named_parameters():
                if not p.requires_grad:
                    continue
                if "match" in oc and re.search(oc["match"], name) is None:
                    continue
                params.append(p)
                names.append(name)
            if params:
                optimizers.append(OptimRegime(params, optimization_config=oc, lr_scheduler_config=lc))
            logging.info("PARAMS for optimization config %s: %s", oc, names)
        return optimizers

    def update(self, epoch, train_steps, update_optimizer=False):
        """utils/optim.py:104-134."""
        if self.optimization_config is None:
            return
        if self.current_optimization_config_phase is None:
            update_optimizer = True
            for phase, setting in enumerate(self.optimization_config):
                if epoch >= setting.get('epoch', 0) or train_steps >= setting.get('step', 0):
                    self.current_optimization_config_phase = phase
                    break
        if len(self.optimization_config) > self.current_optimization_config_phase + 1:
            nxt = self.current_optimization_config_phase + 1
            if (epoch >= self.optimization_config[nxt].get('epoch', float('inf'))
                    or train_steps >= self.optimization_config[nxt].get('step', float('inf'))):
                self.current_optimization_config_phase = nxt
                update_optimizer = True
        if update_optimizer:                 # (the reference copies the configs on every call; only this branch reads them)
            optimizer_config = deepcopy(self.optimization_config[self.current_optimization_config_phase])
            lr_scheduler_config = None
            if self.lr_scheduler_config is not None:
                lr_scheduler_config = deepcopy(self.lr_scheduler_config[self.current_optimization_config_phase])
            self.adjust(optimizer_config, lr_scheduler_config)

    def get_current_setting(self):
        return deepcopy(self.optimization_config[self.current_optimization_config_phase])

    def adjust(self, optimizer_config: dict, lr_scheduler_config: dict):
        """utils/optim.py:139-160. The new optimizer is built on the OLD param_groups, so keys that both
        optimizers share (eps, weight_decay, lr ...) keep their previous values unless the config
        overrides them."""
        if 'optimizer' in optimizer_config:
            self.optimizer = _optimizer_class(optimizer_config['optimizer'])(self.optimizer.param_groups)
            logging.info('OPTIMIZER - setting method = %s', optimizer_config['optimizer'])
        for group in self.optimizer.param_groups:
            for key in group.keys():
                if key in optimizer_config and optimizer_config[key] != group[key]:
                    logging.info('OPTIMIZER - setting %s = %s', key, optimizer_config[key])
                    group[key] = optimizer_config[key]
        if lr_scheduler_config is not None and 'lr_scheduler' in lr_scheduler_config:
            method = torch.optim.lr_scheduler.__dict__[lr_scheduler_config.pop('lr_scheduler')]
            lr_scheduler_config.pop('epoch', None)
            self.lr_scheduler = method(optimizer=self.optimizer, **lr_scheduler_config)
        self.setting = deepcopy(optimizer_config)

    def state_dict(self):
        return {'optimizer_state': self.optimizer.state_dict(), 'regime': self.optimization_config}

    def load_state_dict(self, state_dict, epoch, train_steps, reset_optimizer=True):
        self.optimization_config = state_dict['regime']
        self.update(epoch, train_steps, update_optimizer=True)
        if not reset_optimizer:
            self.optimizer.load_state_dict(state_dict['optimizer_state'])

    def zero_grad(self):
        self.optimizer.zero_grad()

    def step(self, closure=None):
        self.optimizer.step(closure)

    def lr_scheduler_step(self, criterion_value, epoch):
        if self.lr_scheduler is not None:
            self.lr_scheduler.step(criterion_value)

    def add_param_group(self, param_group):
        self.optimizer.add_param_group(param_group)
