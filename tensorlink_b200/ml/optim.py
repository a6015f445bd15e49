"""Distributed optimizer front end (/root/reference/tensorlink/ml/optim.py:81-205).

The reference subclasses the user's optimizer class and fans ``step``/``zero_grad`` out to workers as OPTIMIZER
packets acknowledged by 1-second polling (:131-203) — and skips ``offloaded_group`` shards entirely (:114,144,172); on
the worker the step is the plain ``torch.optim`` step of whatever class the user named (ml/worker.py:1309-1327).
Here every rank owns its stage's flat parameter / gradient arenas, so there is no wire traffic:

* ``Adam`` / ``AdamW`` (and ``None``): ``StageAdam`` — one fused AdamW kernel over the arena (``tl_adamw_step``), fp32 moments;
* any other ``torch.optim.Optimizer`` subclass (SGD, RMSprop, Adagrad, ...): ``StageTorchOptimizer`` — exactly what the
  reference's worker does: that class's own ``step()`` on the device, over ONE parameter that is the whole bf16 arena with
  the gradient arena as its ``.grad`` (element-wise optimizers do not care how the parameters are grouped).

``scheduler_type`` is stored on the model and not used, like in the reference (ml/module.py:315-317).
"""
from __future__ import annotations

import torch


class StageTorchOptimizer:
    """``create_optimizer(**kw)`` result for optimizer classes without a fused kernel here."""

    def __init__(self, dm, optimizer_type, **optimizer_kwargs):
        from .train import _trainer
        self.dm = dm
        p = dm.stage.params
        if p.grad is None:
            raise RuntimeError("DistributedModel was built with training=False; no gradient arena")
        self._trainer = _trainer
        self.param = torch.nn.Parameter(p.flat, requires_grad=True)        # shares the arena's storage
        self.param.grad = p.grad
        self.opt = optimizer_type([self.param], **optimizer_kwargs)
        self.param_groups = self.opt.param_groups

    def zero_grad(self, set_to_none: bool = False):
        self._trainer(self.dm).zero_grad()                # (never set_to_none: the gradient arena is the kernels' target)

    def step(self, closure=None):
        tr = self._trainer(self.dm)
        tr.settle_grads()                                 # lazily-zeroed matrices get their zeros before anything reads them
        self.param.grad = self.dm.stage.params.grad
        return self.opt.step(closure)


def create_distributed_optimizer(model, optimizer_type=None, **optimizer_kwargs):
    if optimizer_type is not None and not (isinstance(optimizer_type, type) and issubclass(optimizer_type, torch.optim.Optimizer)):
        raise TypeError(f"optimizer must be a torch.optim.Optimizer subclass (or None for Adam), got {optimizer_type!r}")
    name = optimizer_type.__name__ if optimizer_type is not None else "Adam"
    if name in ("Adam", "AdamW"):
        from .train import StageAdam
        return StageAdam(model, decoupled=(name == "AdamW"), **optimizer_kwargs)
    return StageTorchOptimizer(model, optimizer_type, **optimizer_kwargs)
