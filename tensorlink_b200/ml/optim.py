"""Distributed optimizer front end (/root/reference/tensorlink/ml/optim.py:81-205).

The reference subclasses the user's optimizer class and fans ``step``/``zero_grad`` out to workers as OPTIMIZER
packets acknowledged by 1-second polling (:131-203) — and skips ``offloaded_group`` shards entirely (:114,144,172).
Here every rank owns its stage's flat parameter / gradient arenas, so ``step`` is one fused AdamW kernel per rank
(``tl_adamw_step``) with no wire traffic.  ``optimizer_type`` other than Adam/AdamW/None raises.
"""
from __future__ import annotations

import torch


def create_distributed_optimizer(model, optimizer_type=None, **optimizer_kwargs):
    name = getattr(optimizer_type, "__name__", "Adam") if optimizer_type is not None else "Adam"
    if name not in ("Adam", "AdamW"):
        raise NotImplementedError(f"only Adam/AdamW are implemented on the B200 stage, got {name}")
    from .train import StageAdam
    return StageAdam(model, decoupled=(name == "AdamW"), **optimizer_kwargs)
