"""One data-parallel training step: the body of the reference's training loop
(engine.py:120-165 `train_one_epoch`: H2D of the batch, forward, criterion,
backward with the gradient all-reduce, clip_grad_norm_, AdamW) as a reusable
object, laid out for B200:

  * all trainable parameters live in ONE flat fp32 buffer and their gradients in
    another, so the data-parallel exchange is a single NCCL all-reduce over
    NVLink/NVSwitch (98 MB at the CoDA configuration; the reference's DDP splits
    it into 25 MB buckets), the global-norm clip is one reduction and AdamW one
    fused kernel over one tensor (weight decay is uniform, optimizer.py:4-36);
  * BatchNorm statistics stay per GPU (no SyncBatchNorm): the step has exactly
    one gradient collective plus the 4-byte box-count all-reduce of the criterion;
  * the batch is copied from pinned host memory on the compute stream and the loss
    is left on the device: the caller decides when to synchronise.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import ops
from .utils.dist import get_world_size, is_distributed


def adjust_learning_rate(args, optimizer, curr_epoch: float) -> float:
    """Cosine schedule with linear warm-up (reference engine.py:33-55)."""
    if curr_epoch <= args.warm_lr_epochs and args.warm_lr_epochs > 0:
        lr = args.warm_lr + (curr_epoch / args.warm_lr_epochs) * (args.base_lr - args.warm_lr)
    else:
        lr = args.final_lr + 0.5 * (args.base_lr - args.final_lr) * (1 + math.cos(math.pi * curr_epoch / args.max_epoch))
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr


class FlatParameters:
    """Re-homes every trainable parameter of `module` (and its gradient) into one
    contiguous fp32 buffer each."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        assert params and all(p.dtype == torch.float32 for p in params)
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.flat_param = torch.nn.Parameter(torch.empty(total, dtype=torch.float32, device=dev))
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat_param.data[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param.data[off:off + n].view_as(p)
                p.grad = self.flat_grad[off:off + n].view_as(p)
                off += n
        self.flat_param.grad = self.flat_grad
        self.params = params

    def zero_grad(self):
        self.flat_grad.zero_()

    def nbytes(self) -> int:
        return self.flat_grad.numel() * 4


class TrainStep:
    def __init__(self, args, model, criterion, device):
        self.args, self.model, self.criterion, self.device = args, model, criterion, device
        self.flat = FlatParameters(model)
        self.optimizer = torch.optim.AdamW([self.flat.flat_param], lr=args.base_lr, weight_decay=args.weight_decay,
                                           fused=True)
        self.world = get_world_size()

    def to_device(self, batch_host: dict) -> dict:
        """engine.py:125-129: every tensor of the collated batch to the device (async from pinned memory)."""
        return {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                for k, v in batch_host.items()}

    def __call__(self, batch: dict, curr_epoch: float = 0.0):
        """One optimiser step on a batch that is already on the device.  Returns (loss, loss_dict)."""
        adjust_learning_rate(self.args, self.optimizer, curr_epoch)
        self.flat.zero_grad()
        outputs = self.model(batch, curr_epoch=int(curr_epoch))
        loss, loss_dict = self.criterion(outputs, batch)
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat.flat_grad)          # the single gradient collective
            self.flat.flat_grad.div_(self.world)
        if self.args.clip_gradient > 0:
            torch.nn.utils.clip_grad_norm_([self.flat.flat_param], self.args.clip_gradient)
        self.optimizer.step()
        ops.invalidate_weight_cache()  # packed bf16 weight planes are stale now
        return loss.detach(), loss_dict
