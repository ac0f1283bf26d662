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


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """The step's single gradient collective: sum over ranks (NCCL over NVLink on GPUs), then /world."""
    if is_distributed() and get_world_size() > 1:
        dist.all_reduce(flat)
        flat.div_(get_world_size())
    return flat


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
    """The step runs either eagerly or, after `capture()`, as ONE CUDA graph: the ~6 000 kernel
    launches of a step (8 decoder layers x many small ops) are launch-bound on the host otherwise.
    Everything inside the step is static-shaped and free of host synchronisation (the criterion
    matches on the GPU); the only per-step host work -- the learning rate, the random choice of the
    32 boxes per scene that get CLIP crops, the batch upload -- happens before the replay and is
    handed over through persistent device buffers."""

    def __init__(self, args, model, criterion, device):
        self.args, self.model, self.criterion, self.device = args, model, criterion, device
        self.flat = FlatParameters(model)
        self.lr = torch.tensor(float(args.base_lr), device=device)
        self.optimizer = torch.optim.AdamW([self.flat.flat_param], lr=self.lr, weight_decay=args.weight_decay,
                                           fused=True, capturable=True)
        self.world = get_world_size()
        self.graph = None
        self.launches_per_step = None
        self.static_batch = None
        self.static_out = None
        nsel = getattr(model, "distillation_box_num", 32)
        self._sel_host = None
        self._nsel = nsel

    def to_device(self, batch_host: dict) -> dict:
        """engine.py:125-129: every tensor of the collated batch to the device (async from pinned memory)."""
        return {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                for k, v in batch_host.items()}

    def _set_lr(self, curr_epoch: float):
        a = self.args
        if curr_epoch <= a.warm_lr_epochs and a.warm_lr_epochs > 0:
            lr = a.warm_lr + (curr_epoch / a.warm_lr_epochs) * (a.base_lr - a.warm_lr)
        else:
            lr = a.final_lr + 0.5 * (a.base_lr - a.final_lr) * (1 + math.cos(math.pi * curr_epoch / a.max_epoch))
        self.lr.fill_(lr)
        return lr

    def _draw_selection(self, bsz: int):
        """Host RNG draw of the CLIP-crop boxes for the coming step -> persistent device buffer."""
        m = self.model
        if not getattr(m, "if_with_clip_train", False):
            return
        sel = torch.from_numpy(m.draw_box_selection(bsz))
        if self._sel_host is None or self._sel_host.shape != sel.shape:
            self._sel_host = torch.empty_like(sel).pin_memory()
            m.external_selection = torch.empty_like(sel, device=self.device)
        self._sel_host.copy_(sel)
        m.external_selection.copy_(self._sel_host, non_blocking=True)

    def capture(self, example_batch: dict, warmup: int = 3):
        """Captures the whole step into a CUDA graph (static copy of `example_batch` as input)."""
        self.static_batch = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        bsz = self.static_batch["point_clouds"].shape[0]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._draw_selection(bsz)
                self._body(self.static_batch, 0)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._draw_selection(bsz)
        ops.invalidate_weight_cache()
        from . import _lib

        n0 = _lib.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._body(self.static_batch, 0)
        self.launches_per_step = _lib.LAUNCHES - n0  # C-ABI kernel launches recorded in the graph
        return self

    def __call__(self, batch: dict, curr_epoch: float = 0.0):
        """One optimiser step on a batch that is already on the device.  Returns (loss, loss_dict)."""
        self._set_lr(curr_epoch)
        bsz = batch["point_clouds"].shape[0]
        self._draw_selection(bsz)
        if self.graph is not None:
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    self.static_batch[k].copy_(v, non_blocking=True)
            self.graph.replay()
            return self.static_out
        return self._body(batch, int(curr_epoch))

    def _body(self, batch: dict, curr_epoch: int):
        from . import attention_launch

        attention_launch.advance_seed(self.device)
        self.flat.zero_grad()
        outputs = self.model(batch, curr_epoch=int(curr_epoch))
        loss, loss_dict = self.criterion(outputs, batch)
        loss.backward()
        allreduce_mean_(self.flat.flat_grad)          # the single gradient collective
        if self.args.clip_gradient > 0:
            torch.nn.utils.clip_grad_norm_([self.flat.flat_param], self.args.clip_gradient)
        self.optimizer.step()
        ops.invalidate_weight_cache()  # packed bf16 weight planes are stale now
        return loss.detach(), loss_dict
