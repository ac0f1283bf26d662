"""torch.distributed helpers with the reference's names (utils/dist.py:9-185).

One process per GPU, NCCL over NVLink; every helper degrades to the identity
when no process group is initialised (single-GPU runs), as the reference does.
"""
from __future__ import annotations

import pickle

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def is_primary() -> bool:
    return get_rank() == 0


def get_world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


def barrier():
    if is_distributed():
        dist.barrier()


def setup_print_for_distributed(is_primary_flag: bool):
    """Silences print() on non-primary ranks unless force=True is passed."""
    import builtins

    builtin_print = builtins.print

    def print(*args, **kwargs):  # noqa: A001
        if kwargs.pop("force", False) or is_primary_flag:
            builtin_print(*args, **kwargs)

    builtins.print = print


def init_distributed(gpu_id, global_rank, world_size, dist_url, dist_backend):
    torch.cuda.set_device(gpu_id)
    dist.init_process_group(backend=dist_backend, init_method=dist_url, world_size=world_size, rank=global_rank)
    dist.barrier()
    setup_print_for_distributed(is_primary())


def all_reduce_sum(tensor):
    if not is_distributed():
        return tensor
    out = tensor.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def all_reduce_average(tensor):
    val = all_reduce_sum(tensor)
    return val / get_world_size()


def reduce_dict(input_dict, average: bool = True):
    """All-reduces the (sorted-key) values of a dict of scalars in one collective."""
    world_size = get_world_size()
    if world_size < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k] for k in names], dim=0)
        dist.all_reduce(values)
        if average:
            values /= world_size
        return dict(zip(names, values))


def all_gather_pickle(data, max_size, device):
    """all_gather of an arbitrary picklable object through fixed-size byte tensors."""
    world_size = get_world_size()
    if world_size == 1:
        return [data]
    buffer = torch.ByteTensor(torch.ByteStorage.from_buffer(pickle.dumps(data))).to(device)
    local_size = torch.tensor([buffer.numel()], device=device)
    size_list = [torch.tensor([0], device=device) for _ in range(world_size)]
    dist.all_gather(size_list, local_size)
    size_list = [int(s.item()) for s in size_list]
    assert max(size_list) <= max_size
    tensor_list = [torch.empty((max_size,), dtype=torch.uint8, device=device) for _ in size_list]
    if local_size != max_size:
        buffer = torch.cat((buffer, torch.empty((max_size - int(local_size),), dtype=torch.uint8, device=device)))
    dist.all_gather(tensor_list, buffer)
    return [pickle.loads(t.cpu().numpy().tobytes()[:s]) for s, t in zip(size_list, tensor_list)]


def all_gather_dict(data):
    """all_gather + concat (dim 0) of every tensor of a (possibly nested) dict."""
    world_size = get_world_size()
    if world_size == 1:
        return data
    out = {}
    for key, val in data.items():
        if isinstance(val, torch.Tensor):
            parts = [torch.empty_like(val) for _ in range(world_size)]
            dist.all_gather(parts, val)
            out[key] = torch.cat(parts, dim=0)
        elif isinstance(val, dict):
            out[key] = all_gather_dict(val)
    return out
