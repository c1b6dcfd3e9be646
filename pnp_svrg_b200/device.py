"""Small helpers around torch device memory (torch is plumbing here: allocations and streams)."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    _lib.init_device()
    return torch.device('cuda', torch.cuda.current_device())


def pin_to_local_rank(local_rank=None, local_world=None):
    """One process per GPU on a shared host: give every local rank its own slice of the cores this process may run on
    (``sched_setaffinity``), so that the ranks' host-side sampler threads (``pnp_host_draws_*``) and Python loops do not
    fight over the same cores -- the worker-thread count of the look-ahead draw queue follows the affinity mask.
    Arguments default to torchrun's LOCAL_RANK / LOCAL_WORLD_SIZE.  Returns the cores kept (None: nothing changed)."""
    import os
    if not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        lr = int(os.environ.get('LOCAL_RANK', '0')) if local_rank is None else int(local_rank)
        lw = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1'))) if local_world is None else int(local_world)
        cores = sorted(os.sched_getaffinity(0))
        if lw <= 1 or len(cores) < 2 * lw:
            return None
        share = len(cores) // lw
        mine = cores[lr * share:(lr + 1) * share]
        os.sched_setaffinity(0, mine)
        return mine
    except (OSError, ValueError):
        return None


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


_PINNED = {}


def pinned_buffer(key, shape, dtype):
    """A reusable pinned host buffer (allocating page-locked memory costs milliseconds; the loops upload Xinit and
    download z on every call).  One buffer per key: callers use it synchronously, between a copy and its wait."""
    want = (tuple(shape), dtype)
    hit = _PINNED.get(key)
    if hit is None or hit[0] != want:
        hit = (want, torch.empty(shape, dtype=dtype).pin_memory())
        _PINNED[key] = hit
    return hit[1]


def to_lines(z, H, W, device=None):
    """host/device (N,) or (H, W) array -> float32 device tensor in the transposed line layout
    [W][H] (line c = original column c)."""
    if isinstance(z, torch.Tensor):
        t = z.detach().to(device=device or 'cuda', dtype=torch.float32).reshape(H, W)
        return t.t().contiguous()
    # cast on the host with torch (multi-threaded, contiguous), transpose on the device: a strided NumPy transpose of a
    # 2048^2 float64 image costs 20-50 ms of host time, this path well under a millisecond plus the copy
    a = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(H, W))
    if not a.flags.writeable:
        a = a.copy()                                   # torch.from_numpy wants a writable buffer
    dev = torch.device(device or 'cuda')
    if dev.type != 'cuda':
        return torch.from_numpy(a).to(torch.float32).to(dev).t().contiguous()
    # float64 -> float32 straight into a pinned staging buffer (one pass on the host), DMA from there
    stage = pinned_buffer(('up', H * W), (H, W), torch.float32)
    stage.copy_(torch.from_numpy(a))
    t = stage.to(dev, non_blocking=True).t().contiguous()
    torch.cuda.current_stream(dev).synchronize()       # complete on return (any stream may use it; the staging buffer is free again)
    return t


def from_lines(t, H, W):
    """device [W][H] float32 -> host float64 (N,) in the reference's raveled (row-major) order."""
    d = t.reshape(W, H).t().contiguous()
    if not d.is_cuda:
        return d.to(torch.float64).numpy().ravel()
    stage = pinned_buffer(('down', H * W), (H, W), torch.float32)
    stage.copy_(d, non_blocking=True)
    torch.cuda.current_stream(d.device).synchronize()
    return stage.to(torch.float64).numpy().ravel()


def lines_to_image_tensor(t, H, W):
    """device [W][H] -> device (H, W) view-contiguous tensor"""
    return t.reshape(W, H).t().contiguous()
