"""Small helpers around torch device memory (torch is plumbing here: allocations and streams)."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    _lib.init_device()
    return torch.device('cuda', torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


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
    t = torch.from_numpy(a).to(torch.float32).to(device or 'cuda', non_blocking=False).t().contiguous()
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()      # complete on return, like the copy itself (any stream may use it)
    return t


def from_lines(t, H, W):
    """device [W][H] float32 -> host float64 (N,) in the reference's raveled (row-major) order."""
    return t.reshape(W, H).t().contiguous().cpu().to(torch.float64).numpy().ravel()


def lines_to_image_tensor(t, H, W):
    """device [W][H] -> device (H, W) view-contiguous tensor"""
    return t.reshape(W, H).t().contiguous()
