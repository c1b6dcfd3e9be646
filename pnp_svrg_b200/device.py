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
    a = np.asarray(z, dtype=np.float64).reshape(H, W)
    t = torch.from_numpy(np.ascontiguousarray(a.T, dtype=np.float32))
    return t.to(device or 'cuda', non_blocking=False)


def from_lines(t, H, W):
    """device [W][H] float32 -> host float64 (N,) in the reference's raveled (row-major) order."""
    return t.reshape(W, H).t().contiguous().cpu().numpy().astype(np.float64).ravel()


def lines_to_image_tensor(t, H, W):
    """device [W][H] -> device (H, W) view-contiguous tensor"""
    return t.reshape(W, H).t().contiguous()
