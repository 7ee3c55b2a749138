"""Hot-path helpers of the reference's ``Utils/utils.py`` (indexing only, bit exact)."""
import torch


def set_gpu(mode, verbose=False):
    """Utils/utils.py:9-16.  This package is CUDA-only: the device is always the current CUDA device."""
    return torch.device("cuda" if (mode and torch.cuda.is_available()) else "cpu")


def batch_reduce(x, reduce=torch.sum, batch_dim=0):
    """Utils/utils.py:25-28: reduce everything but the batch dimension."""
    batch_size = x.size(batch_dim)
    return reduce(x.view(batch_size, -1), dim=-1)


def split_feature(tensor, type="split"):
    """Utils/utils.py:86-91: 'split' = channel halves, 'cross' = even / odd channels (views)."""
    C = tensor.size(1)
    if type == "split":
        return tensor[:, :C // 2, ...], tensor[:, C // 2:, ...]
    elif type == "cross":
        return tensor[:, 0::2, ...], tensor[:, 1::2, ...]
    raise ValueError(f"unknown split type {type!r}")
