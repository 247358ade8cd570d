"""``split_dataset`` -- the reference's multi-GPU partitioner (misc/torchutils.py:66-68)."""
import numpy as np
from torch.utils.data import Subset


def split_indices(n_items, n_splits):
    """rank r gets items r, r+n, r+2n, ... (so per-rank file sets match the reference's)."""
    return [np.arange(i, n_items, n_splits) for i in range(n_splits)]


def split_dataset(dataset, n_splits):
    return [Subset(dataset, idx) for idx in split_indices(len(dataset), n_splits)]
