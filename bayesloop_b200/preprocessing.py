"""Data segmentation (reference: bayesloop/preprocessing.py:14-26)."""
import numpy as np


def movingWindow(rawData, n):
    """Overlapping windows of `n` consecutive rows: result[i] = rawData[i:i+n]."""
    raw = np.asarray(rawData)
    count = raw.shape[0] - (n - 1)
    if count <= 0:
        return np.empty((0, n) + raw.shape[1:], dtype=raw.dtype)
    idx = np.arange(count)[:, None] + np.arange(n)[None, :]
    return raw[idx]
