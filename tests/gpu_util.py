import numpy as np
import pytest


def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def relrms(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))
