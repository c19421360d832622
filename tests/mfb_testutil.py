"""Shared helpers for the test-suite (imported as a top-level module; tests/ has no __init__)."""
import numpy as np
import torch


def t32(a):
    return torch.from_numpy(np.asarray(a)).to(torch.float32)


def cuda(a):
    return t32(a).cuda()
