"""Helpers shared by the `-m gpu` parity tests (all of them call the CUDA kernels through the C ABI)."""
import numpy as np
import torch

from conftest import relerr  # noqa: F401

DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def t(a, grad=False):
    x = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(DEV)
    return x.requires_grad_(grad)


def grad_tol(g, key, floor=1e-4):
    """compare with the fp64 reference, allow max(floor, 2 x the reference's own fp32-vs-fp64 error)."""
    return max(floor, 2.0 * relerr(g[key + "_f32"], g[key + "_f64"]))
