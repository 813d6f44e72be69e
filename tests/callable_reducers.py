"""Callable `reducefn`s shared by tests/golden/make_golden_callable.py (run on the unmodified reference) and the GPU tests."""
import torch


def ramp_weighted(img):
    """Order-sensitive: weights the j-th segment / sample by a ramp 0.5 .. 1.5 along the ray."""
    w = torch.linspace(0.5, 1.5, img.shape[-1], dtype=img.dtype, device=img.device)
    return (img * w).sum(dim=-1)


def root_sum_squares(img):
    return (img.square().sum(dim=-1) + 1e-6).sqrt()


REDUCERS = {"ramp": ramp_weighted, "rss": root_sum_squares}
