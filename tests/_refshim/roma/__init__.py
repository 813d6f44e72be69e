"""Stand-in for roma.is_orthonormal_matrix (test infrastructure only); pose.py:11,59."""
import torch


def is_orthonormal_matrix(R, epsilon=1e-7):
    eye = torch.eye(R.shape[-1], dtype=R.dtype, device=R.device)
    return bool(torch.all(torch.abs(R @ R.mT - eye) < epsilon))
