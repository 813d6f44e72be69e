"""SE(3) poses for the DRR module: `RigidTransform` and `convert` with the reference's semantics.

Mirrors the interface of reference diffdrr/pose.py:14-190 (RigidTransform.forward/compose/inverse, `convert`
from nine parameterisations).  This is host-side glue -- a handful of 3x3 products per step that autograd
differentiates -- not part of the CUDA hot path.  Conventions kept from the reference:
  * `RigidTransform(x)` applies  x -> M[:, :3] @ [x; 1]                    (pose.py:45-48)
  * `a.compose(b)` has matrix  b.matrix @ a.matrix                         (pose.py:69-71)
  * `convert(rot, xyz, ...)` builds  [R | R @ xyz]  (camera centre = R t)   (pose.py:149-157)
  * Euler angles follow the PyTorch3D convention: R = R_c0(a0) @ R_c1(a1) @ R_c2(a2)
"""
from __future__ import annotations

import torch

PARAMETERIZATIONS = [
    "axis_angle", "euler_angles", "matrix", "quaternion", "quaternion_adjugate", "rotation_6d", "rotation_9d",
    "rotation_10d", "se3_log_map",
]


class RigidTransform(torch.nn.Module):
    """A batch of 4x4 rigid (or affine) transforms acting on point clouds of shape (B, N, 3)."""

    def __new__(cls, matrix, eps=1e-6):
        if isinstance(matrix, cls):
            return matrix
        return super().__new__(cls)

    def __init__(self, matrix, eps: float = 1e-6):
        if isinstance(matrix, type(self)):
            return
        super().__init__()
        self.register_buffer("matrix", matrix if matrix.dim() == 3 else matrix.unsqueeze(0))
        self.eps = eps

    def __len__(self):
        return self.matrix.shape[0]

    def __getitem__(self, idx):
        return type(self)(self.matrix[idx])

    def __matmul__(self, other):
        return other.compose(self)

    @property
    def rotation(self):
        return self.matrix[..., :3, :3]

    @property
    def translation(self):
        return self.matrix[..., :3, 3]

    def forward(self, x):
        # x -> R x + t with broadcasting over the batch (B or 1) -- same result as the reference's padded einsum
        R, t = self.matrix[:, :3, :3], self.matrix[:, :3, 3]
        return x @ R.mT + t.unsqueeze(1)

    def _is_rigid(self):
        R = self.rotation
        eye = torch.eye(3, dtype=R.dtype, device=R.device)
        return bool(((R @ R.mT - eye).abs() < self.eps).all())

    def inverse(self):
        if self._is_rigid():
            Rt = self.rotation.mT
            return type(self)(make_matrix(Rt, -(Rt @ self.translation.unsqueeze(-1)).squeeze(-1)))
        return type(self)(torch.linalg.inv(self.matrix))

    def compose(self, other):
        return type(self)(other.matrix @ self.matrix)

    def convert(self, parameterization, convention=None, degrees=False):
        """Inverse of `convert`: (rotation parameters, translation) of this pose."""
        R = self.rotation
        translation = -self.inverse().translation
        if parameterization == "matrix":
            return R, translation
        if parameterization == "euler_angles":
            rot = matrix_to_euler_angles(R, convention)
            return (rot * (180.0 / torch.pi) if degrees else rot), translation
        if parameterization == "quaternion":
            return matrix_to_quaternion(R), translation
        if parameterization == "axis_angle":
            return quaternion_to_axis_angle(matrix_to_quaternion(R)), translation
        if parameterization == "rotation_6d":
            return R[..., :2, :].reshape(*R.shape[:-2], 6), translation
        if parameterization == "rotation_9d":
            return R.flatten(start_dim=1), translation
        if parameterization == "quaternion_adjugate":   # +q q^T, upper triangle (pose.py:250-253)
            return quaternion_to_quaternion_adjugate(matrix_to_quaternion(R)), translation
        if parameterization == "rotation_10d":          # -q q^T, upper triangle (pose.py:229-232)
            return quaternion_to_rotation_10d(matrix_to_quaternion(R)), translation
        if parameterization == "se3_log_map":           # pose.py:96-102: the log's own translation part replaces it
            params = self.get_se3_log()
            return params[..., 3:], params[..., :3]
        raise ValueError(f"Must be in {PARAMETERIZATIONS}, not {parameterization}")

    def get_se3_log(self):
        """(B, 6) = [log translation | log rotation] of this pose (pose.py:104-105)."""
        return se3_log_map(self.matrix.mT)


def make_matrix(R, t):
    if len(R) != len(t):
        raise AssertionError("rotation and translation batch sizes differ")
    top = torch.cat([R, t.unsqueeze(-1)], dim=-1)
    bottom = torch.zeros_like(top[:, :1, :])
    bottom[..., 3] = 1.0
    return torch.cat([top, bottom], dim=-2)


def convert(*args, parameterization, convention=None, degrees=False) -> RigidTransform:
    """Rotation parameters + translation (camera centre in the rotated frame) -> RigidTransform."""
    if parameterization == "matrix":
        return RigidTransform(args[0])
    if parameterization not in PARAMETERIZATIONS:
        raise ValueError(f"Must be in {PARAMETERIZATIONS}, not {parameterization}")
    rotation, translation = args
    if parameterization == "se3_log_map":
        return RigidTransform(se3_exp_map(torch.cat([translation, rotation], dim=-1)).mT)
    if parameterization == "euler_angles":
        if convention is None:
            raise ValueError("convention for Euler angles must be specified as a 3 letter combination of [X, Y, Z]")
        R = euler_angles_to_matrix(rotation * (torch.pi / 180.0) if degrees else rotation, convention)
    elif parameterization == "axis_angle":
        R = axis_angle_to_matrix(rotation)
    elif parameterization == "quaternion":
        R = quaternion_to_matrix(rotation)
    elif parameterization == "quaternion_adjugate":
        R = quaternion_to_matrix(quaternion_adjugate_to_quaternion(rotation))
    elif parameterization == "rotation_6d":
        R = rotation_6d_to_matrix(rotation)
    elif parameterization == "rotation_9d":
        R = rotation_9d_to_matrix(rotation)
    else:  # rotation_10d
        R = quaternion_to_matrix(rotation_10d_to_quaternion(rotation))
    centre = (R @ translation.unsqueeze(-1)).squeeze(-1)
    return RigidTransform(make_matrix(R, centre))


# ---- rotation parameterisations ------------------------------------------------------------------------
def _axis_rotation(axis: str, angle: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(angle), torch.sin(angle)
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    rows = {
        "X": (o, z, z, z, c, -s, z, s, c),
        "Y": (c, z, s, z, o, z, -s, z, c),
        "Z": (c, -s, z, s, c, z, z, z, o),
    }
    if axis not in rows:
        raise ValueError("letter must be either X, Y or Z.")
    return torch.stack(rows[axis], dim=-1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles: torch.Tensor, convention: str) -> torch.Tensor:
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3 or any(c not in "XYZ" for c in convention):
        raise ValueError(f"Invalid convention {convention}.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    r0, r1, r2 = (_axis_rotation(c, euler_angles[..., i]) for i, c in enumerate(convention))
    return r0 @ r1 @ r2


def matrix_to_euler_angles(matrix: torch.Tensor, convention: str) -> torch.Tensor:
    """Tait-Bryan / proper Euler angles of R = R_c0(a0) R_c1(a1) R_c2(a2)."""
    i0, i1, i2 = ("XYZ".index(c) for c in convention)
    tait_bryan = i0 != i2
    if tait_bryan:
        sign = -1.0 if (i0 - i2) in (-1, 2) else 1.0
        central = torch.asin(torch.clamp(matrix[..., i0, i2] * sign, -1.0, 1.0))
    else:
        central = torch.acos(torch.clamp(matrix[..., i0, i0], -1.0, 1.0))

    def angle_from_tan(axis, other, data, horizontal):
        j1, j2 = {0: (2, 1), 1: (0, 2), 2: (1, 0)}[axis]
        if horizontal:
            j2, j1 = j1, j2
        even = (axis, other) in ((0, 1), (1, 2), (2, 0))
        if horizontal == even:
            return torch.atan2(data[..., j1], data[..., j2])
        if tait_bryan:
            return torch.atan2(-data[..., j2], data[..., j1])
        return torch.atan2(data[..., j2], -data[..., j1])

    first = angle_from_tan(i0, i1, matrix[..., i2], False)
    last = angle_from_tan(i2, i1, matrix[..., i0, :], True)
    return torch.stack((first, central, last), dim=-1)


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Real-first quaternion (not necessarily unit) -> rotation matrix."""
    w, x, y, z = torch.unbind(q, -1)
    k = 2.0 / (q * q).sum(-1)
    m = torch.stack((
        1 - k * (y * y + z * z), k * (x * y - z * w), k * (x * z + y * w),
        k * (x * y + z * w), 1 - k * (x * x + z * z), k * (y * z - x * w),
        k * (x * z - y * w), k * (y * z + x * w), 1 - k * (x * x + y * y),
    ), dim=-1)
    return m.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> real-first unit quaternion (largest-component branch for stability)."""
    m = matrix
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    q_abs = torch.sqrt(torch.clamp(torch.stack([
        1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1), min=0.0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]], -1),
        torch.stack([m[..., 2, 1] - m[..., 1, 2], q_abs[..., 1] ** 2, m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0]], -1),
        torch.stack([m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] + m[..., 0, 1], q_abs[..., 2] ** 2, m[..., 2, 1] + m[..., 1, 2]], -1),
        torch.stack([m[..., 1, 0] - m[..., 0, 1], m[..., 2, 0] + m[..., 0, 2], m[..., 2, 1] + m[..., 1, 2], q_abs[..., 3] ** 2], -1),
    ], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(dim=-1)
    return standardize_quaternion(torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2))


def standardize_quaternion(q: torch.Tensor) -> torch.Tensor:
    return torch.where(q[..., :1] < 0, -q, q)


def axis_angle_to_quaternion(axis_angle: torch.Tensor) -> torch.Tensor:
    angle = torch.linalg.norm(axis_angle, dim=-1, keepdim=True)
    half = 0.5 * angle
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    k = torch.where(small, 0.5 - angle * angle / 48.0, torch.sin(half) / safe)
    return torch.cat([torch.cos(half), axis_angle * k], dim=-1)


def axis_angle_to_matrix(axis_angle: torch.Tensor) -> torch.Tensor:
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def quaternion_to_axis_angle(q: torch.Tensor) -> torch.Tensor:
    n = torch.linalg.norm(q[..., 1:], dim=-1, keepdim=True)
    angle = 2.0 * torch.atan2(n, q[..., :1])
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    k = torch.where(small, 0.5 - angle * angle / 48.0, torch.sin(0.5 * angle) / safe)
    return q[..., 1:] / k


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """Gram-Schmidt on two 3-vectors (Zhou et al., CVPR 2019); rows of the result are the orthonormal frame."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def rotation_9d_to_matrix(rotation: torch.Tensor) -> torch.Tensor:
    """Nearest rotation (in Frobenius norm) to an arbitrary 3x3 via SVD (Levinson et al., NeurIPS 2020)."""
    u, _, vt = torch.linalg.svd(rotation.reshape(-1, 3, 3))
    det = torch.linalg.det(u @ vt)
    fix = torch.ones_like(vt[..., 0])
    fix[..., 2] = det
    return u @ (fix.unsqueeze(-1) * vt)


def _sym4(vec: torch.Tensor) -> torch.Tensor:
    i, j = torch.triu_indices(4, 4)
    A = vec.new_zeros(len(vec), 4, 4)
    A[:, i, j] = vec
    A[:, j, i] = vec
    return A


def rotation_10d_to_quaternion(rotation: torch.Tensor) -> torch.Tensor:
    """Eigenvector of the smallest eigenvalue of the symmetric 4x4 built from 10 numbers (Peretroukhin et al. 2020)."""
    return torch.linalg.eigh(_sym4(rotation)).eigenvectors[..., 0]


def quaternion_adjugate_to_quaternion(rotation: torch.Tensor) -> torch.Tensor:
    """Dominant column of the quaternion-adjugate matrix q q^T (Lin et al. 2022), up to scale."""
    A = _sym4(rotation)
    col_norm = A.norm(dim=1)
    k = col_norm.argmax(dim=1)
    return A[torch.arange(len(A)), k] / col_norm.amax(dim=1, keepdim=True)


# ---- se(3) exponential (PyTorch3D row-vector convention: the reference transposes the result) ----------
def _hat(v: torch.Tensor) -> torch.Tensor:
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack((o, -z, y, z, o, -x, -y, x, o), dim=-1).reshape(v.shape[:-1] + (3, 3))


def _upper_triangle_4x4(A: torch.Tensor) -> torch.Tensor:
    idx, jdx = torch.triu_indices(4, 4)
    return A[..., idx, jdx]


def quaternion_to_quaternion_adjugate(q: torch.Tensor) -> torch.Tensor:
    """Unit quaternion -> the 10 upper-triangle entries of q q^T (inverse of quaternion_adjugate_to_quaternion)."""
    return _upper_triangle_4x4(q.unsqueeze(-1) * q.unsqueeze(-2))


def quaternion_to_rotation_10d(q: torch.Tensor) -> torch.Tensor:
    """Unit quaternion -> the 10 upper-triangle entries of -q q^T (inverse of rotation_10d_to_quaternion)."""
    return _upper_triangle_4x4(-(q.unsqueeze(-1) * q.unsqueeze(-2)))


def so3_log_map(R: torch.Tensor) -> torch.Tensor:
    """Rotation matrices (B, 3, 3) -> axis-angle logarithms (B, 3)."""
    return quaternion_to_axis_angle(matrix_to_quaternion(R))


def _se3_V(w: torch.Tensor, eps: float):
    theta2 = (w * w).sum(-1).clamp(min=eps * eps)
    theta = theta2.sqrt()
    K = _hat(w)
    K2 = K @ K
    b = ((1 - torch.cos(theta)) / theta2)[..., None, None]
    c = ((theta - torch.sin(theta)) / (theta2 * theta))[..., None, None]
    eye = torch.eye(3, dtype=w.dtype, device=w.device).expand_as(K)
    return eye + b * K + c * K2


def se3_log_map(transform: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """(B, 4, 4) with rows [R^T 0; T 1] (the transpose convention of se3_exp_map) -> (B, 6) = [V^-1 T | log R]."""
    if transform.dim() != 3 or transform.shape[-2:] != (4, 4):
        raise ValueError("Input tensor shape has to be (N, 4, 4).")
    if not torch.allclose(transform[:, :3, 3], torch.zeros_like(transform[:, :3, 3])):
        raise ValueError("All elements of `transform[:, :3, 3]` should be 0.")
    w = so3_log_map(transform[:, :3, :3].mT)
    v = torch.linalg.solve(_se3_V(w, eps), transform[:, 3, :3].unsqueeze(-1)).squeeze(-1)
    return torch.cat([v, w], dim=-1)


def se3_exp_map(log_transform: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """(B,6) = [translation-part v | rotation-part w]  ->  (B,4,4) with rows [R^T 0; (V v)^T 1]."""
    v, w = log_transform[..., :3], log_transform[..., 3:]
    theta2 = (w * w).sum(-1).clamp(min=eps * eps)
    theta = theta2.sqrt()
    K = _hat(w)
    K2 = K @ K
    a = (torch.sin(theta) / theta)[..., None, None]
    b = ((1 - torch.cos(theta)) / theta2)[..., None, None]
    c = ((theta - torch.sin(theta)) / (theta2 * theta))[..., None, None]
    eye = torch.eye(3, dtype=w.dtype, device=w.device).expand_as(K)
    R = eye + a * K + b * K2
    V = eye + b * K + c * K2
    T = torch.zeros(*w.shape[:-1], 4, 4, dtype=w.dtype, device=w.device)
    T[..., :3, :3] = R.mT
    T[..., 3, :3] = (V @ v.unsqueeze(-1)).squeeze(-1)
    T[..., 3, 3] = 1.0
    return T
