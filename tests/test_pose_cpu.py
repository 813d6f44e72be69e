"""Host-side pose algebra (diffdrr_b200/pose.py) against the UNMODIFIED reference (diffdrr/pose.py), when it is available
(build container; /root/reference or baseline/_ref), plus round trips that hold anywhere: all nine parameterisations in
BOTH directions (se3_log_map / quaternion_adjugate / rotation_10d were missing from RigidTransform.convert)."""
import os
import sys

import pytest
import torch

from conftest import ROOT
from diffdrr_b200.pose import PARAMETERIZATIONS, RigidTransform, convert


def _poses(B=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    rot = (torch.rand(B, 3, generator=g) * 2 - 1) * 1.2
    xyz = (torch.rand(B, 3, generator=g) * 2 - 1) * 200.0
    return convert(rot, xyz, parameterization="euler_angles", convention="ZXY")


@pytest.mark.parametrize("param", [p for p in PARAMETERIZATIONS])
def test_convert_round_trip(param):
    pose = _poses()
    kw = dict(convention="ZYX") if param == "euler_angles" else {}
    back = pose.convert(param, **kw)
    again = convert(*back, parameterization=param, **kw) if param != "matrix" else RigidTransform(pose.matrix)
    assert torch.allclose(again.matrix, pose.matrix, atol=2e-4, rtol=1e-5), param


def _reference_pose_module():
    for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(base, "diffdrr")):
            for p in (os.path.join(ROOT, "tests", "_refshim"), base):
                if p not in sys.path:
                    sys.path.insert(0, p)
            try:
                import diffdrr.pose as ref
                return ref
            except Exception:
                return None
    return None


@pytest.mark.parametrize("param", [p for p in PARAMETERIZATIONS if p != "matrix"])
def test_convert_matches_the_reference_both_ways(param):
    ref = _reference_pose_module()
    if ref is None:
        pytest.skip("reference sources not available on this box")
    pose = _poses(seed=3)
    kw = dict(convention="ZXY") if param == "euler_angles" else {}
    ours = pose.convert(param, **kw)
    theirs = ref.RigidTransform(pose.matrix.clone()).convert(param, **kw)
    for a, b in zip(ours, theirs):
        assert torch.allclose(a, b, atol=5e-4, rtol=1e-4), param
    fwd_ours = convert(*theirs, parameterization=param, **kw).matrix
    fwd_ref = ref.convert(*theirs, parameterization=param, **kw).matrix
    assert torch.allclose(fwd_ours, fwd_ref, atol=5e-4, rtol=1e-4), param
    assert torch.allclose(pose.get_se3_log(), ref.RigidTransform(pose.matrix.clone()).get_se3_log(), atol=5e-4, rtol=1e-4)
