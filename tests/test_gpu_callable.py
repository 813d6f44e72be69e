"""Callable `reducefn` (reference renderers.py:175-183): the un-reduced kernels of csrc/literal.cu hand the callable the same
(B, N, M-1) segment / (B, N, n_points) sample tensor the reference builds, and their backward takes its per-segment gradient.
Goldens: tests/golden/make_golden_callable.py (unmodified reference, fp32 and fp64, images + autograd)."""
import numpy as np
import pytest
import torch

from callable_reducers import REDUCERS
from conftest import load_golden, relerr
from gpu_common import DEV

pytestmark = pytest.mark.gpu

CASES = [("siddon", "siddon_nc_b4_callable", {}), ("trilinear", "trilinear_nc_b4_callable", dict(n_points=160))]


def _t(a, dt, grad=False):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(DEV).requires_grad_(grad)


@pytest.mark.parametrize("kind,name,fkw", CASES)
@pytest.mark.parametrize("rname", sorted(REDUCERS))
@pytest.mark.parametrize("dt_tag,dt,tol_img,tol_grad", [("f32", torch.float32, 1e-4, 1e-3), ("f64", torch.float64, 1e-9, 1e-7)])
def test_callable_reducefn_matches_the_reference(kind, name, fkw, rname, dt_tag, dt, tol_img, tol_grad):
    from diffdrr_b200 import Siddon, Trilinear
    g = load_golden(name)
    mod = (Siddon if kind == "siddon" else Trilinear)(reducefn=REDUCERS[rname])
    v, s, tg, l = (_t(g[k], dt, True) for k in ("volume", "source", "target", "raylen"))
    img = mod(v, s, tg, l, **fkw)
    assert img.dtype == dt and img.shape == g[f"{rname}_img_f64"].shape
    assert relerr(img.detach().cpu().numpy(), g[f"{rname}_img_f64"]) < tol_img
    (img * _t(g["w"], dt)).sum().backward()
    for key, x in (("g_volume", v), ("g_source", s), ("g_target", tg), ("g_raylen", l)):
        assert relerr(x.grad.cpu().numpy(), g[f"{rname}_{key}_f64"]) < tol_grad, key


def test_segment_tensor_reduces_to_the_fused_kernels():
    """sum / max of the un-reduced tensor == the fused kernels' images (same rays), for both renderers."""
    from diffdrr_b200 import Siddon, Trilinear
    g = load_golden("siddon_nc_b4")
    args = [_t(g[k], torch.float32) for k in ("volume", "source", "target", "raylen")]
    for cls, fkw in ((Siddon, {}), (Trilinear, dict(n_points=96))):
        for name, fn in (("sum", lambda x: x.sum(-1)), ("max", lambda x: x.max(-1).values)):
            a = cls(reducefn=fn)(*args, **fkw)
            b = cls(reducefn=name)(*args, **fkw)
            assert relerr(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, (cls.__name__, name)
