"""CPU-side checks of the drop-in boundary: the C ABI library loads and exports every symbol include/b200drr.h
declares, argument validation works without a GPU, and the host-side modules mirror the reference's interface."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from conftest import ROOT
from diffdrr_b200 import DRR, Siddon, Trilinear, _lib, synthetic
from diffdrr_b200 import build as b200build


@pytest.fixture(scope="module")
def lib():
    b200build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "b200drr.h")).read()
    declared = sorted(set(re.findall(r"\b(b200drr_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in include/b200drr.h"
    raw = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported by libb200drr.so"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding and header disagree"
    assert lib.b200drr_version() == 200  # 0.2.0: + b200drr_ncc_* (round 2)


def test_argument_validation_without_gpu(lib):
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.b200drr_siddon_fwd(None, 4, 4, 4, None, None, None, None, 1, 1, 0.5, 1e-8, 0, 0, None) == -1
    assert lib.b200drr_trilinear_fwd(None, 4, 4, 4, None, None, None, None, 1, 1, 0.5, 1e-8, 1, None, 0, 0, None) == -1
    assert lib.b200drr_siddon_visits(0, 4, 4, None, None, None, 1, 1, 0.5, 1e-8, None) == -1
    # every entry point taking pointers rejects an all-NULL call with B200DRR_EINVAL instead of touching CUDA
    n = None
    assert lib.b200drr_siddon_fwd_sens(n, 4, 4, 4, n, n, n, n, n, 1, 1, 0.5, 1e-8, n) == -1
    assert lib.b200drr_siddon_fwd_sens_grid(n, 4, 4, 4, n, n, n, n, n, 1, 2, 2, 0.5, 1e-8, 0, n) == -1
    assert lib.b200drr_siddon_fwd_sens_pose(n, 4, 4, 4, n, n, n, n, n, n, n, 1, 2, 2, 0.5, 1e-8, n) == -1
    assert lib.b200drr_siddon_bwd_sens(n, n, n, n, n, 1, 4, 0, n) == -1
    assert lib.b200drr_siddon_bwd_sens_pose(n, n, n, n, n, n, n, n, 1, 2, 2, 0, n) == -1
    assert lib.b200drr_trilinear_fwd_sens(n, 4, 4, 4, n, n, n, n, n, 1, 4, 0, 0, 0.5, 1e-8, 10, n, 0, n) == -1
    assert lib.b200drr_trilinear_fwd_sens_packed(n, 4, 4, 4, n, n, n, n, n, 1, 2, 2, 0.5, 1e-8, 10, n, 0, n) == -1
    assert lib.b200drr_trilinear_bwd_sens(n, n, n, n, n, n, 1, 4, n) == -1
    assert lib.b200drr_euler_pose_fwd(n, n, 2, 0, 1, 1.0, n, 1, n) == -1
    assert lib.b200drr_euler_pose_bwd(n, n, 2, 0, 1, 1.0, n, n, n, 1, n) == -1
    assert lib.b200drr_pose_rays_fwd(n, n, n, n, n, n, n, 1, n) == -1
    assert lib.b200drr_pose_rays_bwd(n, n, n, n, n, n, n, 1, n) == -1
    # bad Euler axis triple (middle axis repeated / out of range) is refused even with valid-looking pointers
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.b200drr_euler_pose_fwd(p, p, 0, 0, 1, 1.0, p, 1, n) == -1
    assert lib.b200drr_euler_pose_fwd(p, p, 0, 3, 1, 1.0, p, 1, n) == -1
    assert b"invalid argument" in lib.b200drr_error_string(-1)
    assert b"unsupported" in lib.b200drr_error_string(-2)
    assert lib.b200drr_error_string(0) == b"success"


def test_no_cpu_fallback():
    vol = torch.rand(8, 8, 8)
    src = torch.zeros(1, 1, 3)
    tgt = torch.rand(1, 5, 3)
    img = torch.ones(1, 1, 5)
    with pytest.raises(_lib.B200DRRError, match="no CPU fallback"):
        Siddon()(vol, src, tgt, img)
    with pytest.raises(_lib.B200DRRError, match="no CPU fallback"):
        Trilinear()(vol, src, tgt, img, alphamin=0.0, alphamax=1.0)


def test_renderer_signatures_match_reference():
    # reference renderers.py:14-22,34-42 and 189-195,205-216
    assert list(inspect.signature(Siddon.__init__).parameters)[1:] == [
        "voxel_shift", "mode", "stop_gradients_through_grid_sample", "filter_intersections_outside_volume", "reducefn", "eps"]
    assert list(inspect.signature(Siddon.forward).parameters)[1:] == [
        "volume", "source", "target", "img", "align_corners", "mask"]
    assert list(inspect.signature(Trilinear.__init__).parameters)[1:] == ["voxel_shift", "mode", "reducefn", "eps"]
    assert list(inspect.signature(Trilinear.forward).parameters)[1:] == [
        "volume", "source", "target", "img", "n_points", "align_corners", "mask", "alphamin", "alphamax"]
    assert list(inspect.signature(DRR.__init__).parameters)[1:] == [
        "subject", "sdd", "height", "delx", "width", "dely", "x0", "y0", "p_subsample", "reshape", "reverse_x_axis",
        "patch_size", "renderer", "voxel_shift", "persistent", "compile_renderer", "checkpoint_gradients", "renderer_kwargs"]


def test_drr_module_surface_on_cpu():
    vol = synthetic.make_volume(16, "phantom")
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(12), renderer="trilinear")
    assert isinstance(drr.renderer, Trilinear)
    assert set(dict(drr.named_buffers())) >= {"_affine", "_affine_inverse", "density", "detector.source", "detector.target"}
    assert drr.density.shape == (16, 16, 16) and drr.device.type == "cpu" and drr.dtype == torch.float32
    with pytest.raises(ValueError, match="renderer must be"):
        DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(12), renderer="nope")
    with pytest.raises(ValueError):
        Siddon(reducefn="median")(vol, vol, vol, vol) if False else __import__("diffdrr_b200.renderers", fromlist=["x"])._reduce_code("median")
    # intrinsics editing keeps the x0/y0 sign quirk of the reference (Q9)
    drr2 = DRR(synthetic.make_subject(vol), sdd=1000.0, height=10, delx=2.0, x0=3.0, y0=-4.0)
    assert drr2.detector.x0 == -3.0 and drr2.detector.y0 == 4.0
    drr2.set_intrinsics_(delx=1.0)
    assert drr2.detector.x0 == -3.0 and drr2.detector.delx == 1.0
    drr2.rescale_detector_(0.5)
    assert drr2.detector.height == 5 and drr2.detector.delx == 2.0
    # projection helpers round-trip
    from diffdrr_b200.pose import convert
    rot, xyz = synthetic.make_poses(2)
    pose = convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    px = torch.tensor([[[2.0, 3.0], [7.0, 1.0]], [[4.0, 4.0], [0.5, 9.0]]])
    world = drr2.inverse_projection(pose, px.clone())
    back = drr2.perspective_projection(pose, world)
    assert torch.allclose(back, px, atol=1e-3)


def test_render_grid_shape_is_validated_on_cpu():
    """`DRR.render(grid_shape=...)` (row-block hint used by ray sharding) refuses anything but whole detector rows."""
    vol = synthetic.make_volume(16, "phantom")
    drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(12))
    src, tgt = torch.zeros(1, 1, 3), torch.zeros(1, 24, 3)
    with pytest.raises(ValueError, match="grid_shape"):
        drr.render(drr.density, src, tgt, grid_shape=(2, 11))      # wrong width
    with pytest.raises(ValueError, match="grid_shape"):
        drr.render(drr.density, src, tgt, grid_shape=(3, 12))      # 36 rays declared, 24 given
    sub = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(12), p_subsample=0.5)
    with pytest.raises(ValueError, match="grid_shape"):
        sub.render(sub.density, src, tgt, grid_shape=(2, 12))      # sub-sampled detector has no row blocks
    from diffdrr_b200.parallel import render_sharded
    assert render_sharded.__kwdefaults__["shard"] == "auto"


def test_compile_renderer_flag_is_not_silently_ignored():
    """compile_renderer=True (reference drr.py:102-103) warns and marks the renderer opaque to dynamo instead of no-op'ing."""
    import warnings

    vol = synthetic.make_volume(8, "smooth")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        drr = DRR(synthetic.make_subject(vol), **synthetic.detector_kwargs(6), compile_renderer=True)
    assert drr.compile_renderer and any("compile_renderer" in str(x.message) for x in w)


def test_ctypes_signatures_have_the_arity_of_the_header_prototypes():
    """Every ctypes argtypes list has exactly as many entries as the C prototype in include/b200drr.h has parameters (a
    mismatch only shows up as a TypeError on the GPU box otherwise)."""
    header = open(os.path.join(ROOT, "include", "b200drr.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char \*)\s*(b200drr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 40
    for name, args in protos:
        n = 0 if args.strip() in ("void", "") else len(args.split(","))
        assert len(_lib._SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters, ctypes binding {len(_lib._SIGNATURES[name][1])}"
