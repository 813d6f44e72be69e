"""Real-data fixture (SURVEY.md 7.3 / VERDICT r1 items 7, 9): a crop of the reference's example CT, HU -> density with the
reference's own rule (data.py:214-227), rendered by the unmodified reference renderers (tests/golden/make_golden_realdata.py).
CPU: our HU -> density map, the oracle and the device math (host emulation); GPU: the kernels through the modules."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, relerr


@pytest.fixture(scope="module")
def g():
    rec = dict(np.load(os.path.join(GOLDEN, "realdata_cxr_crop.npz")))
    from diffdrr_b200.data import transform_hu_to_density
    rec["density"] = transform_hu_to_density(torch.from_numpy(rec["hu_crop"].astype(np.float32)),
                                             float(rec["bone_attenuation_multiplier"])).numpy()
    return rec


def test_hu_to_density_matches_the_reference_rule(g):
    assert g["density"].dtype == np.float32 and g["density"].min() == 0.0 and g["density"].max() == 1.0
    assert np.array_equal(g["density"][::3, ::3, ::3], g["density_sub3"])
    assert abs(g["density"].astype(np.float64).sum() - float(g["density_sum"])) < 1e-6 * float(g["density_sum"])


def test_nifti_reader_round_trip(tmp_path):
    """read_nifti on a hand-written NIfTI-1 file (int16, scl_inter, sform), gzipped."""
    import gzip
    import struct

    from diffdrr_b200.data import read_nifti
    vol = (np.arange(4 * 5 * 6, dtype=np.int16).reshape(6, 5, 4) - 50)            # file order: i fastest
    hdr = bytearray(352)
    hdr[0:4] = struct.pack("<i", 348)
    hdr[40:56] = struct.pack("<8h", 3, 4, 5, 6, 1, 1, 1, 1)
    hdr[70:72] = struct.pack("<h", 4)
    hdr[72:74] = struct.pack("<h", 16)
    hdr[108:112] = struct.pack("<f", 352.0)
    hdr[112:120] = struct.pack("<2f", 1.0, -1024.0)
    hdr[254:256] = struct.pack("<h", 1)
    hdr[280:328] = struct.pack("<12f", -0.7, 0, 0, 166.0, 0, 0.7, 0, -187.6, 0, 0, 2.5, -340.0)
    hdr[344:348] = b"n+1\0"
    path = tmp_path / "t.nii.gz"
    with gzip.open(path, "wb") as f:
        f.write(bytes(hdr) + vol.tobytes())
    arr, aff = read_nifti(str(path))
    assert arr.shape == (4, 5, 6) and arr[1, 2, 3] == vol[3, 2, 1] - 1024.0
    assert np.allclose(aff[:3, 3], [166.0, -187.6, -340.0]) and np.allclose(np.diag(aff)[:3], [-0.7, 0.7, 2.5])


def test_oracle_and_device_math_on_real_data(g):
    from hostemu import emu
    from oracle import oracle
    args = (g["density"], g["source"], g["target"], g["raylen"])
    assert relerr(oracle.siddon_fwd(*args, dtype=np.float64), g["siddon_f64"]) < 1e-9
    assert relerr(oracle.siddon_fwd(*args, dtype=np.float32), g["siddon_f32"]) < 2e-5
    assert relerr(emu.siddon_fwd(*args), g["siddon_f64"]) < 1e-4
    amin, amax = oracle.alpha_minmax(g["density"].shape, g["source"], g["target"], 0.5, 1e-8, np.float32)
    tri = oracle.trilinear_fwd(*args, n_points=300, alphamin=amin, alphamax=amax, dtype=np.float64)
    assert relerr(tri, g["trilinear_f64"]) < 1e-5


@pytest.mark.gpu
def test_kernels_on_real_data(g):
    from diffdrr_b200 import Siddon, Trilinear
    from gpu_common import t
    vol, src, tgt, img = t(g["density"]), t(g["source"]), t(g["target"]), t(g["raylen"])
    assert relerr(Siddon()(vol, src, tgt, img).cpu().numpy(), g["siddon_f64"]) < 1e-4
    assert relerr(Trilinear()(vol, src, tgt, img, n_points=300).cpu().numpy(), g["trilinear_f64"]) < 1e-4
