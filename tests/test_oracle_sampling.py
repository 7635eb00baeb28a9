"""CPU: the torch restatement of the reference's sampling / activation math
(oracle/torch_ref.py) against golden vectors produced by the reference's own
Python (tests/golden/make_sampling_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sampling_*.npz")))


def load(path):
    z = np.load(path)
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def run_oracle(d):
    cp = d["curve_points"].clone().requires_grad_(True)
    width = d["width"].clone().requires_grad_(True)
    opl = d["opacity_logit"].clone().requires_grad_(True)
    mask = d["mask_logit"].clone().requires_grad_(True)
    n = int(d["n"])
    xyz, rot, scal = torch_ref.sample_curves(cp, width, d["is_bezier"], n)
    _, opacity, scales, rot_n, colors, all_map = torch_ref.raster_inputs(
        xyz, rot, scal, opl, n, mask, d["cam_center"], d["world_view"], use_mask=bool(d["use_mask"]))
    return (cp, width, opl, mask), (xyz, rot, scal, opacity, rot_n, scales, all_map[:, :3])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_forward_matches_reference_python(path):
    d = load(path)
    _, outs = run_oracle(d)
    names = ["xyz", "rotation", "scaling", "opacity", "rot_normalized", "scales_masked", "local_axis"]
    for name, o in zip(names, outs):
        torch.testing.assert_close(o, d[name], rtol=1e-6, atol=1e-7, msg=name)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_gradients_match_reference_autograd(path):
    d = load(path)
    leaves, outs = run_oracle(d)
    ws = [d[k] for k in ("w_xyz", "w_rot", "w_scal", "w_opacity", "w_rot_n", "w_scales", "w_local")]
    loss = sum((a * b).sum() for a, b in zip(outs, ws))
    loss.backward()
    for leaf, key in zip(leaves, ("g_curve_points", "g_width", "g_opacity", "g_mask")):
        g = leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)
        ref = d[key]
        scale = ref.abs().max().item() + 1e-12
        assert (g - ref).abs().max().item() <= 2e-5 * scale, key


def test_golden_present():
    assert len(GOLD) >= 3
