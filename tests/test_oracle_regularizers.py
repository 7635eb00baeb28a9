"""CPU: the oracle restatements of the curve-side regularisers (oracle/torch_ref.py) against a literal
transcription of the reference's train.py lines (default torch.cdist mode, pytorch3d-style matrix)."""
import torch

from curve_gaussian_b200 import synth
from oracle import torch_ref


def test_endpoint_connectivity_equals_literal_train_py():
    cp, _, _, _ = synth.random_curves(400, seed=2)
    cp = ((cp - 0.5) * 0.5 + 0.5).double()
    # train.py:133-146, verbatim
    start_points, end_points = cp[:, 0], cp[:, -1]
    all_points = torch.cat([start_points, end_points], dim=0)
    mask = torch.eye(len(start_points), dtype=torch.bool)
    mask = torch.cat([torch.cat([mask, mask], dim=1), torch.cat([mask, mask], dim=1)], dim=0)
    dist = torch.cdist(all_points, all_points, p=2)
    valid_mask = (dist < 0.05) & (~mask)
    assert valid_mask.any()
    want = dist[valid_mask].mean()
    got = torch_ref.endpoint_connectivity(cp, 0.05)
    assert abs(got.item() - want.item()) <= 1e-9
    assert torch_ref.endpoint_connectivity(cp[:1], 0.05).item() == 0.0


def test_curve_smoothness_is_zero_for_lines_and_positive_for_bends():
    n = 10
    cp, width, _, _ = synth.random_curves(20, seed=1)
    isb = torch.zeros(20, dtype=torch.bool)
    _, rot, _ = torch_ref.sample_curves(cp, width, isb, n)
    assert torch_ref.curve_smoothness(rot.double(), n).item() < 1e-6      # straight segments: constant direction
    _, rot_b, _ = torch_ref.sample_curves(cp, width, ~isb, n)
    assert torch_ref.curve_smoothness(rot_b.double(), n).item() > 1e-6
