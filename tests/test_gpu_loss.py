"""GPU: the fused edge+SSIM loss op and the direction-channel rotation against (a) the reference's own
formulation (torch edge_aware_loss, utils/loss_utils.py:94-115 + fused_ssim, train.py:101-107) and (b) the
CPU oracle. Tolerance 1e-5 relative on the scalar and on dL/dimage (fp32 summation order differs)."""
import pytest
import torch

from curve_gaussian_b200.loss import edge_aware_loss, edge_ssim_loss, rotate_channels
from curve_gaussian_b200.ssim import fused_ssim
from oracle import cpu_pipeline as CP

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _images(H, W, seed, edge_frac=0.08):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, H, W, generator=g)
    gt = (torch.rand(1, H, W, generator=g) < edge_frac).float() * torch.rand(1, H, W, generator=g)
    return img, gt


@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (150, 200), (1080, 1920)])
def test_fused_loss_matches_reference_formulation(cuda_dev, shape):
    H, W = shape
    img, gt = _images(H, W, seed=H + W)
    a = img.to(cuda_dev).requires_grad_(True)
    b = gt.to(cuda_dev)
    loss = edge_ssim_loss(a, b, threshold=0.1, lambda_mse=10.0, lambda_dssim=0.1)
    assert loss.shape == () and loss.is_cuda
    w = torch.tensor(1.7, device=cuda_dev)
    (loss * w).backward()

    a2 = img.to(cuda_dev).requires_grad_(True)
    ref = 10.0 * (0.9 * edge_aware_loss(a2, b) + 0.1 * (1.0 - fused_ssim(a2[None], b[None])))
    (ref * w).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert rel(a.grad, a2.grad) <= 1e-5


def test_fused_loss_matches_cpu_oracle_and_edge_cases(cuda_dev):
    img, gt = _images(70, 90, seed=3)
    o_loss, o_grad = CP.edge_ssim_loss_ref(img, gt)
    a = img.to(cuda_dev).requires_grad_(True)
    loss = edge_ssim_loss(a, gt.to(cuda_dev))
    loss.backward()
    assert abs(loss.item() - o_loss) <= 1e-5 * abs(o_loss)
    assert rel(a.grad, o_grad) <= 1e-5
    # no edge pixel at all / every pixel an edge: the class weights stay finite like the reference's
    for gt2 in (torch.zeros(1, 70, 90), torch.ones(1, 70, 90)):
        o_loss, o_grad = CP.edge_ssim_loss_ref(img, gt2)
        a = img.to(cuda_dev).requires_grad_(True)
        loss = edge_ssim_loss(a, gt2.to(cuda_dev))
        loss.backward()
        assert abs(loss.item() - o_loss) <= 1e-5 * abs(o_loss)
        assert rel(a.grad, o_grad) <= 1e-5
    # inference: no grad requested -> no partial maps saved, same value
    with torch.no_grad():
        v = edge_ssim_loss(img.to(cuda_dev), gt.to(cuda_dev))
    o_loss, _ = CP.edge_ssim_loss_ref(img, gt)
    assert abs(v.item() - o_loss) <= 1e-5 * abs(o_loss)
    with pytest.raises(Exception):
        edge_ssim_loss(torch.rand(3, 8, 8, device=cuda_dev), torch.rand(3, 8, 8, device=cuda_dev))


@pytest.mark.parametrize("shape", [(48, 64), (33, 35)])
def test_rotate_channels_matches_matmul(cuda_dev, shape):
    H, W = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, H, W, generator=g).to(cuda_dev)
    wvt = torch.randn(4, 4, generator=g).to(cuda_dev)
    planes = x[0:3].clone().requires_grad_(True)
    out = rotate_channels(planes, wvt[:3, :3])
    p2 = x[0:3].clone().requires_grad_(True)
    ref = (p2.permute(1, 2, 0) @ (wvt[:3, :3].T)).permute(2, 0, 1)   # gaussian_renderer/__init__.py:144
    assert rel(out, ref) <= 1e-6
    gw = torch.randn(3, H, W, generator=g).to(cuda_dev)
    (out * gw).sum().backward()
    (ref * gw).sum().backward()
    assert rel(planes.grad, p2.grad) <= 1e-6


def test_fused_clamp_equals_torch_clamp_then_loss(cuda_dev):
    """clamp=True on the raw render == render()'s torch clamp(0,1) followed by the loss, value and gradient
    (zero outside [0,1], passed on the closed interval's ends, like torch.clamp's adjoint)."""
    g = torch.Generator().manual_seed(11)
    raw = torch.rand(1, 90, 130, generator=g) * 1.6 - 0.3          # a good share below 0 and above 1
    raw[0, 0, :4] = torch.tensor([0.0, 1.0, -0.0, 1.0000001])
    gt = (torch.rand(1, 90, 130, generator=g) < 0.1).float()
    a = raw.to(cuda_dev).requires_grad_(True)
    loss = edge_ssim_loss(a, gt.to(cuda_dev), clamp=True)
    loss.backward()
    b = raw.to(cuda_dev).requires_grad_(True)
    ref = edge_ssim_loss(b.clamp(0, 1), gt.to(cuda_dev))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert rel(a.grad, b.grad) <= 1e-6
    outside = (raw < 0) | (raw > 1)
    assert float(a.grad.cpu()[outside].abs().max()) == 0.0
