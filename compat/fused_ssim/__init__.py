from curve_gaussian_b200.ssim import FusedSSIMMap, fused_ssim, allowed_padding  # noqa: F401
