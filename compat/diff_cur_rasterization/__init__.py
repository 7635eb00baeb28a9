from curve_gaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                            rasterize_gaussians)
