from curve_gaussian_b200.renderer import render  # noqa: F401
