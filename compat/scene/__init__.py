from curve_gaussian_b200.curve_model import GaussianCurveModel  # noqa: F401
