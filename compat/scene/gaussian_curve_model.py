from curve_gaussian_b200.curve_model import GaussianCurveModel, initialize_bezier_curves  # noqa: F401
