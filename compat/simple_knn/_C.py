from curve_gaussian_b200.knn import distCUDA2  # noqa: F401
