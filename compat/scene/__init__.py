"""`scene` shim: the B200-native GaussianCurveModel under the reference's package name.

Everything else in the reference's `scene` package - `Scene`, `scene.cameras`, `scene.dataset_readers`,
`scene.gaussian_model` - is reached through the NEXT `scene` directory on sys.path (the reference checkout), so
`from scene import Scene, GaussianCurveModel` (train.py:10) keeps working with this directory placed ahead of it:
the reference's `Scene` then drives our model (`create_from_pcd`, `save_ply`, `get_exposure_from_name`, ...).
"""
import os
import sys

from curve_gaussian_b200.curve_model import GaussianCurveModel  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_scene_dir():
    for p in sys.path:
        d = os.path.abspath(os.path.join(p or ".", "scene"))
        if d != _HERE and os.path.isfile(os.path.join(d, "__init__.py")):
            return d
    return None


_REF = _reference_scene_dir()
if _REF is not None:
    __path__.append(_REF)   # our gaussian_curve_model.py stays first; the other submodules come from the reference


def __getattr__(name):
    # `Scene` lives in the reference's scene/__init__.py; run that file in this module's namespace on first use
    # (its `from scene.gaussian_curve_model import GaussianCurveModel` resolves to ours, first on __path__)
    if name == "Scene" and _REF is not None:
        path = os.path.join(_REF, "__init__.py")
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), globals())
        return globals()["Scene"]
    raise AttributeError(f"module 'scene' has no attribute {name!r}")
