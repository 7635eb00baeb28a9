"""Generates tests/golden/merge_endpoints.npz with the reference's own edge_extraction/merging.py:merge_endpoints
(pure numpy / scipy; open3d is stubbed at import). Run: python tests/golden/make_endpoints_golden.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("open3d", types.ModuleType("open3d"))
sys.path.insert(0, os.environ.get("REF", "/root/reference"))
from edge_extraction.merging import merge_endpoints  # noqa: E402

rng = np.random.default_rng(0)
hubs = rng.random((12, 3))
pick = lambda: hubs[rng.integers(0, 12)] + rng.normal(0, 0.004, 3)       # end points cluster around 12 junctions
lines = np.stack([np.concatenate([pick(), pick()]) for _ in range(25)])
curves = np.stack([np.concatenate([pick(), rng.random(3), rng.random(3), pick()]) for _ in range(20)])
far = rng.random((5, 6)) + 3.0                                            # isolated segments stay untouched
lines = np.concatenate([lines, far])
out_l, out_c = merge_endpoints(lines.copy(), curves.copy(), 0.015)
np.savez_compressed(os.path.join(HERE, "merge_endpoints.npz"), lines=lines, curves=curves, out_lines=out_l, out_curves=out_c)
print("moved", int((np.abs(out_l - lines) > 0).any(1).sum()), "lines,", int((np.abs(out_c - curves) > 0).any(1).sum()), "curves")
