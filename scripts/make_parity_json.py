#!/usr/bin/env python
"""gpurun_out/parity/*.jsonl (written by the -m gpu tests through tests/parity.py) -> profiles/parity_r02.json:
every comparison the GPU tests made against the reference (CUDA extensions recompiled for sm_100, the reference's
Python step, the reference-made golden vectors), with the reference's own run-to-run noise next to each figure."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity")
out = {"how": "python -m pytest tests -m gpu on a B200 (gpurun); metrics in tests/parity.py", "suites": {}}
for f in sorted(glob.glob(os.path.join(src, "*.jsonl"))):
    rows = [json.loads(l) for l in open(f)]
    name = os.path.basename(f)[:-6]
    grads = [r for r in rows if "max_rel" in r and str(r.get("quantity", "")).lower().find("dl") >= 0]
    summ = {"comparisons": len(rows)}
    if grads:
        w = max(grads, key=lambda r: r["max_rel"])
        summ["worst_gradient_max_rel"] = {k: w.get(k) for k in ("case", "quantity", "max_rel", "ref_self_noise_max_rel", "tol")}
        summ["gradients_within_1e-5_max_rel"] = sum(r["max_rel"] <= 1e-5 for r in grads)
        summ["gradient_comparisons"] = len(grads)
    e2e = [r for r in rows if r.get("quantity") == "E2E gradients"]
    if e2e:
        summ["end_to_end_dL_dcontrol_points_max_rel"] = {
            r["case"]: {"ours_vs_reference": r["g_curve_points_max_rel"], "reference_vs_itself": r["g_curve_points_ref_self_noise_max_rel"]}
            for r in e2e}
    over = [r for r in grads if r["max_rel"] > 1e-5]
    if over:
        summ["gradients_beyond_1e-5"] = [{k: r.get(k) for k in ("case", "quantity", "max_rel", "ref_self_noise_max_rel", "tol")} for r in over]
    bit = [r for r in rows if r.get("bit_identical") is True]
    summ["bit_identical_comparisons"] = len(bit)
    out["suites"][name] = {"summary": summ, "records": rows}
dst = os.path.join(ROOT, "profiles", "parity_r02.json")
json.dump(out, open(dst, "w"), indent=1)
for k, v in out["suites"].items():
    print(k, json.dumps(v["summary"]))
print("wrote", dst)
