#!/usr/bin/env python
"""Full-size agreement statistics of the CUDA path against the reference forward (oracle/_ref) on the
GPU box; writes profiles/<tag>_ref_parity.json.   python tools/ref_parity.py [tag]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refparity  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
res = {}
cases = [("cfg2", dict(kind="SV_DGCNN_CLS", B=4, N=1024, k=20, binary=True, ncls=40, seed=1002)),
         ("cfg3", dict(kind="SV_DGCNN_CLS", B=4, N=1024, k=20, binary=False, ncls=15, seed=1003)),
         ("cfg4", dict(kind="SV_DGCNN_PSEG", B=2, N=2048, k=40, binary=True, ncls=50, seed=1004))]
for name, kw in cases:
    out, (ours, rnet, x, extra, y_ref) = refparity.measure(**kw)
    if name == "cfg2":
        out["reference_cpu_vs_reference_cuda"] = refparity.reference_cpu_vs_cuda(rnet, x, extra, y_ref)
    res[name] = out
    print(name, json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for d in ("gpurun_out", "profiles"):
    json.dump(res, open(os.path.join(ROOT, d, tag + "_ref_parity.json"), "w"), indent=1)
