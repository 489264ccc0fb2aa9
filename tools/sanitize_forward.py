#!/usr/bin/env python
"""One small forward of every SV model (binary and full precision) through the CUDA library: the command that
runs under `compute-sanitizer --tool memcheck|racecheck` (profiles/r2_sanitizer_*.log).  Shapes are chosen so that
the tcgen05 kernels run: tensor-core kNN (N >= 64), tensor-core edge kernels (k = 20 / 40), vector linears,
conv5's binary linear (rows >= 2048)."""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv  # noqa: E402
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict  # noqa: E402

CASES = [("SV_DGCNN_CLS", dict(k=20, binary=True), 40, 2, 1024, False),
         ("SV_DGCNN_CLS", dict(k=20, binary=False), 15, 2, 256, False),
         ("SV_DGCNN_PSEG", dict(k=40, binary=True), 50, 1, 2048, True),
         ("SV_DGCNN_PSEG", dict(k=40, binary=False), 50, 1, 256, True),
         ("SV_PointNet_CLS", dict(k=20, binary=True), 40, 2, 256, False),
         ("SV_PointNet_CLS", dict(k=20, binary=False), 40, 2, 256, False),
         ("SV_PointNet_PSEG", dict(k=20, binary=True), 50, 1, 256, True)]
for kind, margs, ncls, B, N, lab in CASES:
    with contextlib.redirect_stdout(io.StringIO()):
        net = getattr(sv, kind)(make_args(**margs), ncls)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=3))
    net = net.cuda().eval()
    x = synthetic_clouds(B, N, 3).cuda()
    ins = (x, one_hot_labels(B).cuda()) if lab else (x,)
    if kind == "SV_PointNet_PSEG":
        ins = (x, one_hot_labels(B).view(B, 1, 16).cuda())
    with torch.no_grad():
        y = net(*ins)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    print("%-18s binary=%-5s B=%d N=%d k=%d -> %s ok" % (kind, margs["binary"], B, N, margs["k"], tuple(y.shape)), flush=True)
