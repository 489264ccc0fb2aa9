"""Throughput of the other BASELINE.json configs on one GPU (cfg3 fp cls, cfg4 binary pseg, cfg5 sweep).
Device-resident inputs, CUDA-event timing, 3 warm-up + best-of-5.  Output: one JSON object."""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv  # noqa: E402
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict  # noqa: E402


def build(cls, args, ncls, seed):
    with contextlib.redirect_stdout(io.StringIO()):
        net = cls(args, ncls)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=seed))
    return net.cuda().eval()


def time_ms(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    return best


def main():
    out = {}
    with torch.no_grad():
        net = build(sv.SV_DGCNN_CLS, make_args(k=20, binary=False), 15, 1003)
        x = synthetic_clouds(256, 1024, 1003).cuda()
        ms = time_ms(lambda: net(x))
        out["cfg3 SV-DGCNN fp ScanObjectNN-shaped cls B=256 N=1024 k=20"] = {"ms": ms, "clouds_per_s": 256 / ms * 1e3}
        net = build(sv.SV_DGCNN_PSEG, make_args(k=40, binary=True), 50, 1004)
        x = synthetic_clouds(16, 2048, 1004).cuda()
        l = one_hot_labels(16).cuda()
        ms = time_ms(lambda: net(x, l))
        out["cfg4 SV-DGCNN binary pseg B=16/GPU N=2048 k=40"] = {"ms": ms, "clouds_per_s": 16 / ms * 1e3}
        net = build(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40, 1005)
        for B, N in [(32, 1024), (128, 1024), (512, 1024), (2048, 1024), (4096, 1024), (32, 2048), (128, 2048), (512, 2048), (32, 4096), (128, 4096)]:
            x = synthetic_clouds(B, N, 1005).cuda()
            ms = time_ms(lambda: net(x), reps=3 if B <= 512 else 2)
            out["cfg5 SV-DGCNN binary cls B=%d N=%d k=20" % (B, N)] = {"ms": ms, "clouds_per_s": B / ms * 1e3}
        net = build(sv.SV_PointNet_CLS, make_args(k=20, binary=False), 40, 1001)
        x = synthetic_clouds(32, 1024, 1001).cuda()
        ms = time_ms(lambda: net(x))
        out["cfg1 SV-PointNet fp cls B=32 N=1024 k=20 (GPU)"] = {"ms": ms, "clouds_per_s": 32 / ms * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
