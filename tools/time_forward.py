"""Developer tool: time one SV-DGCNN forward on cuda:0 with CUDA events (eager and CUDA graph)."""
import argparse
import contextlib
import io
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv  # noqa: E402
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--fp", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    with contextlib.redirect_stdout(io.StringIO()):
        net = sv.SV_DGCNN_CLS(make_args(k=a.k, binary=not a.fp), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1002))
    net = net.cuda().eval()
    x = synthetic_clouds(a.B, a.N, 1002).cuda()
    with torch.no_grad():
        for _ in range(3):
            y = net(x)
        torch.cuda.synchronize()
        fwd = lambda: net(x)
        if a.graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                net(x)
            torch.cuda.current_stream().wait_stream(s)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                yg = net(x)
            fwd = gr.replay
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fwd()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            fwd()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
    print("B=%d N=%d k=%d %s graph=%s: %.3f ms/forward  %.1f clouds/s  (%.2f us/cloud)" % (
        a.B, a.N, a.k, "fp" if a.fp else "bin", a.graph, ms, a.B / ms * 1e3, ms * 1e3 / a.B))
    print("logits[0,:4] =", y[0, :4].tolist())
    if a.profile:
        from svnet_b200 import _native as nv
        nv.PROFILE[0] = set(nv.EXPORTS)
        nv.ORDER.clear()
        with torch.no_grad():
            net(x)
        torch.cuda.synchronize()
        tot = 0.0
        for name, s0, s1 in nv.ORDER:
            t = s0.elapsed_time(s1)
            tot += t
            print("  %-28s %8.3f ms" % (name, t))
        print("  sum %.3f ms" % tot)


if __name__ == "__main__":
    main()
