"""Developer check: tcgen05 vector-linear kernel vs the CUDA-core kernel on a conv5-shaped problem."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svnet_b200 import _native as nv  # noqa: E402


def run(R, K, N, rows_per_cloud, env):
    os.environ["SVNET_TCGEN05"] = env.get("SVNET_TCGEN05", "0")
    os.environ["SVNET_NO_TC"] = env.get("SVNET_NO_TC", "0")
    g = torch.Generator().manual_seed(0)
    v = torch.randn(R, 3, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    sc = (torch.rand(N, generator=g) + 0.5).cuda()
    a = (torch.rand(N, generator=g) + 0.5).cuda()
    c = (torch.randn(N, generator=g) * 0.1).cuda()
    B = R // rows_per_cloud
    gate = torch.rand(B, N, generator=g).cuda()
    out = torch.zeros(R, 3, N, device="cuda")
    nv.linear_rows(v, v.stride(0), v.stride(1), 3, 3 * R, K, W, N, out, out.stride(0), out.stride(1), sign_w=True,
                   colscale=sc, bn=(a, c), vbn=True, gate=gate, groups_per_cloud=rows_per_cloud)
    torch.cuda.synchronize()
    return out


def main():
    for (R, K, N, rpc) in [(128, 83, 170, 64), (1000, 83, 170, 250), (4096, 48, 96, 1024), (32768, 83, 170, 1024)]:
        ref = run(R, K, N, rpc, {"SVNET_NO_TC": "1", "SVNET_TCGEN05": "0"})
        got = run(R, K, N, rpc, {"SVNET_TCGEN05": "1"})
        err = (got - ref).abs().max().item()
        rel = err / ref.abs().max().item()
        print("R=%d K=%d N=%d: max abs err %.3g (rel %.3g) finite=%s" % (R, K, N, err, rel, bool(torch.isfinite(got).all())))
        assert rel < 1e-5, "tcgen05 result differs"
    # timing
    os.environ["SVNET_TCGEN05"] = "1"
    for env in ({"SVNET_TCGEN05": "1"}, {"SVNET_TCGEN05": "0"}, {"SVNET_NO_TC": "1", "SVNET_TCGEN05": "0"}):
        run(32768, 83, 170, 1024, env)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.Generator().manual_seed(0)
        v = torch.randn(32768, 3, 83, generator=g).cuda()
        W = torch.randn(170, 83, generator=g).cuda()
        sc = torch.rand(170).cuda(); a = torch.rand(170).cuda(); c = torch.rand(170).cuda(); gate = torch.rand(32, 170).cuda()
        out = torch.zeros(32768, 3, 170, device="cuda")
        os.environ["SVNET_TCGEN05"] = env.get("SVNET_TCGEN05", "0")
        os.environ["SVNET_NO_TC"] = env.get("SVNET_NO_TC", "0")
        f = lambda: nv.linear_rows(v, v.stride(0), v.stride(1), 3, 3 * 32768, 83, W, 170, out, out.stride(0), out.stride(1),
                                   sign_w=True, colscale=sc, bn=(a, c), vbn=True, gate=gate, groups_per_cloud=1024)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(env, "%.1f us" % (e0.elapsed_time(e1) * 100))


if __name__ == "__main__":
    main()
