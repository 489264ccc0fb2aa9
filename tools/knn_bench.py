"""Developer tool: kNN kernel alone -- tensor-core path (knn_tc.cu) vs CUDA-core path (knn.cu):
bit-exact agreement with each other and with the CPU oracle, filter statistics, CUDA-event times."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svnet_b200 import _native as nv  # noqa: E402


def stats(reset=True):
    out = (ctypes.c_ulonglong * 12)()
    rc = nv.lib().svnet_knn_tc_stats(out, ctypes.c_int(1 if reset else 0))
    assert rc == 0
    return list(out)


def run(feat, B, N, k, tc):
    os.environ["SVNET_KNN_TC"] = "1" if tc else "0"
    return nv.knn(nv.view_of(feat, None), B, N, k)[0]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    shapes = [(2, 130, 62, 20), (1, 200, 3, 1), (1, 1024, 3, 20), (4, 1024, 62, 20), (2, 1000, 127, 20),
              (32, 1024, 3, 20), (32, 1024, 62, 20), (32, 1024, 127, 20), (8, 4096, 62, 20)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        shapes = shapes[:5]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        shapes = shapes[5:8]
    check_oracle = "--oracle" in sys.argv
    for (B, N, C, k) in shapes:
        g = torch.Generator().manual_seed(B * 131 + N + C)
        x = torch.randn((B * N, C), generator=g)
        if C > 3:
            x = x * 0.3 + 1.0          # common offset: norms >> distances (the hard case for the filter)
        xd = x.cuda()
        stats()
        os.environ["SVNET_KNN_TC_STATS"] = "1"
        a = run(xd, B, N, k, True)
        torch.cuda.synchronize()
        os.environ["SVNET_KNN_TC_STATS"] = "0"
        st = stats()
        b = run(xd, B, N, k, False)
        torch.cuda.synchronize()
        same = bool((a == b).all())
        msg = "B=%d N=%d C=%d k=%d  tc==cuda-core: %s" % (B, N, C, k, same)
        if not same:
            bad = (a != b).any(dim=2).sum().item()
            msg += " (%d rows differ)" % bad
        if check_oracle and B * N <= 8192:
            from oracle import svnet_oracle as orc
            ref = orc.knn(x.view(B, N, C).numpy(), k)
            msg += "  tc==oracle: %s" % bool((a.cpu().numpy() == ref).all())
        rows = max(st[0], 1)
        msg += "  | rows %d exact %.1f%% brute %d survivors/row %.1f" % (st[0], 100.0 * st[1] / rows, st[2], st[3] / rows)
        import struct
        msg += "  | max err %.3g (2^%.1f)" % ((lambda e: (e, np.log2(max(e, 1e-30))))(struct.unpack("f", struct.pack("I", st[9] & 0xffffffff))[0]))
        if st[4]:
            msg += "  | cycles/CTA passA %d thr %d passB %d finish %d" % tuple(v // st[4] for v in st[5:9])
        t_tc = timeit(lambda: run(xd, B, N, k, True))
        t_cc = timeit(lambda: run(xd, B, N, k, False))
        msg += "  | tc %.1f us  cuda-core %.1f us" % (t_tc, t_cc)
        print(msg, flush=True)


if __name__ == "__main__":
    main()
