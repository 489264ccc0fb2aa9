"""Developer probe: does a spatial (Morton) order of the points inside a cloud speed the gathers up?
Times the per-call kernels of one cfg2 forward with the clouds as generated and with Morton-sorted points."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict


def morton(x):           # x (B,3,N) in [-1,1]
    q = ((x.clamp(-1, 1) * 0.5 + 0.5) * 1023).long()
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    perm = code.argsort(dim=1)
    return torch.gather(x, 2, perm.unsqueeze(1).expand(-1, 3, -1)).contiguous()


with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_CLS(make_args(k=20, binary=True), 40)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1002))
net = net.cuda().eval()
x0 = synthetic_clouds(32, 1024, 1002).cuda()
fused.CONCURRENT_HALVES = False
for tag, x in (("as generated", x0), ("morton order", morton(x0))):
    best = {}
    with torch.no_grad():
        for it in range(8):
            nv.PROFILE[0] = set(nv.EXPORTS) if it >= 3 else None
            nv.ORDER.clear()
            net(x)
            torch.cuda.synchronize()
            cnt = {}
            for name, e0, e1 in nv.ORDER:
                i = cnt.get(name, 0); cnt[name] = i + 1
                best[(name, i)] = min(best.get((name, i), 1e9), e0.elapsed_time(e1))
    agg = {}
    for (n, i), t in best.items():
        agg.setdefault(n, []).append((i, t))
    print(tag, "total %.1f us" % (1e3 * sum(best.values())))
    for n in ("svnet_knn_ws", "svnet_svblock_edge_fwd", "svnet_edge_xyz_fwd", "svnet_linear_rows_ws", "svnet_gate_edge"):
        print("   %-24s %s" % (n, " ".join("%.1f" % (1e3 * t) for i, t in sorted(agg.get(n, [])))))
