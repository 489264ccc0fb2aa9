"""Developer tool: per-layer CUDA-event times of the edge layers (and totals) of one cfg2 forward, best of 5."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_CLS(make_args(k=20, binary=True), 40)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1002))
net = net.cuda().eval()
x = synthetic_clouds(B, 1024, 1002).cuda()
fused.CONCURRENT_HALVES = False
best = {}
with torch.no_grad():
    for it in range(8):
        nv.PROFILE[0] = {"svnet_svblock_edge_fwd", "svnet_knn_ws", "svnet_linear_rows_ws"} if it >= 3 else None
        nv.ORDER.clear()
        net(x)
        torch.cuda.synchronize()
        cnt = {}
        for name, e0, e1 in nv.ORDER:
            i = cnt.get(name, 0); cnt[name] = i + 1
            t = e0.elapsed_time(e1)
            best[(name, i)] = min(best.get((name, i), 1e9), t)
print(" ".join("%s[%d]=%.1fus" % (n.replace("svnet_", "").replace("_fwd", "").replace("_ws", ""), i, 1e3 * t) for (n, i), t in sorted(best.items())))
