"""Developer tool: per-C-ABI-call CUDA-event times of one forward of a chosen model/config.
usage: python tools/profile_calls.py [cls|cls_fp|pseg|pointnet] [B] [N] [k]"""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict

kind = sys.argv[1] if len(sys.argv) > 1 else "cls"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
k = int(sys.argv[4]) if len(sys.argv) > 4 else 20
with contextlib.redirect_stdout(io.StringIO()):
    if kind == "pseg":
        net, extra = sv.SV_DGCNN_PSEG(make_args(k=k, binary=True), 50), (one_hot_labels(B).cuda(),)
    elif kind == "pointnet":
        net, extra = sv.SV_PointNet_CLS(make_args(k=k, binary=False), 40), ()
    else:
        net, extra = sv.SV_DGCNN_CLS(make_args(k=k, binary=(kind == "cls")), 40), ()
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1004))
net = net.cuda().eval()
x = synthetic_clouds(B, N, 1004).cuda()
fused.CONCURRENT_HALVES = False
with torch.no_grad():
    for _ in range(3):
        net(x, *extra)
    torch.cuda.synchronize()
    nv.PROFILE[0] = set(nv.EXPORTS)
    nv.ORDER.clear()
    net(x, *extra)
    torch.cuda.synchronize()
tot = 0.0
agg = {}
for name, e0, e1 in nv.ORDER:
    t = e0.elapsed_time(e1)
    tot += t
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1; agg[name][1] += t
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s x%-3d %8.3f ms  %5.1f%%" % (name, n, t, 100 * t / tot))
print("  sum %.3f ms for %d clouds" % (tot, B))
