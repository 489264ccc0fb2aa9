"""Developer tool: per-call CUDA-event times, in call order, of one SV-PointNet fp classification forward (cfg1 shape)."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_PointNet_CLS(make_args(k=20, binary=False), 40)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1001))
net = net.cuda().eval()
x = synthetic_clouds(B, 1024, 1001).cuda()
fused.CONCURRENT_HALVES = False
with torch.no_grad():
    for _ in range(3): net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(x); e1.record(); torch.cuda.synchronize()
    print("whole forward (eager) %.3f ms" % e0.elapsed_time(e1))
    nv.PROFILE[0] = set(nv.EXPORTS); nv.ORDER.clear()
    orig = nv.linear_rows
    shapes = []
    def lr(A, lda_g, lda_x, G, M, K, W, N, *a, **kw):
        shapes.append((G, M, K, N, kw.get("vbn", False)))
        return orig(A, lda_g, lda_x, G, M, K, W, N, *a, **kw)
    nv.linear_rows = lr
    net(x); torch.cuda.synchronize()
tot = 0.0; li = 0
for name, e0, e1 in nv.ORDER:
    t = e0.elapsed_time(e1); tot += t
    extra = ""
    if name == "svnet_linear_rows_ws":
        extra = str(shapes[li]); li += 1
    print("%-28s %8.1f us %s" % (name, 1e3 * t, extra))
print("sum %.3f ms for %d clouds" % (tot, B))
