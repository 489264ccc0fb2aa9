"""Developer tool: per-layer statistics of the tensor-core kNN filter on real SV-DGCNN features."""
import contextlib, ctypes, io, os, struct, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict

def stats():
    out = (ctypes.c_ulonglong * 12)()
    assert nv.lib().svnet_knn_tc_stats(out, ctypes.c_int(1)) == 0
    return list(out)

orig = nv.knn
def knn(view, B, N, k, **kw):
    torch.cuda.synchronize(); stats()
    os.environ["SVNET_KNN_TC_STATS"] = "1"
    orig(view, B, N, k, **kw); torch.cuda.synchronize()
    os.environ["SVNET_KNN_TC_STATS"] = "0"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(view, B, N, k, **kw); e1.record(); torch.cuda.synchronize()
    st = stats(); rows = max(st[0], 1)
    err = struct.unpack("f", struct.pack("I", st[9] & 0xffffffff))[0]
    cyc = tuple(v // max(st[4], 1) for v in st[5:9])
    print("  knn C=%d: %.1f us | exact rows %.1f%% brute %d survivors/row %.1f flagged/exact row %.1f max err %.2g | cycles/CTA A %d thr %d B %d fin %d" % (
        view.Cs + 3 * view.Cv, e0.elapsed_time(e1) * 1e3, 100.0 * st[1] / rows, st[2], st[3] / rows, st[11] / max(st[1], 1), err, *cyc))
    return r
nv.knn = knn
B = int(os.environ.get("KB", 32)); N = int(os.environ.get("KN", 1024)); SEED = int(os.environ.get("KSEED", 1002))
for binary in (True, False):
    with contextlib.redirect_stdout(io.StringIO()):
        net = sv.SV_DGCNN_CLS(make_args(k=20, binary=binary), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=SEED))
    net = net.cuda().eval()
    x = synthetic_clouds(B, N, SEED).cuda()
    print("binary" if binary else "fp")
    with torch.no_grad():
        net(x); print(" --"); net(x)
