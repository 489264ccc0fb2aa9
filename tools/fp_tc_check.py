"""Developer tool: full-precision tensor-core edge kernel (csrc/edge_fp_tc.cu) against the CUDA-core kernel
(SVNET_EDGE_FP_TC=0) on the same inputs and the same kNN graphs; per-layer max differences and times."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_CLS(make_args(k=20, binary=False), 15)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1003))
net = net.cuda().eval()
x = synthetic_clouds(B, 1024, 1003).cuda()
fused.CONCURRENT_HALVES = False
outs = {}
with torch.no_grad():
    os.environ["SVNET_EDGE_FP_TC"] = "0"
    rec0 = {}
    y0 = net(x, record=rec0)
    idx = rec0["idx"]
    for tc in ("0", "1"):
        os.environ["SVNET_EDGE_FP_TC"] = tc
        rec = {}
        y = net(x, forced_idx=idx, record=rec)
        torch.cuda.synchronize()
        outs[tc] = (y.clone(), rec["s_cat"].clone(), rec["v_cat"].clone())
    y0, s0, v0 = outs["0"]
    y1, s1, v1 = outs["1"]
    so = 0
    for li, cs in enumerate((32, 32, 64, 128)):
        d = (s0[:, so:so + cs] - s1[:, so:so + cs]).abs().max().item()
        m = s0[:, so:so + cs].abs().max().item()
        print("layer %d s: max |diff| %.3g (max |s| %.3g)" % (li + 1, d, m))
        so += cs
    print("v_cat max |diff| %.3g (max %.3g)" % ((v0 - v1).abs().max().item(), v0.abs().max().item()))
    print("logits max |diff| %.3g, argmax equal %s" % ((y0 - y1).abs().max().item(), bool((y0.argmax(1) == y1.argmax(1)).all())))
    for tc in ("0", "1"):
        os.environ["SVNET_EDGE_FP_TC"] = tc
        for _ in range(2):
            net(x)
        torch.cuda.synchronize()
        nv.PROFILE[0] = set(nv.EXPORTS); nv.ORDER.clear()
        net(x); torch.cuda.synchronize()
        t = [e0.elapsed_time(e1) * 1e3 for name, e0, e1 in nv.ORDER if name == "svnet_svblock_edge_fwd"]
        tot = sum(e0.elapsed_time(e1) for name, e0, e1 in nv.ORDER)
        nv.PROFILE[0] = None
        print("tc=%s edge layers us: %s | forward %.3f ms for %d clouds" % (tc, ["%.0f" % v for v in t], tot, B))
