"""Developer probe: how much of the tensor-core edge kernel's time is gather latency?  Same layer, same arithmetic,
three neighbour patterns: random rows of the cloud (like a real graph), the 20 following rows (sequential, L1/L2
friendly), and the point itself 20 times (every gather hits the line the centre loads just brought in)."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import fused
from svnet_b200.synthetic import synthetic_state_dict

B, N, k = 32, 1024, 20
for (Cs, Cv, Cout, Cvo) in ((32, 10, 32, 10), (32, 10, 64, 21), (64, 21, 128, 42)):
    with contextlib.redirect_stdout(io.StringIO()):
        blk = sv.SVBlock((2 * Cs, 2 * Cv), (Cout, Cvo), True)
    blk.load_state_dict(synthetic_state_dict(blk.state_dict(), seed=7))
    blk = blk.cuda().eval()
    R = B * N
    g = torch.Generator().manual_seed(1)
    s_in = torch.randn(R, Cs, generator=g).cuda()
    v_in = torch.randn(R, 3, Cv, generator=g).cuda()
    s_out = torch.empty(R, Cout, device="cuda")
    v_out = torch.empty(R, 3, Cvo, device="cuda")
    ar = torch.arange(N, dtype=torch.int32)
    pats = {"random": torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32),
            "next-20": ((ar.view(N, 1) + torch.arange(k, dtype=torch.int32).view(1, k)) % N).expand(B, N, k).contiguous(),
            "self": ar.view(1, N, 1).expand(B, N, k).contiguous()}
    out = []
    for name, idx in pats.items():
        idx = idx.cuda()
        with torch.no_grad():
            for _ in range(3):
                fused.sv_edge_layer(s_in, v_in, B, N, k, blk, s_out, v_out, idx32=idx)
            best = 1e9
            for _ in range(5):
                from svnet_b200 import _native as nv
                nv.PROFILE[0] = {"svnet_svblock_edge_fwd"}; nv.ORDER.clear()
                fused.sv_edge_layer(s_in, v_in, B, N, k, blk, s_out, v_out, idx32=idx)
                torch.cuda.synchronize()
                best = min(best, nv.ORDER[-1][1].elapsed_time(nv.ORDER[-1][2]))
                nv.PROFILE[0] = None
        out.append("%s %.1f us" % (name, 1e3 * best))
    print((Cs, Cv, Cout, Cvo), " | ".join(out))
