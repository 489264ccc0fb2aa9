"""Developer tool: small-batch latency of the binary SV-DGCNN classifier (N=1024, k=20): CUDA-graph replay of the nn.Module
forward and of the whole-model C entry."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_CLS(make_args(k=20, binary=True), 40)
sd = synthetic_state_dict(net.state_dict(), seed=1002)
net.load_state_dict(sd)
net = net.cuda().eval()
native = sv.NativeModel("SV_DGCNN_CLS", sd, k=20, binary=True, num_class=40)

def best_ms(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
for B in (1, 2, 4, 8, 16, 32):
    x = synthetic_clouds(B, 1024, 1002).cuda()
    with torch.no_grad():
        g = sv.GraphedForward(net, x)
        t_g = best_ms(lambda: g(x))
        gc = torch.cuda.CUDAGraph()
        native(x); torch.cuda.synchronize()
        with torch.cuda.graph(gc):
            y = native(x)
        t_c = best_ms(lambda: gc.replay())
        t_e = best_ms(lambda: native(x))
    print("B=%2d: module graph %.3f ms | C entry graph %.3f ms | C entry eager %.3f ms | %.0f clouds/s" % (B, t_g, t_c, t_e, B / min(t_g, t_c) * 1e3))
