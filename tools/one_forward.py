"""Developer tool: exactly two forwards of the cfg2 model on one stream (the first one packs the weights); the
command that runs under `ncu --set full` for profiles/<tag>_ncu_full_all_kernels.md."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import fused
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_CLS(make_args(k=20, binary=True), 40)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1002))
net = net.cuda().eval()
x = synthetic_clouds(32, 1024, 1002).cuda()
fused.CONCURRENT_HALVES = False
fused.SIDE_STREAM = False
with torch.no_grad():
    for _ in range(2):
        y = net(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
