import sys, os, torch
sys.path.insert(0, os.getcwd())
from svnet_b200 import _native as nv
B,N,C,k = 32,1024,62,20
g = torch.Generator().manual_seed(1)
x = (torch.randn((B*N, C), generator=g)*0.3+1.0).cuda()
for _ in range(2):
    nv.knn(nv.view_of(x, None), B, N, k)
torch.cuda.synchronize()
