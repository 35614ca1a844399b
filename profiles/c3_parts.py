"""Per-kernel times of the configs[2] step (256 point clouds x 1024 points, d 3 -> 256, fwd+bwd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsw_gnn_b200 import FSW_embedding, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
mod = FSW_embedding(d_in=3, d_out=256, device=dev, learnable_slices=True, learnable_freqs=True)
X = torch.randn(256, 1024, 3, device=dev, requires_grad=True)
for _ in range(3):
    mod(X).square().sum().backward()
torch.cuda.synchronize()
_lib.profile_enable(True)
n = 5
for _ in range(n):
    mod(X).square().sum().backward()
torch.cuda.synchronize()
tot = 0.0
for k, v in sorted(_lib.profile_read().items(), key=lambda kv: -kv[1][1]):
    print("  %-28s x%-3d %.3f ms per step" % (k, v[0] // n, v[1] / n))
    tot += v[1] / n
print("library kernels: %.3f ms per step" % tot)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.profile_enable(False)
e0.record()
for _ in range(20):
    mod(X).square().sum().backward()
e1.record(); torch.cuda.synchronize()
print("step: %.3f ms" % (e0.elapsed_time(e1) / 20))
