"""Small driver for ncu captures: one FSW_conv(100,100) layer fwd+bwd on a scaled products-like graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsw_gnn_b200 import FSW_conv
from fsw_gnn_b200 import synthetic as syn

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
N, E = int(2_400_000 * scale), int(62_000_000 * scale)
deg = syn.products_like_degrees(N, E, seed=0, device=dev)
ei = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev)
torch.manual_seed(0)
conv = FSW_conv(100, 100, device=dev)
x = torch.randn(N, 100, device=dev, requires_grad=True)
for _ in range(iters):
    x.grad = None
    conv(x, ei).square().sum().backward()
torch.cuda.synchronize()
print("ok", N, int(deg.sum()))
