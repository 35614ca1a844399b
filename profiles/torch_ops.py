"""Lists the CUDA kernels torch itself launches around the library calls (one FSW_conv layer fwd+bwd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from fsw_gnn_b200 import FSW_conv
from fsw_gnn_b200 import synthetic as syn

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
dev = torch.device("cuda:0")
N, E = int(2_400_000 * scale), int(62_000_000 * scale)
deg = syn.products_like_degrees(N, E, seed=0, device=dev)
ei = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev)
torch.manual_seed(0)
conv = FSW_conv(100, 100, device=dev)
x = torch.randn(N, 100, device=dev, requires_grad=True)
for _ in range(2):
    conv(x, ei).square().sum().backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    conv(x, ei).square().sum().backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
