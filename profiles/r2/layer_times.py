"""Per-kernel CUDA-event times of ONE FSW_conv(100,100) layer fwd+bwd on the configs[3] graph (library profile timers)."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from fsw_gnn_b200 import FSW_conv, _lib
from fsw_gnn_b200 import synthetic as syn

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda:0")
N, E = int(2_400_000 * scale), int(62_000_000 * scale)
deg = syn.products_like_degrees(N, E, seed=0, device=dev)
ei = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev)
torch.manual_seed(0)
conv = FSW_conv(100, 100, device=dev)
x = torch.randn(N, 100, device=dev, requires_grad=True)


def step():
    x.grad = None
    conv(x, ei).square().sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_read()
n = 3
for _ in range(n):
    step()
torch.cuda.synchronize()
rec = _lib.profile_read()
out = {k: round(v[1] / n, 3) for k, v in sorted(rec.items(), key=lambda kv: -kv[1][1])}
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("FSW_")}, "total_ms": round(sum(out.values()), 2), "kernels": out}))
