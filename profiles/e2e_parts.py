"""Where the end-to-end step spends its time beyond the resident step: H2D copies and graph preparation (K0)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsw_gnn_b200 import synthetic as syn, _lib
from fsw_gnn_b200.graph import cached_graph, clear_graph_cache

dev = torch.device("cuda:0")
N, E = 2_400_000, 62_000_000
deg = syn.products_like_degrees(N, E, seed=0, device=dev)
ei = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev)
ei_host = ei.cpu().pin_memory()
x_host = torch.randn(N, 100).pin_memory()
ei_dev = torch.empty_like(ei); x_dev = torch.empty(N, 100, device=dev)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
        t.append((e0.elapsed_time(e1), (w1 - w0) * 1e3))
    return min(a for a, b in t), min(b for a, b in t)


print("H2D edge_index (%.2f GB): gpu %.1f ms wall %.1f ms" % ((ei_host.numel() * 8 / 1e9,) + timed(lambda: ei_dev.copy_(ei_host, non_blocking=True))))
print("H2D features   (%.2f GB): gpu %.1f ms wall %.1f ms" % ((x_host.numel() * 4 / 1e9,) + timed(lambda: x_dev.copy_(x_host, non_blocking=True))))


def prep():
    clear_graph_cache()
    return cached_graph(ei_dev, N, 0, "unit", 1.0, torch.float32)


print("graph preparation (CSR + plan): gpu %.1f ms wall %.1f ms" % timed(prep))
csr, plan = prep()


def tr():
    plan._transpose = None
    plan.transpose(N)


print("transposition: gpu %.1f ms wall %.1f ms" % timed(tr))
_lib.profile_enable(True)
prep(); csr, plan = prep(); plan._transpose = None; plan.transpose(N)
torch.cuda.synchronize()
for k, v in sorted(_lib.profile_read().items(), key=lambda kv: -kv[1][1]):
    print("  %-28s x%d %.2f ms" % (k, v[0], v[1]))
