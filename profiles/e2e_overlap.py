import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from fsw_gnn_b200 import synthetic as syn
from fsw_gnn_b200.graph import cached_graph, clear_graph_cache
dev = torch.device("cuda:0")
N, E = 2_400_000, 62_000_000
deg = syn.products_like_degrees(N, E, seed=0, device=dev)
ei = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev)
ei_host = ei.cpu().pin_memory()
x_host = torch.randn(N, 100).pin_memory()
ei_dev = torch.empty_like(ei); x_dev = torch.empty(N, 100, device=dev)
cs = torch.cuda.Stream(device=dev)

def seq():
    x_dev.copy_(x_host, non_blocking=True)
    ei_dev.copy_(ei_host, non_blocking=True)
    clear_graph_cache()
    csr, plan = cached_graph(ei_dev, N, 0, "unit", 1.0, torch.float32)
    plan._transpose = None; plan.transpose(N)
    return x_dev.sum()

def ovl():
    ei_dev.copy_(ei_host, non_blocking=True)
    cs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cs):
        x_dev.copy_(x_host, non_blocking=True)
    clear_graph_cache()
    csr, plan = cached_graph(ei_dev, N, 0, "unit", 1.0, torch.float32)
    plan._transpose = None; plan.transpose(N)
    torch.cuda.current_stream(dev).wait_stream(cs)
    return x_dev.sum()

for name, fn in (("sequential", seq), ("overlapped", ovl), ("sequential", seq), ("overlapped", ovl)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(name, "%.1f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
