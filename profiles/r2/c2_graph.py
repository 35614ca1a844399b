"""configs[1] (FSW_conv(64,64), N = 10k, E = 100k, fwd+bwd) is launch-latency bound: eager step vs the same step captured in a
CUDA graph and replayed (inputs in static buffers; the graph plan, the transposition and every scratch buffer are cached before
the capture, so the captured region holds kernel launches and allocator-pool memory only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from fsw_gnn_b200 import FSW_conv, _lib

dev = torch.device("cuda:0")
torch.manual_seed(1)
c2 = FSW_conv(64, 64, device=dev)
x2 = torch.randn(10000, 64, device=dev, requires_grad=True)
e2 = torch.randint(0, 10000, (2, 100000), device=dev)


def step():
    x2.grad = None
    for p in c2.parameters():
        p.grad = None
    loss = c2(x2, e2).square().sum()
    loss.backward()
    return loss


def timed(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


l0 = _lib.launch_count()
step()
launches = _lib.launch_count() - l0
ms_eager = timed(step)
ref_loss = float(step())
ref_grad = x2.grad.clone()

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss_static = step()
ms_graph = timed(g.replay)
g.replay()
torch.cuda.synchronize()
err = float((x2.grad - ref_grad).abs().max() / ref_grad.abs().max())
print("configs[1] step: eager %.3f ms (%d library launches per step), CUDA graph replay %.3f ms; loss %.6g vs %.6g, dX rel diff %.2e"
      % (ms_eager, launches, ms_graph, float(loss_static), ref_loss, err))
