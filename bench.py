#!/usr/bin/env python
"""Benchmark of the FSW hot path (BASELINE.json metric: FSW_conv edges/s & FSW_embedding multisets/s,
forward + backward, and % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU port of the reference algorithm (oracle/)

Headline workload (N = 1 and the 1/2/4/8-GPU scaling runs): BASELINE.json configs[3] - three
FSW_conv(100, 100) layers on an ogbn-products-shaped synthetic graph (2.4M vertices, ~62M edges), fp32,
destination vertices sharded edge-balanced over the GPUs, strong scaling.  `value` counts every edge once
per layer per step (edges x layers / step time), forward + backward.  One JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FSW_conv edges/sec & FSW_embedding multisets/sec fwd+bwd, % of HBM roofline"
N_VERT, N_EDGE, D_FEAT, N_LAYERS = 2_400_000, 62_000_000, 100, 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debugging only; invalid as a bench value)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (C2 / C3) and the CPU baseline")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def label_buckets(label):
    """buckets of the segment plan covered by a profiled kernel label, e.g. fwdr_small_u24_f32 -> uniform 17..24,
    bwd_rank_u64_f32 -> uniform 0..64, fwd_medium_u128_f32 -> uniform 65..128."""
    parts = label.split("_")
    if len(parts) < 3 or parts[2][:1] not in ("u", "g") or not parts[2][1:].isdigit():
        return []
    kind = 0 if parts[2][0] == "u" else 1
    size = int(parts[2][1:])
    base = kind * 519
    if parts[1] == "small":
        lo = ({4: 0, 8: 5, 12: 9, 16: 13, 24: 17, 32: 25, 48: 33, 64: 49} if kind == 0 else {4: 0, 8: 5, 16: 9, 32: 17, 64: 33})[size]
        return [base + b for b in range(lo, size + 1)]
    if parts[1] == "rank" and size <= 64:
        return [base + b for b in range(0, size + 1)]
    if parts[1] == "rankT":   # source-major rank backward: every uniform segment (up to 32768 elements)
        return [b for b in range(0, 519)]
    if parts[1] in ("coop", "cloud"):   # packed-key classes: two per power of two (3/4 and all of the slots)
        coop = {48: (33, 48), 64: (49, 64), 96: (65, 96), 128: (97, 128), 192: (129, 192), 256: (193, 256), 384: (257, 384),
                512: (385, 512), 1024: (513, 513)}
        lo, hi = coop.get(size, (518, 518))
        return [base + b for b in range(lo, hi + 1)]
    ranges = {64: (33, 64), 128: (65, 128), 256: (129, 256), 512: (257, 512), 1024: (513, 513), 2048: (514, 514), 4096: (515, 515),
              8192: (516, 516), 32768: (517, 517)}
    lo, hi = ranges.get(size, (518, 518))
    return [base + b for b in range(lo, hi + 1)]


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, by kernel label, from the `ncu --set full` capture of
    THIS binary at full size: profiles/r2/ncu_traffic.json, written by profiles/r2/ncu_traffic.py from the .ncu-rep that
    profiles/r2/ncu_fullscale.sh records (one layer of configs[3]).  Absent file / label -> None."""
    p = os.path.join(ROOT, "profiles", "r2", "ncu_traffic.json")
    try:
        return {k: float(v["dram_bytes_per_launch"]) for k, v in json.load(open(p))["kernels"].items()}
    except Exception:
        return {}


def class_fill(label, plan):
    """elements / sorting slots of a forward size class (the networks sort `size` slots per segment; the merge path has none)"""
    parts = label.split("_")
    bs = label_buckets(label)
    if not bs or parts[1] not in ("small", "coop"):
        return None
    size = int(parts[2][1:])
    segs = sum(plan.bucket_counts[b] for b in bs)
    elems = sum(plan.bucket_elems[b] for b in bs)
    return round(elems / max(segs * size, 1), 3)


def algorithmic_bytes(label, plan, K):
    """the fused kernel moves 4 B per (element, slice) gathered, 4 B per element (column id) and
    4 B per (segment, slice) written (forward) or read (backward: upstream gradient); the backward also
    scatters 4 B per (element, slice) into dXp."""
    bs = label_buckets(label)
    elems = sum(plan.bucket_elems[b] for b in bs)
    segs = sum(plan.bucket_counts[b] for b in bs)
    per_es = 8 if label.startswith("bwd") else 4
    return per_es * elems * K + 4 * elems + 4 * segs * K


# ----------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the C port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_conv_step(sample_edges, seed=0):
    """One fwd+bwd of the 3-layer FSW_conv stack on a destination-row SUBSAMPLE of the synthetic graph,
    computed by oracle/fsw_oracle.c (fp32 build, OpenMP over all host cores) + numpy for the small
    Linear/LeakyReLU part.  Returns (seconds, edges in the sample, threads)."""
    import numpy as np
    from oracle import c_oracle as C
    C.set_threads()   # every host core, whatever OMP_NUM_THREADS the launcher exported (torchrun sets 1)
    rng = np.random.default_rng(seed)
    mean_deg = N_EDGE / N_VERT
    mu = np.log(mean_deg) - 0.5
    S = max(int(sample_edges / mean_deg), 16)
    deg = np.clip(np.round(np.exp(mu + rng.standard_normal(S))), 1, 17000).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    E = int(rowptr[-1])
    Nsrc = max(S, 4096)  # sources live in a vertex set at least as large as the sampled rows
    col = rng.integers(0, Nsrc, E).astype(np.int32)
    K = 2 * D_FEAT - 1
    xs = [rng.standard_normal((Nsrc, D_FEAT)).astype(np.float32) for _ in range(N_LAYERS)]
    theta = rng.standard_normal((K, D_FEAT)).astype(np.float32)
    theta /= np.linalg.norm(theta, axis=1, keepdims=True)
    u = (0.5 + np.arange(K)) / K
    xi = (u / (1 - u)).astype(np.float32)
    Wl = (rng.standard_normal((D_FEAT, K + 1 + D_FEAT)) / np.sqrt(K + 1 + D_FEAT)).astype(np.float32)
    t0 = time.perf_counter()
    for layer in range(N_LAYERS):
        x = xs[layer]
        g_core = np.ones((S, K), dtype=np.float32)
        out, mass, dX, dth, dxi = C.embed_forward_backward(x, rowptr, col, None, theta, xi, g_core, dtype=np.float32)
        emb = np.concatenate([mass[:, None].astype(np.float32), out, x[:S]], axis=1)  # degree channel | embedding | self
        h = emb @ Wl.T
        h = np.where(h >= 0, h, 0.2 * h)
        gh = np.where(h >= 0, 1.0, 0.2).astype(np.float32) * (2.0 / h.size) * h
        _ = gh.T @ emb     # dW
        _ = gh @ Wl        # d emb
    dt = time.perf_counter() - t0
    return dt, E * 1, C.threads()


def reference_python_c2(steps=1):
    """The UNMODIFIED reference (fsw_conv.py / fsw_embedding.py, torch on the host cores, load_custom_cuda_lib off) on
    BASELINE.json configs[1]: FSW_conv(64, 64) fwd+bwd, N = 10k, E = 100k, fp32 - the largest configuration its
    ~126 B per (edge, slice) of live memory lets it run whole.  Needs the git-ignored copy baseline/_ref/ (made by
    __graft_entry__.build() where /root/reference exists); returns None when it is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "fsw_conv.py")):
        return None
    try:
        os.environ["FSW_REFERENCE_DIR"] = ref_dir
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
        import ref_loader
        emb_mod, conv_mod = ref_loader.load_reference()
        emb_mod.fsw_embedding_produce_error_on_custom_library_loading_failure = False   # public switch (fsw_embedding.py:116)
        torch.manual_seed(0)
        conv = conv_mod.FSW_conv(64, 64, device="cpu", dtype=torch.float32)
        emb_mod.libfsw_embedding = None   # CPU tensors: the reference's pure-torch segcumsum (its .so holds CUDA kernels only)
        x = torch.randn(10000, 64, requires_grad=True)
        ei = torch.randint(0, 10000, (2, 100000))
        ts = []
        for i in range(1 + steps):
            x.grad = None
            t0 = time.perf_counter()
            conv(x, ei).square().sum().backward()
            if i > 0:
                ts.append(time.perf_counter() - t0)
        t = sum(ts) / len(ts)
        return {"value": 100000 / t, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": "configs[1] whole (N=10k, E=100k, d=64, K=127, one layer fwd+bwd), reference Python on the host cores, %.1f s per step after one warm-up" % t}
    except Exception as e:   # the arm must never fail the bench
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args):
    """`--impl reference`: the reference's algorithm on the host cores (oracle port; the reference is a
    Python/torch program that cannot travel to the GPU box - DESIGN.md 'Reference arm')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 250_000
    times = []
    E = thr = 0
    for i in range(args.warmup + args.steps):
        dt, E, thr = cpu_conv_step(sample, seed=i)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / max(len(times), 1)
    val = E * N_LAYERS / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: 3 x FSW_conv(100,100) on an ogbn-products-shaped synthetic graph; CPU port timed on a "
                               "destination-row subsample of ~%d edges per step (rows are independent)" % E},
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": thr, "kind": "port",
                         "sample": "destination-row subsample with ~%d edges x %d layers per step (rows are independent), fp32, oracle/fsw_oracle.c (OpenMP) + numpy MLP" % (E, N_LAYERS)},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    ref_py = reference_python_c2()
    if ref_py is not None:
        line["reference_python_c2"] = ref_py
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from fsw_gnn_b200 import FSW_conv, FSW_embedding, _lib
    from fsw_gnn_b200 import dist as fdist
    from fsw_gnn_b200 import synthetic as syn
    from fsw_gnn_b200.graph import cached_graph, clear_graph_cache

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the FSW kernels have no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.load()

    Nv = max(int(N_VERT * args.scale), 1000)
    Ne = max(int(N_EDGE * args.scale), 10000)
    K = 2 * D_FEAT - 1

    # ---- workload: this rank's destination range of the synthetic graph ----
    deg = syn.products_like_degrees(Nv, Ne, seed=0, device=dev)
    ranges = syn.balanced_row_ranges(deg, world)
    lo, hi = ranges[rank]
    ei_local = syn.edges_for_rows(deg, lo, hi, Nv, seed=0, device=dev)
    E_total = int(deg.sum())
    E_local = int(ei_local.shape[1])
    n_local = hi - lo
    torch.manual_seed(0)
    layers = [FSW_conv(D_FEAT, D_FEAT, device=dev) for _ in range(N_LAYERS)]
    gx = torch.Generator(device=dev)
    gx.manual_seed(100 + rank)
    x_local = torch.randn(n_local, D_FEAT, device=dev, generator=gx)

    if world == 1:
        def prepare(ei):
            return ei

        def step(x, graph):
            h = x
            for conv in layers:
                h = conv(h, graph)
            loss = h.square().sum() / (Nv * D_FEAT)
            loss.backward()
            return loss
    else:
        def prepare(ei):
            return fdist.ShardedGraph(ei, ranges, rank, 1.0, torch.float32)

        def step(x, graph):
            h = x
            for conv in layers:
                h = fdist.sharded_conv_forward(conv, h, graph)
            loss = h.square().sum() / (Nv * D_FEAT)
            loss.backward()
            fdist.all_reduce_gradients(layers)
            return loss

    def zero_grads():
        for m in layers:
            for p in m.parameters():
                p.grad = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, device time from CUDA events, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # ---- device-resident run: inputs in HBM, graph plan prepared once (K0 is cacheable) ----
    graph = prepare(ei_local)
    x_req = x_local.clone().requires_grad_(True)

    def resident_step():
        zero_grads()
        x_req.grad = None
        step(x_req, graph)

    for _ in range(max(args.warmup, 3)):
        resident_step()
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed(resident_step, args.steps)
    clocks = sampler.stop()
    gpu_launches = _lib.launch_count() - launches0

    # ---- end-to-end run: host buffers in, loss out, graph preparation inside the step ----
    x_host = x_local.cpu().pin_memory()
    ei_host = ei_local.cpu().pin_memory()
    x_dev = torch.empty_like(x_local)
    ei_dev = torch.empty_like(ei_local)

    copy_stream = torch.cuda.Stream(device=dev)

    def prepare_training_graph(ei):
        """K0 for a training step: CSR + segment plan on the compute stream; the transposition the backward walks is built on a side
        stream under the first forward kernels (all cached on the graph, all inside the timed region)"""
        if world > 1:
            g = prepare(ei)
            g.plan.transpose_async(world * g.max_rows)   # side stream: first read by the backward
            return g
        cached_graph(ei, Nv, 0, "unit", 1.0, torch.float32)   # FSW_conv.forward starts the transposition itself (side stream)
        return ei

    def e2e_step():
        zero_grads()
        # the graph is needed first: edge_index goes over PCIe on the compute stream, the features follow on a side
        # stream and arrive while K0 prepares the graph (both copies are inside the timed region)
        ei_dev.copy_(ei_host, non_blocking=True)   # bumps the tensor version -> the graph is prepared again
        copy_stream.wait_stream(torch.cuda.current_stream(dev))   # x_dev is free again, and the link is ours after ei
        with torch.cuda.stream(copy_stream):
            x_dev.copy_(x_host, non_blocking=True)
        g = prepare_training_graph(ei_dev)
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        xr = x_dev.detach().requires_grad_(True)
        loss = step(xr, g)
        return float(loss.item())

    clear_graph_cache()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, max(2, min(args.steps, 5)))
    h2d = x_host.numel() * 4 + ei_host.numel() * 8
    d2h = 4

    # ---- roofline of the dominant kernel: per-kernel CUDA-event timers of the library, separate pass ----
    roofline = roofline_forward = None
    plan = graph.plan if world > 1 else cached_graph(ei_local, Nv, 0, "unit", 1.0, torch.float32)[1]
    resident_step()   # the e2e pass cleared the graph cache: prepare the resident graph (and its side-stream transposition) untimed
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    if world > 1:
        fdist.EXCHANGE_TIMING["on"] = True
    nprof = 2
    for _ in range(nprof):
        resident_step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    nccl_ms = None
    if world > 1:
        fdist.EXCHANGE_TIMING["on"] = False
        # time the compute stream spent between issuing a collective and having waited for it (skew between the ranks included)
        nccl_ms = {k: round(v / nprof, 3) for k, v in fdist.exchange_timing_read().items()}
    peaks, peak_kind = measured_peaks()
    traffic = ncu_traffic()
    kern = {k: v for k, v in prof.items() if k.startswith(("fwd", "bwd"))}
    breakdown = {k: round(v[1] / nprof, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    if kern:
        top = max(kern, key=lambda k: kern[k][1])
        cnt, tot_ms = kern[top]
        bytes_per_launch = algorithmic_bytes(top, plan, K)
        ach = bytes_per_launch / (tot_ms / cnt * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"],
                    "traffic": (traffic.get(top) if (world == 1 and args.scale == 1.0) else None),
                    "peak_source": peak_kind,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "ms_per_launch": tot_ms / cnt,
                    "share_of_step": (tot_ms / nprof) / ms_step}
        # the whole fused forward family (all size classes) for the north-star 70 % target
        fwd = {k: v for k, v in kern.items() if k.startswith("fwd")}
        fb = sum(algorithmic_bytes(k, plan, K) * v[0] for k, v in fwd.items())
        ft = sum(v[1] for v in fwd.values()) * 1e-3
        roofline["fused_forward_all_classes"] = {"achieved": fb / ft / 1e9, "frac": fb / ft / 1e9 / peaks["hbm_gbs"]}
        # the kernel family the north star sets its 70 % target on, as a roofline object of its own, with every size class
        # DRAM bytes of one layer's forward: one launch per size class; the merge-path classes share one entry in the capture
        ftraffic = [traffic.get(k, traffic.get("fwdr_medium_all") if "_medium_" in k else None) for k in fwd]
        roofline_forward = {"bound": "hbm", "kernel": "fused sort->cumsum->Fourier forward, all size classes (fwd*)",
                            "achieved": fb / ft / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fb / ft / 1e9 / peaks["hbm_gbs"],
                            "traffic": (sum(ftraffic) if ftraffic and all(t is not None for t in ftraffic) and world == 1 and args.scale == 1.0 else None),
                            "ms_per_step": ft * 1e3 / nprof, "share_of_step": ft * 1e3 / nprof / ms_step,
                            "classes": {k: {"ms_per_launch": v[1] / v[0], "GBps": algorithmic_bytes(k, plan, K) / (v[1] / v[0] * 1e-3) / 1e9,
                                            "elements": int(sum(plan.bucket_elems[b] for b in label_buckets(k))),
                                            "slot_fill": class_fill(k, plan)} for k, v in sorted(fwd.items())}}

    line = {
        "metric": METRIC, "value": E_total * N_LAYERS / (ms_step * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: %d x FSW_conv(%d,%d) (K=%d slices, learnable embedding) fwd+bwd on an ogbn-products-shaped "
                               "synthetic graph, N=%d vertices, E=%d edges (lognormal in-degree, mean %.1f, max 17000)"
                               % (N_LAYERS, D_FEAT, D_FEAT, K, Nv, E_total, E_total / Nv),
                   "edges_counted": "edges x layers per step", "parallelism": "dst-sharded x%d, all-gather(X) per layer" % world,
                   "l2": "inputs (>= 1 GB per tensor) far exceed the 126 MB L2", "scale": args.scale},
        "clocks": clocks,
        "e2e": {"value": E_total * N_LAYERS / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "d2h_is": "the 4-byte loss of the training step (the step's only host-visible result)",
                "includes": "H2D of edge_index, then of the vertex features (side stream, under K0) from pinned memory, graph preparation (K0: CSR, plan, transposition), fwd+bwd, D2H of the loss"},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline,
        "roofline_forward": roofline_forward,
        "hbm_reserved_gb": round(torch.cuda.max_memory_reserved(dev) / 1e9, 1),
        "kernel_ms_per_step": breakdown,
        "nccl_exposed_ms_per_step": nccl_ms,
    }

    # ---- secondary workloads + CPU baseline (rank 0 does the printing; every rank runs its shard) ----
    extras = {}
    if not args.no_extras:
        # C3: batched point clouds 256 x 1024 x 3 -> 256, batch sharded over the ranks, no communication
        B, n, d3, K3 = 256 // world, 1024, 3, 256
        emb = FSW_embedding(d3, K3, device=dev)
        X3 = torch.randn(B, n, d3, device=dev, requires_grad=True)

        def pc_step():
            X3.grad = None
            emb(X3).square().sum().backward()
        for _ in range(3):
            pc_step()
        ms_pc = timed(pc_step, 10)
        bytes_pc = (n * (16 * K3 + 12 * d3) + 8 * K3) * B * world
        extras["pointcloud_c3"] = {"multisets_per_s": B * world / (ms_pc * 1e-3), "ms_per_step": ms_pc,
                                   "roofline_frac_model": bytes_pc / (ms_pc * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                   "config": "configs[2]: 256 x 1024 points, d_in=3, d_out=256, fwd+bwd (dX), batch sharded x%d" % world}
        if rank == 0:
            # C2: demo_conv-shaped single layer (N=10k, E=100k, d=64), one GPU
            torch.manual_seed(1)
            c2 = FSW_conv(64, 64, device=dev)
            x2 = torch.randn(10000, 64, device=dev, requires_grad=True)
            e2 = torch.randint(0, 10000, (2, 100000), device=dev)

            def c2_step():
                x2.grad = None
                for p in c2.parameters():
                    p.grad = None
                c2(x2, e2).square().sum().backward()
            for _ in range(3):
                c2_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                c2_step()
            b.record()
            torch.cuda.synchronize()
            ms_c2 = a.elapsed_time(b) / 20
            l0 = _lib.launch_count()
            c2_step()
            c2_launches = _lib.launch_count() - l0
            extras["conv_c2"] = {"edges_per_s": 100000 / (ms_c2 * 1e-3), "ms_per_step": ms_c2, "library_launches_per_step": c2_launches,
                                 "config": "configs[1]: FSW_conv(64,64) fwd+bwd, N=10k, E=100k (launch-latency bound)"}
            # the same step captured in a CUDA graph and replayed: graph plan, transposition and scratch are cached before the
            # capture, so the captured region holds kernel launches and allocator-pool memory only (profiles/r2/c2_graph.py)
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        c2_step()
                torch.cuda.current_stream(dev).wait_stream(side)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    c2_step()
                for _ in range(3):
                    cg.replay()
                torch.cuda.synchronize()
                a.record()
                for _ in range(50):
                    cg.replay()
                b.record()
                torch.cuda.synchronize()
                ms_c2g = a.elapsed_time(b) / 50
                extras["conv_c2"]["cuda_graph_ms_per_step"] = ms_c2g
                extras["conv_c2"]["cuda_graph_edges_per_s"] = 100000 / (ms_c2g * 1e-3)
                del cg
            except Exception as ex:   # the eager number above stands
                extras["conv_c2"]["cuda_graph_error"] = str(ex)[:200]
        if rank == 0 and world == 1:
            # C5: configs[4] - power-law graph, 1M vertices, one 100 000-edge hub, d_in = 256 -> d_out = 512 (K = 511 slices):
            # the skewed segment sort and the tensor-core projection.  One FSW_embedding layer fwd+bwd through embed_plan.
            from fsw_gnn_b200.ops import SegmentPlan
            N5, d5 = 1_000_000, 256
            g5 = torch.Generator(device=dev); g5.manual_seed(5)
            u = torch.rand(N5, device=dev, generator=g5).clamp_(min=1e-9)
            deg5 = torch.floor(u.pow(-1.0 / 1.1)).clamp_(1, 100000).to(torch.int64)
            deg5[123] = 100000
            rowptr5 = torch.zeros(N5 + 1, dtype=torch.int32, device=dev)
            rowptr5[1:] = torch.cumsum(deg5, 0).to(torch.int32)
            E5 = int(rowptr5[-1])
            col5 = torch.randint(0, N5, (E5,), device=dev, generator=g5, dtype=torch.int32)
            plan5 = SegmentPlan(N5, E5, rowptr5, 0, col5, None, 1.0, torch.float32, dev)
            torch.manual_seed(5)
            emb5 = FSW_embedding(d_in=d5, d_out=512, encode_total_mass=True, learnable_slices=True, learnable_freqs=True,
                                 freqs_init="spread", minimize_slice_coherence=False, device=dev)
            X5 = torch.randn(N5, d5, device=dev, generator=g5).requires_grad_(True)

            def c5_step():
                X5.grad = None
                for p_ in emb5.parameters():
                    p_.grad = None
                emb5.embed_plan(X5, plan5).square().sum().backward()
            for _ in range(2):
                c5_step()
            ms_c5 = timed(c5_step, 3)
            _lib.profile_enable(True)
            _lib.profile_read()
            c5_step()
            torch.cuda.synchronize()
            p5 = _lib.profile_read()
            _lib.profile_enable(False)
            K5 = emb5.projVecs.shape[0]
            flops5 = 2.0 * N5 * d5 * K5
            tf32_peak = peaks.get("bf16_tflops_sustained", 1389.6) / 2.0
            umma = {k: {"ms": round(v[1] / v[0], 4), "fp32_tflops": round(flops5 / (v[1] / v[0] * 1e-3) / 1e12, 1),
                        "tensor_pipe_frac_of_tf32_peak": round(3 * flops5 / (v[1] / v[0] * 1e-3) / 1e12 / tf32_peak, 3)}
                    for k, v in p5.items() if k.startswith("umma")}
            extras["powerlaw_c5"] = {"edges_per_s": E5 / (ms_c5 * 1e-3), "ms_per_step": ms_c5, "edges": E5, "max_degree": int(deg5.max()),
                                     "K1_tensor_cores": umma, "tf32_peak_tflops_assumed": tf32_peak,
                                     "kernel_ms": {k: round(v[1], 3) for k, v in sorted(p5.items(), key=lambda kv: -kv[1][1])[:8]},
                                     "config": "configs[4]: FSW_embedding(256, 512) graph mode fwd+bwd, N=1M vertices, power-law in-degrees "
                                               "(Pareto 1.1, mean %.1f) with a 100 000-edge hub; K1 = 3xTF32 tcgen05 kernels, tensor-pipe "
                                               "fraction = 3 x fp32 flops / time / (bf16_tflops_sustained / 2)" % (E5 / N5)}
            del X5, plan5, emb5, col5
            torch.cuda.empty_cache()
            # Kseg: the standalone segmented scan behind segcumsum() (4 B read + 4 B written + 8 B of int64 segment id per element)
            from fsw_gnn_b200.ops import segcumsum_cuda
            nseg_el = 1 << 27
            vals = torch.rand(nseg_el, device=dev)
            ids = torch.repeat_interleave(torch.arange(nseg_el // 37 + 1, device=dev), 37)[:nseg_el].contiguous()
            for _ in range(3):
                segcumsum_cuda(vals, ids)
            ms_seg = timed(lambda: segcumsum_cuda(vals, ids), 10)
            extras["segcumsum_kseg"] = {"elements": nseg_el, "ms": ms_seg, "GBps": 16.0 * nseg_el / (ms_seg * 1e-3) / 1e9,
                                        "frac_of_hbm_peak": 16.0 * nseg_el / (ms_seg * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "model": "4 B read + 4 B written + 8 B int64 id per element, fp32, segments of 37"}
            del vals, ids
    if rank == 0 and world == 1 and not args.no_extras:
        dt, E_s, thr = cpu_conv_step(250_000, seed=0)  # warm
        dts = []
        for i in range(2):
            dt, E_s, thr = cpu_conv_step(250_000, seed=1 + i)
            dts.append(dt)
        t = sum(dts) / len(dts)
        line["cpu_baseline"] = {"value": E_s * N_LAYERS / t, "unit": "edges/s", "cores": thr, "kind": "port",
                                "sample": "destination-row subsample with ~%d edges x %d layers per step (rows are independent), fp32, "
                                          "oracle/fsw_oracle.c (OpenMP) + numpy MLP; %.1f s per step" % (E_s, N_LAYERS, t)}
        ref_py = reference_python_c2()
        if ref_py is not None:
            line["cpu_baseline"]["reference_python_c2"] = ref_py
    line["extra"] = extras
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
