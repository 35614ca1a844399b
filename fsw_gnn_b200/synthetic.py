"""Synthetic workloads of the BASELINE.json configs (seeded, generated on the device).

C4: ogbn-products-shaped graph - N = 2.4M vertices, ~62M directed edges, in-degrees lognormal with
mean 25.8 clipped to [1, 17000], sources uniform.  Any destination range of it can
be generated independently (same per-vertex degrees and per-edge sources on every rank), which is what
the destination-sharded multi-GPU runs use.
"""
import math

import torch


def lognormal_degrees(N, mean_deg, sigma, lo, hi, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    mu = math.log(mean_deg) - 0.5 * sigma * sigma
    z = torch.randn(N, generator=g, device=device, dtype=torch.float32)
    deg = torch.exp(mu + sigma * z).round().clamp_(lo, hi).to(torch.int64)
    return deg


def products_like_degrees(N=2_400_000, E_target=62_000_000, seed=0, device="cuda"):
    deg = lognormal_degrees(N, E_target / N, 1.0, 1, 17000, seed, device)
    # rescale so that the edge count lands on the target (clipping / rounding shift the mean a little)
    scale = E_target / float(deg.sum())
    deg = (deg.to(torch.float64) * scale).round().clamp_(1, 17000).to(torch.int64)
    return deg


def edges_for_rows(deg, row_lo, row_hi, N, seed, device, shuffle=True):
    """edge_index [2, E_local] (row 0 = source, row 1 = destination) of the destination rows
    [row_lo, row_hi).  Sources are a counter-based hash of the global edge number, so every rank
    produces the same graph whatever the partition."""
    d = deg[row_lo:row_hi]
    E = int(d.sum())
    start = int(deg[:row_lo].sum())
    dst = torch.repeat_interleave(torch.arange(row_lo, row_hi, device=device, dtype=torch.int64), d)
    eidx = torch.arange(start, start + E, device=device, dtype=torch.int64)
    # splitmix64-style hash (wraps in int64 arithmetic)
    x = eidx * (-7046029254386353131) + (seed * 2654435761 + 0x1234567)
    x = (x ^ (x >> 30)) * (-4658895280553007687)
    x = (x ^ (x >> 27)) * (-7723592293110705685)
    x = x ^ (x >> 31)
    src = (x & 0x7FFFFFFFFFFFFFFF) % N
    if shuffle:
        g = torch.Generator(device=device)
        g.manual_seed(seed + 17)
        perm = torch.randperm(E, generator=g, device=device)
        src, dst = src[perm], dst[perm]
    return torch.stack((src, dst), dim=0).contiguous()


def balanced_row_ranges(deg, parts):
    """Contiguous destination ranges with (nearly) equal edge counts."""
    csum = torch.cumsum(deg, 0)
    total = int(csum[-1])
    bounds = [0]
    for p in range(1, parts):
        target = total * p // parts
        bounds.append(int(torch.searchsorted(csum, torch.tensor(target, device=deg.device))))
    bounds.append(deg.numel())
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]
