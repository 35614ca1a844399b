"""Per-kernel DRAM traffic of the full-size capture (profiles/r2/ncu_fullscale.sh) -> profiles/r2/ncu_traffic.json, the file
bench.py reads `roofline.traffic` from.  Kernel names are mapped to the library's profile labels:
    fsw_rank_bwdS_kernel / fsw_rank_bwdT_kernel -> bwd_rankT_u32768_f32
    fsw_coop_fwd_kernel<R, L, col, rank>      -> fwd[r]_coop_u{R*L}_f32
    fsw_small_fwd_kernel<float, NP, col, rank>-> fwd[r]_small_u{NP}_f32
    fsw_umma_kernel<MODE>                     -> umma_nt / umma_nn / umma_tn (average over the launches captured)
python profiles/r2/ncu_traffic.py gpurun_out/r2_full_raw.csv"""
import csv, json, os, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(d, k):
    try:
        return float(d[ix[k]].replace(",", ""))
    except Exception:
        return None


def unit_scale(k):
    u = rows[1][ix[k]].lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)


def label(name):
    m = re.search(r"fsw_coop_fwd_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(bool\))?(\w+), (?:\(bool\))?(\w+)(?:, (?:\(bool\))?\w+)?>", name)
    if m:
        return ("fwdr" if m.group(4) in ("1", "true") else "fwd") + "_coop_u%d_f32" % (int(m.group(1)) * int(m.group(2)))
    m = re.search(r"fsw_small_fwd_kernel<float, (?:\(int\))?(\d+), (?:\(bool\))?(\w+), (?:\(bool\))?(\w+)>", name)
    if m:
        return ("fwdr" if m.group(3) in ("1", "true") else "fwd") + "_small_u%d_f32" % int(m.group(1))
    if "fsw_rank_bwdT_kernel" in name or "fsw_rank_bwdS_kernel" in name:
        return "bwd_rankT_u32768_f32"
    m = re.search(r"fsw_umma_kernel<(?:\(int\))?(\d)>", name)
    if m:
        return ("umma_nt", "umma_nn", "umma_tn")[int(m.group(1))]
    if "fsw_scale_grad_kernel" in name:
        return "bwd_scale_grad"
    if "fsw_medium_kernel" in name:
        return "fwdr_medium_all"
    return None


agg = {}
for d in data:
    lb = label(d[ix["Kernel Name"]])
    if lb is None:
        continue
    rd = num(d, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum")
    wr = num(d, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
    e = agg.setdefault(lb, dict(n=0, bytes=0.0, ns=0.0, regs=None, dram_pct=[], tensor_pct=[], issue_pct=[], occ=[]))
    e["n"] += 1
    e["bytes"] += rd + wr
    t = num(d, "gpu__time_duration.sum")
    tu = rows[1][ix["gpu__time_duration.sum"]].lower()
    e["ns"] += t * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(tu, 1.0)
    e["regs"] = num(d, "launch__registers_per_thread")
    for key, fld in (("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        if fld in ix and num(d, fld) is not None:
            e[key].append(num(d, fld))
out = {"source": os.path.basename(sys.argv[1]), "note": "ncu --set full --clock-control none, one layer of configs[3] at full size; "
       "times are cold-cache and serialised (compare shares, not absolutes)", "kernels": {}}
for lb, e in sorted(agg.items()):
    avg = lambda v: round(sum(v) / len(v), 2) if v else None
    out["kernels"][lb] = {"launches": e["n"], "dram_bytes_per_launch": e["bytes"] / e["n"], "ms_per_launch_under_ncu": e["ns"] / e["n"] / 1e6,
                          "registers": e["regs"], "dram_throughput_pct": avg(e["dram_pct"]), "tensor_pipe_pct": avg(e["tensor_pct"]),
                          "issue_active_pct": avg(e["issue_pct"]), "warps_active_pct": avg(e["occ"])}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
