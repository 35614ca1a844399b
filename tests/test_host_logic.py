"""Host-side logic that needs no GPU: column chunking of the multi-GPU exchange, the benchmark's mapping from
profiled kernel labels to plan buckets (the roofline accounting), constants shared between the header and Python."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_column_chunks_cover_all_slices_in_multiples_of_8():
    from fsw_gnn_b200.ops import column_chunks
    for K in (0, 1, 7, 8, 9, 63, 64, 199, 200, 255, 1023):
        for n in (1, 2, 3, 4, 8, 100):
            ch = column_chunks(K, n)
            if K == 0:
                assert ch == [(0, 0)]
                continue
            assert ch[0][0] == 0 and ch[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert all(k0 % 8 == 0 and k1 > k0 for k0, k1 in ch)
            assert len(ch) <= max(1, min(n, (K + 7) // 8))


def test_bench_label_buckets_partition_the_uniform_buckets():
    b = _bench()
    # every forward class label of the C4 run maps to a disjoint bucket range; together they cover 0..516
    labels = (["fwdr_small_u%d_f32" % n for n in (4, 8, 12, 16, 24, 32)] +
              ["fwdr_coop_u%d_f32" % n for n in (48, 64, 96, 128, 192, 256, 384, 512)] +
              ["fwdr_medium_u%d_f32" % n for n in (1024, 2048, 4096, 8192, 32768, 131072)])
    seen = []
    for lab in labels:
        bs = b.label_buckets(lab)
        assert bs, lab
        seen += bs
    assert sorted(seen) == list(range(0, 519)), "forward labels must tile the uniform-weight buckets exactly once"
    assert b.label_buckets("bwd_rankT_u32768_f32") == list(range(0, 519))
    assert b.label_buckets("bwd_rank_dense_f32") == [] and b.label_buckets("coef_tables") == []
    # general-weight labels live in the second kind
    assert b.label_buckets("fwd_small_g16_f32") == [519 + n for n in range(9, 17)]


def test_python_constants_match_the_header():
    from fsw_gnn_b200 import _lib, ops
    hdr = open(os.path.join(ROOT, "include", "fsw_embedding.h")).read()
    per_kind = int(re.search(r"#define FSW_PLAN_BUCKETS_PER_KIND (\d+)", hdr).group(1))
    assert _lib.PLAN_BUCKETS_PER_KIND == per_kind == 519
    assert ops.RANKT_NMAX == int(re.search(r"#define FSW_RANKT_NMAX (\d+)", hdr).group(1))
    assert "FSW_RANKT_ELIGIBLE" in hdr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference's algorithm on the host cores) needs no GPU:
    one JSON line with the contract's keys, its own cpu_baseline and a zero-copy e2e."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
