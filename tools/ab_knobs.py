#!/usr/bin/env python
"""Interleaved A/B of scorer knobs on ONE box in ONE process (boxes differ by a few % under the power cap, so variants
must alternate inside the same run).  Variants are per-handle options (rdb_set_option); "tc_debug" needs a library built
with RDB_PROFILING=1.

    python tools/ab_knobs.py N D Q k store metric 'tc_cta_group=1' 'tc_cta_group=2' ['a=1,b=2' ...]

Prints one JSON line per variant: median / min scorer-kernel ms (CUDA events inside the library) and TFLOP/s.
Profiling aid only -- not a product path."""
import importlib
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "radad-retrievalaugmenteddeepfakeaudiodetection_b200"
if os.environ.get("AB_PROF_LIB"):      # profiling build (make RDB_PROFILING=1 OBJDIR=build_prof OUT=../libradad_flat_prof.so)
    _cabi = importlib.import_module(PKG + "._cabi")
    _cabi.LIB_PATH = os.path.join(ROOT, PKG, "libradad_flat_prof.so")
pkg = importlib.import_module(PKG)


DEFAULTS = {"tc_cta_group": 0, "tc_lockstep": 8, "tc_lockstep_spins": 4096, "tc_stages": 64, "tc_chunks": 0, "tc_query_stationary": 1,
            "tc_pivot": 1, "tc_debug": 0, "tier1": 1, "tier1_kc": 0, "largek_scorer": 0, "largek_rows": 0,
            "largek_sample": 1, "largek_split": 1, "tc_list10": 1, "tier1_share2": 1, "host_pipeline": 1}


def main():
    N, D, Q, k = (int(v) for v in sys.argv[1:5])
    store, metric_s = sys.argv[5], sys.argv[6]
    variants = sys.argv[7:] or [""]
    rounds = int(os.environ.get("AB_ROUNDS", 6))
    dev = torch.device("cuda", 0)
    metric = pkg.METRIC_IP if metric_s.upper() == "IP" else pkg.METRIC_L2
    idx = pkg.FlatIndex(D, metric, store)
    idx.reserve(N)
    g = torch.Generator(device=dev)
    for c in range(0, N, 250_000):
        g.manual_seed(1234 + c)
        x = torch.randn((min(250_000, N - c), D), generator=g, device=dev)
        idx.add(x, normalize=(metric_s.upper() == "IP"))
    g.manual_seed(5678)
    q = torch.randn((Q, D), generator=g, device=dev)
    knobs = sorted({kv.split("=")[0] for v in variants for kv in v.split(",") if kv})
    ms = {v: [] for v in variants}
    wall = {v: [] for v in variants}
    ref = None
    for r in range(rounds + 1):
        for v in variants:
            for name in knobs:
                idx.set_option(name, DEFAULTS[name])
            for kv in v.split(","):
                if kv:
                    a, b = kv.split("=")
                    idx.set_option(a, int(b))
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                Dv, Iv = idx.search(q, k, normalize=(metric_s.upper() == "IP"))[:2]
                torch.cuda.synchronize()
                if r > 0:
                    ms[v].append(idx.last_kernel_ms()[0])
                    wall[v].append((time.perf_counter() - t0) * 1e3)
            if ref is None:
                ref = Iv.clone()
            elif not torch.equal(ref, Iv):
                print(json.dumps({"variant": v, "warning": "ids differ from the first variant",
                                  "mismatch_frac": float((ref != Iv).float().mean())}))
    flops = 2.0 * Q * N * D
    for v in variants:
        med, best = statistics.median(ms[v]), min(ms[v])
        print(json.dumps({"config": f"{N}x{D} {store} {metric_s} Q={Q} k={k}", "variant": v or "(default)",
                          "kernel_ms_median": round(med, 3), "kernel_ms_min": round(best, 3),
                          "tflops_median": round(flops / med / 1e9, 1), "samples": len(ms[v]),
                          "search_ms_median": round(statistics.median(wall[v]), 3),
                          "scorer": idx.last_kernel_ms()[1], "splits": idx.last_kernel_ms()[2]}), flush=True)


if __name__ == "__main__":
    main()
