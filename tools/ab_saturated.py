#!/usr/bin/env python3
"""Times ONE saturated k_search_group launch (every resident group busy for MAPAD_PROFILE_ITERS expansions, batch discarded)
on a small forced-wide index, for the library selected with MAPAD_GPU_LIB — a seconds-long A/B of kernel variants.
Measurement tool, not a bench value.  Usage: [MAPAD_GPU_LIB=path] python tools/ab_saturated.py [iters=30000] [genome_bp=20e6]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MAPAD_WS_BYTES", str(8 << 30))
os.environ.setdefault("MAPAD_FORCE_WIDE", "1")
os.environ.setdefault("MAPAD_GROUP", "32")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
genome_bp = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
from mapad_b200 import api, workloads  # noqa: E402
from mapad_b200.specs import cli_spec as cli_params, product_params  # noqa: E402  (oracle-free)

genome = workloads.random_genome_array(genome_bp, seed=42)
index = api.Index.build(workloads.split_contigs(genome, 4), seed=1234, device=0)
mapper = api.Mapper(index, product_params(dict(cli_params("single_stranded"))), device=0)
seq, qual, off = workloads.simulate_batch(genome, 20_000, (86, 100), seed=79, library="single_stranded")
R, keep = api.make_reads(seq, qual, off, np.arange(20_000, dtype=np.uint32))
times = []
for n in (300, iters, iters):
    os.environ["MAPAD_PROFILE_ITERS"] = str(n)
    t0 = time.perf_counter()
    try:
        mapper.map_raw(R, 0)
    except api.MapadError:
        pass
    times.append((time.perf_counter() - t0) * 1e3)
groups = 16 * 148
print("lib=%s iters=%d ms=%s frames_per_s=%.4g" % (os.environ.get("MAPAD_GPU_LIB", "default"), iters, ["%.1f" % t for t in times],
                                                   groups * iters / (min(times[1:]) * 1e-3)), flush=True)
