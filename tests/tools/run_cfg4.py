#!/usr/bin/env python3
"""hg19-scale probe (BASELINE cfg4): builds the 3.1 Gbp synthetic index with the device suffix sorter, verifies the CUDA path
against the oracle on a sample of reads (positions beyond 2^32, wide 64 B occ blocks, HBM-resident index), then maps one
chunk per search-kernel variant and records reads/s, frames/s and the per-read frame counts (gpurun_out/cfg4_frames.npz).
Measurement tool — the numbers it prints are not bench values.
Usage: python tests/tools/run_cfg4.py [genome_bp] [n_reads] [n_parity] [variants, e.g. 8,32,1]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapad_b200 import abi, api, workloads  # noqa: E402
from compare import compare_results  # noqa: E402
from helpers import oracle_params, product_params  # noqa: E402
from ref_cases import cli_params  # noqa: E402

genome_bp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
n_parity = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["8"]
out = dict(genome_bp=genome_bp)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
t = time.time()
genome = workloads.random_genome_array(genome_bp, seed=42)
out["genome_s"] = time.time() - t
t = time.time()
index = api.Index.build(workloads.split_contigs(genome, 24), device=0 if genome_bp > 500_000_000 else None)
out["index_build_s"] = time.time() - t
print("index built", out, flush=True)
spec = cli_params("single_stranded")
params = product_params(spec)
t = time.time()
mapper = api.Mapper(index, params)
out["relayout_upload_s"] = time.time() - t
out["index_bytes_hbm"] = mapper.export_index()[2]
seq, qual, off = workloads.simulate_batch(genome, n_reads, (25, 100), seed=1004000)
if n_parity:
    from oracle import oracle as ora
    a = index.arrays()
    oix = ora.OracleIndex.from_arrays(a["bwt"], a["sa_sample"], a["sa_rate"], a["extra_rows"], a["contigs"], a["orig_pos"], a["orig_sym"])
    del a
    sub = (seq[: int(off[n_parity])], qual[: int(off[n_parity])], off[: n_parity + 1])
    seeds = np.arange(n_parity, dtype=np.uint32)
    t = time.time()
    want = ora.map_batch(oix, oracle_params(spec), None, None, seeds=seeds, n_threads=os.cpu_count(), want_hits=True, packed=sub)
    out["oracle_reads_per_s"] = n_parity / (time.time() - t)
    out["oracle_threads"] = os.cpu_count()
    for g in variants:
        os.environ["MAPAD_GROUP"] = g
        got = mapper.map_batch(seeds=seeds, want_hits=True, packed=sub)
        compare_results(want, got)
    out["parity_reads"] = n_parity
    out["parity_variants"] = variants
    out["max_absolute_pos"] = int(got.records["absolute_pos"].max())
    del oix
    print("parity ok", out, flush=True)
R = api.make_reads(seq, qual, off, np.arange(n_reads, dtype=np.uint32))
runs = []
frames = None
for g in variants:
    os.environ["MAPAD_GROUP"] = g
    t = time.time()
    res = mapper.map_raw(R[0], 0)
    dt = time.time() - t
    recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
    fr = recs["frames_popped"].astype(np.int64)
    frames = fr.copy()
    runs.append(dict(group=g, seconds=dt, reads_per_s=n_reads / dt, frames_per_s=float(fr.sum()) / dt, ms_search=float(res.ms_search),
                     frames_per_read=float(fr.mean()), max_frames=int(fr.max()), limit_reads=int((recs["flags"] & 1).sum()),
                     retry_reads=int((recs["flags"] & 2).sum()), mapped=float(recs["mapped"].mean())))
    print("variant", runs[-1], flush=True)
out["runs"] = runs
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cfg4_frames_%d.npz" % (genome_bp // 1_000_000)), frames=frames.astype(np.uint32),
                    lengths=np.diff(off).astype(np.uint16))
print(json.dumps(out))
