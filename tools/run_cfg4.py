#!/usr/bin/env python3
"""hg19-scale check (BASELINE cfg4): builds the 3.1 Gbp synthetic index with the device suffix sorter, verifies the
CUDA path against the oracle on a sample of reads (positions beyond 2^32, wide 64 B occ blocks, HBM-resident index)
and measures reads/s.  Usage: python tools/run_cfg4.py [genome_bp] [n_reads] [n_parity]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapad_b200 import abi, api, workloads  # noqa: E402
from compare import compare_results  # noqa: E402
from helpers import oracle_params, product_params  # noqa: E402
from ref_cases import cli_params  # noqa: E402

genome_bp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000
n_parity = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
out = dict(genome_bp=genome_bp)
t = time.time()
genome = workloads.random_genome_array(genome_bp, seed=42)
out["genome_s"] = time.time() - t
t = time.time()
index = api.Index.build(workloads.split_contigs(genome, 24), device=0)
out["index_build_s"] = time.time() - t
print("index built", out, flush=True)
spec = cli_params("single_stranded")
params = product_params(spec)
t = time.time()
mapper = api.Mapper(index, params)
out["relayout_upload_s"] = time.time() - t
out["index_bytes_hbm"] = mapper.export_index()[2]
chunks = [workloads.simulate_batch(genome, n_reads, (25, 100), seed=1004000 + i) for i in range(4)]
# parity on a sample, oracle sharing the same index arrays
if n_parity:
    from oracle import oracle as ora
    a = index.arrays()
    oix = ora.OracleIndex.from_arrays(a["bwt"], a["sa_sample"], a["sa_rate"], a["extra_rows"], a["contigs"], a["orig_pos"], a["orig_sym"])
    del a
    seq, qual, off = chunks[0]
    sub = (seq[: int(off[n_parity])], qual[: int(off[n_parity])], off[: n_parity + 1])
    seeds = np.arange(n_parity, dtype=np.uint32)
    t = time.time()
    want = ora.map_batch(oix, oracle_params(spec), None, None, seeds=seeds, n_threads=os.cpu_count(), want_hits=True, packed=sub)
    out["oracle_reads_per_s"] = n_parity / (time.time() - t)
    got = mapper.map_batch(seeds=seeds, want_hits=True, packed=sub)
    compare_results(want, got)
    out["parity_reads"] = n_parity
    out["max_absolute_pos"] = int(got.records["absolute_pos"].max())
    del oix
    print("parity ok", out, flush=True)
# throughput: 4 chunks in flight over 4 handles
import threading
import torch
mappers = [mapper] + [mapper.clone() for _ in range(3)]
rs = [api.make_reads(c[0], c[1], c[2], np.arange(len(c[2]) - 1, dtype=np.uint32)) for c in chunks]
for m, r in zip(mappers, rs):
    m.map_raw(r[0], 0)  # warm-up
torch.cuda.synchronize()
t = time.time()
stats = {}
def work(i):
    res = mappers[i].map_raw(rs[i][0], 0)
    recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
    stats[i] = (int(recs["frames_popped"].astype(np.int64).sum()), int(recs["mapped"].sum()), int(recs["d_ext_steps"].astype(np.int64).sum()))
th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
[x.start() for x in th]; [x.join() for x in th]
torch.cuda.synchronize()
dt = time.time() - t
out["reads_per_s_e2e_4_chunks_in_flight"] = 4 * n_reads / dt
out["seconds_for_%d_reads" % (4 * n_reads)] = dt
out["frames_popped_per_read"] = sum(s[0] for s in stats.values()) / (4 * n_reads)
out["mapped_fraction"] = sum(s[1] for s in stats.values()) / (4 * n_reads)
out["d_ext_steps_per_read"] = sum(s[2] for s in stats.values()) / (4 * n_reads)
print(json.dumps(out))
