#!/usr/bin/env python3
"""(Test infrastructure: uses the CPU oracle as the checker / CPU column, hence under tests/.)
BASELINE cfg5: backtracking stress sweep on the hg19-scale index — `-p` x read length, reporting reads/s, search-frame
counts and the achieved fraction of the memory roofline per cell, with the CPU restatement beside it on a small sample.
Cells are time-boxed (chunks of reads are mapped until the cell's budget is used or its read target is reached), because
the work per read spans five orders of magnitude over the grid.  Measurement tool — writes gpurun_out/cfg5_sweep.json and a
markdown table; the numbers are not bench values.
Usage: python tests/tools/run_cfg5.py [genome_bp=3.1e9] [seconds_per_cell=15] [max_reads_per_cell=100000] [cpu_reads=200]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapad_b200 import abi, api, workloads  # noqa: E402
from helpers import oracle_params, product_params  # noqa: E402
from ref_cases import cli_params  # noqa: E402

genome_bp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
cell_s = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
max_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
cpu_reads = int(sys.argv[4]) if len(sys.argv) > 4 else 200
PS = [float(x) for x in os.environ.get("CFG5_P", "0.01,0.02,0.03,0.04,0.05,0.06").split(",")]
LS = [int(x) for x in os.environ.get("CFG5_L", "25,35,50,75,100,150").split(",")]
CHUNK = int(os.environ.get("CFG5_CHUNK", "5000"))
INFLIGHT = int(os.environ.get("CFG5_INFLIGHT", "4"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

genome = workloads.random_genome_array(genome_bp, seed=42)
t0 = time.time()
index = api.Index.build(workloads.split_contigs(genome, 24), seed=1234, device=0 if genome_bp > 500_000_000 else None)
print("index built in %.1f s" % (time.time() - t0), flush=True)
from oracle import oracle as ora  # noqa: E402
a = index.arrays()
oix = ora.OracleIndex.from_arrays(a["bwt"], a["sa_sample"], a["sa_rate"], a["extra_rows"], a["contigs"], a["orig_pos"], a["orig_sym"])
del a
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
api.plan_handles(0, INFLIGHT + 1)
spec = dict(cli_params("single_stranded"))
first = api.Mapper(index, product_params(spec))
mappers = [first.clone() for _ in range(INFLIGHT)]
rows = []
for p in PS:
    spec["bound"] = ("discrete", p, 0.02)
    params = product_params(spec)
    for m in mappers:
        m.set_params(params)
    k_of = {L: api.allowed_mismatches(params, L) for L in LS}
    for L in LS:
        chunks = [workloads.simulate_batch(genome, CHUNK, (L, L), seed=5000 + 100 * L + c) for c in range(INFLIGHT)]
        structs = [api.make_reads(c[0], c[1], c[2], np.arange(CHUNK, dtype=np.uint32)) for c in chunks]
        stats = dict(reads=0, frames=0, mapped=0, limit=0, max_frames=0)
        lock = threading.Lock()
        t_start = time.time()

        def work(i):
            while True:
                with lock:
                    if stats["reads"] + CHUNK > max_reads and stats["reads"] > 0:
                        return
                    if time.time() - t_start > cell_s and stats["reads"] > 0:
                        return
                    stats["reads"] += CHUNK  # claimed
                res = mappers[i].map_raw(structs[i][0], 0)
                recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
                fr = recs["frames_popped"].astype(np.int64)
                with lock:
                    stats["frames"] += int(fr.sum()); stats["mapped"] += int(recs["mapped"].sum())
                    stats["limit"] += int((recs["flags"] & 1).sum()); stats["max_frames"] = max(stats["max_frames"], int(fr.max()))

        th = [threading.Thread(target=work, args=(i,)) for i in range(INFLIGHT)]
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.time() - t_start
        # CPU restatement on the first reads of chunk 0
        n_cpu = min(cpu_reads, CHUNK)
        seq, qual, off = chunks[0]
        sub = (seq[: int(off[n_cpu])], qual[: int(off[n_cpu])], off[: n_cpu + 1])
        tc = time.time()
        ora.map_batch(oix, oracle_params(spec), None, None, seeds=np.arange(n_cpu, dtype=np.uint32), n_threads=os.cpu_count(), want_hits=False, packed=sub)
        cpu_rps = n_cpu / (time.time() - tc)
        row = dict(p=p, L=L, k=float(k_of[L]), reads=stats["reads"], seconds=round(dt, 2), reads_per_s=stats["reads"] / dt,
                   frames_per_read=stats["frames"] / stats["reads"], max_frames=stats["max_frames"], reads_at_limit=stats["limit"],
                   mapped_fraction=stats["mapped"] / stats["reads"], frames_per_s=stats["frames"] / dt,
                   roofline_frac=128.0 * stats["frames"] / dt / 1e9 / hbm, cpu_reads_per_s=cpu_rps, cpu_threads=os.cpu_count(), cpu_sample=n_cpu)
        rows.append(row)
        print(json.dumps(row), flush=True)
json.dump(dict(genome_bp=genome_bp, chunk=CHUNK, inflight=INFLIGHT, seconds_per_cell=cell_s, rows=rows),
          open(os.path.join(ROOT, "gpurun_out", "cfg5_sweep.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "cfg5_sweep.md"), "w") as f:
    f.write("| -p | L | k(L) | reads | reads/s (GPU) | frames/read | max frames | reads at limit | frames/s | roofline frac | reads/s (CPU, %d thr) |\n" % (os.cpu_count() or 1))
    f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %.2f | %d | %.0f | %d | %.0f | %.0f | %d | %d | %.3g | %.4f | %.1f |\n" % (
            r["p"], r["L"], r["k"], r["reads"], r["reads_per_s"], r["frames_per_read"], r["max_frames"], r["reads_at_limit"], r["frames_per_s"],
            r["roofline_frac"], r["cpu_reads_per_s"]))
