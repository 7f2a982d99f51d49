#!/usr/bin/env python3
"""hg19-scale tuning probe: builds the cfg4 index once, then for every variant maps K chunks concurrently over K handles
(the bench's pipelining) and reports seconds, reads/s, frames/s, deferred reads.  Variants are NAME[:ENV=VALUE,...].
Measurement tool — the numbers it prints are not bench values.
Usage: python tools/probe_cfg4.py [K=8] [reads_per_chunk=25000] variant..."""
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapad_b200 import abi, api, workloads  # noqa: E402
from mapad_b200.specs import cli_spec as cli_params, product_params  # noqa: E402  (oracle-free)

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 25_000
variants = sys.argv[3:] or ["g8:MAPAD_GROUP=8"]
genome_bp = int(float(os.environ.get("PROBE_GENOME_BP", "3.1e9")))
genome = workloads.random_genome_array(genome_bp, seed=42)
t0 = time.time()
index = api.Index.build(workloads.split_contigs(genome, 24), seed=1234, device=0 if genome_bp > 500_000_000 else None)
print("index built in %.1f s" % (time.time() - t0), flush=True)
params = product_params(cli_params("single_stranded"))
chunks = [workloads.simulate_batch(genome, n_reads, (25, 100), seed=1004000 + 20000 + c) for c in range(K)]
structs = [api.make_reads(c[0], c[1], c[2], np.arange(n_reads, dtype=np.uint32)) for c in chunks]
del genome
rows = []
for v in variants:
    parts = v.split(":")
    env = dict(kv.split("=", 1) for kv in parts[1].split(",")) if len(parts) > 1 and parts[1] else {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    api.plan_handles(0, K + 1)
    first = api.Mapper(index, params)
    mappers = [first.clone() for _ in range(K)]
    stats = {}

    def work(i):
        res = mappers[i].map_raw(structs[i][0], 0)
        recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
        sig = hashlib.sha1()
        for f in ("mapped", "tid", "pos", "strand", "mapq", "alignment_score", "nm", "x0", "x1", "best_lower", "best_size", "frames_popped"):
            sig.update(np.ascontiguousarray(recs[f]).tobytes())
        stats[i] = (int(recs["frames_popped"].astype(np.int64).sum()), int(((recs["flags"] & 2) != 0).sum()), time.time(), sig.hexdigest()[:12])

    t = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.time() - t
    row = dict(variant=parts[0], env=env, seconds=round(dt, 2), reads_per_s=K * n_reads / dt, frames_per_s=sum(s[0] for s in stats.values()) / dt,
               deferred=sum(s[1] for s in stats.values()), done_s=sorted(round(s[2] - t, 1) for s in stats.values()),
               # records of every chunk hashed (best hit, position, MAPQ, scores, frame counts): must not change between variants
               signature=hashlib.sha1("".join(stats[i][3] for i in range(K)).encode()).hexdigest()[:12])
    rows.append(row)
    print(json.dumps(row), flush=True)
    for m in mappers:
        m.close()
    first.close()
    for k, val in old.items():
        if val is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = val
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "probe_cfg4.json"), "w"), indent=1)
