#!/usr/bin/env python3
"""Helper PROCESS of bench.py's reference arm: builds the index of a BASELINE workload with the product's indexer (host
SA-IS, or the device suffix sorter beyond 0.5 Gbp) and writes its arrays as .npy files, so that the process timing the
CPU restatement only ever loads oracle/libmapad_oracle.so.  Usage: python tools/index_arrays.py <workload> <out_dir>"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mapad_b200 import api, workloads  # noqa: E402


def main():
    name, out = sys.argv[1], sys.argv[2]
    cfg = workloads.CONFIGS[name]
    os.makedirs(out, exist_ok=True)
    genome = workloads.random_genome_array(cfg["genome_bp"], seed=42)
    dev = int(os.environ.get("LOCAL_RANK", "0")) if cfg["genome_bp"] > 500_000_000 else None
    index = api.Index.build(workloads.split_contigs(genome, cfg["n_contigs"]), seed=1234, device=dev)
    a = index.arrays()
    for k in ("bwt", "sa_sample", "extra_rows", "orig_pos", "orig_sym"):
        np.save(os.path.join(out, k + ".npy"), a[k])
    json.dump(dict(n=a["n"], sa_rate=a["sa_rate"], contigs=a["contigs"], less=a["less"], sentinel_rows=a["sentinel_rows"]),
              open(os.path.join(out, "meta.json"), "w"))


if __name__ == "__main__":
    main()
