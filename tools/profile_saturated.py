"""Launches ONE k_search_group over a chunk of reads on the cfg3 index with every resident group busy and a fixed number
of expansions per group (MAPAD_PROFILE_ITERS), i.e. the saturated phase only, short enough for ncu:

    MAPAD_GROUP=8 ncu --set full --clock-control none --import-source on -k regex:k_search_group -c 1 -o gpurun_out/r2_g8 \\
        python tools/profile_saturated.py [min_len max_len]     # default 50 50 (cfg3); 86 100 = heavy reads, deep heaps

The batch is discarded by the library (MAPAD_ELIMIT) — this is a profiling aid, not a product path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MAPAD_WS_BYTES", str(60 << 30))
os.environ.setdefault("MAPAD_PROFILE_ITERS", "3000")
from mapad_b200 import api, workloads  # noqa: E402
from mapad_b200.specs import cli_spec as cli_params, product_params  # noqa: E402  (oracle-free)

cfg = workloads.CONFIGS[os.environ.get("PROFILE_WORKLOAD", "cfg3")]  # cfg4: hg19 scale (index outside L2, deep heaps with enough MAPAD_PROFILE_ITERS)
genome = workloads.random_genome_array(cfg["genome_bp"], seed=42)
index = api.Index.build(workloads.split_contigs(genome, cfg["n_contigs"]), seed=1234, device=0)
spec = dict(cli_params(cfg["library"]))
if os.environ.get("MAPAD_PROFILE_LIMITS"):  # e.g. "20000,100000": small STACK_LIMIT / EDIT_TREE_LIMIT -> many reads in limit recovery
    spec["limits"] = tuple(int(x) for x in os.environ["MAPAD_PROFILE_LIMITS"].split(","))
mapper = api.Mapper(index, product_params(spec), device=0)
len_range = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else cfg["len_range"]
n_reads = int(os.environ.get("PROFILE_READS", "250000"))
seq, qual, off = workloads.simulate_batch(genome, n_reads, len_range, seed=79, library=cfg["library"])
del genome
R, keep = api.make_reads(seq, qual, off, np.arange(n_reads, dtype=np.uint32))
try:
    mapper.map_raw(R, 0)
    print("unexpected: the batch completed")
except api.MapadError as e:
    print("profiled launch done (batch discarded as intended):", e)
