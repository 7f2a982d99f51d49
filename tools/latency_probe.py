"""Per-frame latency and SIMT-efficiency probe for k_search_pool (GPU only; measurement tool, not a test).

Maps a sample of cfg3 reads once to learn each read's frame count, then times hand-made chunks in a single handle:
  one      1 copy of a heavy read                           -> dependent-chain latency per popped frame
  warp     32 copies (one warp, perfect lock step)          -> the same with 32 lanes active
  full     one copy per resident thread (148 SMs x 4 x 128) -> throughput ceiling of the kernel with perfect SIMT efficiency
  sorted   reads of similar size in every warp              -> divergence without the straggler drain
  random   the natural order                                -> what the bench sees inside one launch
Prints one JSON object.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RESIDENT = int(os.environ.get("MAPAD_PROBE_THREADS", str(148 * 4 * 128)))
os.environ.setdefault("MAPAD_POOL_THREADS", str(RESIDENT))
os.environ.setdefault("MAPAD_WS_BYTES", str(120 << 30))  # one handle: room for a full grid of medium-heavy reads

from mapad_b200 import abi, api, workloads  # noqa: E402
from helpers import product_params  # noqa: E402
from ref_cases import cli_params  # noqa: E402


def main():
    cfg = workloads.CONFIGS[os.environ.get("MAPAD_PROBE_WORKLOAD", "cfg3")]
    genome = workloads.random_genome_array(cfg["genome_bp"], seed=42)
    index = api.Index.build(workloads.split_contigs(genome, cfg["n_contigs"]), seed=1234, device=0)
    mapper = api.Mapper(index, product_params(cli_params(cfg["library"])), device=0)
    n = RESIDENT
    seq, qual, off = workloads.simulate_batch(genome, n, cfg["len_range"], seed=77, library=cfg["library"])

    def run(seq_, qual_, off_, label):
        R, keep = api.make_reads(seq_, qual_, off_, np.arange(len(off_) - 1, dtype=np.uint32))
        mapper.map_raw(R, 0)  # warm: buffers
        t0 = time.perf_counter()
        res = mapper.map_raw(R, 0)
        wall = time.perf_counter() - t0
        recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
        frames = int(recs["frames_popped"].astype(np.int64).sum())
        deferred = int(((recs["flags"] & 2) != 0).sum())
        out = dict(label=label, reads=len(off_) - 1, frames=frames, ms_search=round(float(res.ms_search), 3), wall_ms=round(wall * 1e3, 3),
                   us_per_frame_per_read=round(float(res.ms_search) * 1e3 / max(1, frames) * (len(off_) - 1), 3),
                   mframes_per_s=round(frames / max(1e-9, float(res.ms_search)) * 1e-3, 2), retry_lane_reads=deferred)
        print(json.dumps(out), flush=True)
        return recs["frames_popped"].astype(np.int64).copy()

    def subset(ids):
        ids = np.asarray(ids, dtype=np.int64)
        lens = (off[ids + 1] - off[ids]).astype(np.int64)
        o = np.zeros(len(ids) + 1, dtype=np.uint64)
        o[1:] = np.cumsum(lens)
        idx = np.repeat(off[ids].astype(np.int64) - o[:-1].astype(np.int64), lens) + np.arange(int(o[-1]), dtype=np.int64)
        return seq[idx], qual[idx], o

    frames = run(seq, qual, off, "random")
    order = np.argsort(frames)
    # a heavy read that stays inside the thread lane (tree nodes <= ~3 x frames; keep well below 131072 nodes)
    pick = lambda f: int(order[min(len(order) - 1, np.searchsorted(frames[order], f))])
    heavy, medium, median = pick(30000), pick(8000), int(order[len(order) // 2])
    print(json.dumps(dict(frames_median=int(frames[median]), frames_heavy=int(frames[heavy]), frames_medium=int(frames[medium]), frames_max=int(frames.max()),
                          frames_mean=float(frames.mean()), p99=int(np.percentile(frames, 99)))), flush=True)
    for label, rid, reps in (("heavy", heavy, (1, 32, 128)), ("medium", medium, (1, 32, RESIDENT)), ("median", median, (32, RESIDENT))):
        for copies in reps:
            got = run(*subset([rid] * copies), "%s_x%d" % (label, copies))
            assert (got == frames[rid]).all()
    run(*subset(order), "sorted")
    # sorted, without the heaviest 1 %: no stragglers, little divergence
    run(*subset(order[: int(len(order) * 0.99)]), "sorted_p99")
    rng = np.random.default_rng(1)
    run(*subset(rng.permutation(order[: int(len(order) * 0.99)])), "random_p99")


if __name__ == "__main__":
    main()
