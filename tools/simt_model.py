"""Host-side model of the lane efficiency of k_search_pool's loop structure (see tools/simt_model.cpp).  CPU only.

    python tools/simt_model.py [n_warps=40] [genome_bp=2000000]
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapad_b200 import abi, api, workloads  # noqa: E402
from helpers import product_params  # noqa: E402
from ref_cases import cli_params  # noqa: E402


def build():
    csrc = os.path.join(ROOT, "mapad_b200", "csrc")
    out = os.path.join(ROOT, "tools", "libsimt_model.so")
    srcs = [os.path.join(ROOT, "tools", "simt_model.cpp")] + [os.path.join(csrc, f) for f in ("host_index.cpp", "host_params.cpp", "dev_index_build.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-march=x86-64-v3", "-shared", "-o", out] + srcs)
    return out


def main():
    n_warps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    gbp = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    L = C.CDLL(build())
    L.simt_model.restype = C.c_int
    L.simt_model.argtypes = [C.c_void_p, C.POINTER(abi.Params), C.POINTER(abi.Reads), C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    cfg = workloads.CONFIGS["cfg3"]
    genome = workloads.random_genome_array(gbp, seed=42)
    index = api.Index.build(workloads.split_contigs(genome, 2), seed=1234)
    params = product_params(cli_params(cfg["library"]))
    tot = np.zeros(8)
    heap = np.zeros(8)
    reads_per_warp = 32 * 12   # each lane maps ~12 reads, like a thread of the pool lane maps tens of reads
    for w in range(n_warps):
        seq, qual, off = workloads.simulate_batch(genome, reads_per_warp, cfg["len_range"], seed=5000 + w, library=cfg["library"])
        R, keep = api.make_reads(seq, qual, off, np.arange(reads_per_warp, dtype=np.uint32))
        out = (C.c_double * 8)()
        h8 = (C.c_double * 8)()
        rc = L.simt_model(index.h, C.byref(params), C.byref(R), 1 << 17, out, h8)
        assert rc == 0, rc
        tot += np.array(list(out))
        heap += np.array(list(h8))
    iters, frames, cur, flat, useful, cur_tr, cur_push, skipped = tot
    print(json.dumps(dict(
        warps=n_warps, reads=n_warps * reads_per_warp, frames=int(frames), skipped_reads=int(skipped),
        lanes_busy_per_iteration=round(frames / iters, 2),
        round_trips_per_iteration=dict(current=round(cur / iters, 2), flat=round(flat / iters, 2), of_which_current_trickle=round(cur_tr / iters, 2),
                                       of_which_current_push=round(cur_push / iters, 2)),
        useful_round_trips_per_frame=round(useful / frames, 2),
        lane_efficiency=dict(current=round(useful / (32 * cur), 3), flat=round(useful / (32 * flat), 3)),
        speedup_bound_flat_vs_current=round(cur / flat, 2),
        heap_accesses_per_frame=round(heap[6] / frames, 1),
        heap_cold_sectors_and_lines_per_frame=dict(  # heap entries below the top 31 only; 32-byte sectors / 64-byte lines
            linear_0_based=[round(heap[0] / frames, 2), round(heap[1] / frames, 2)],
            linear_1_based=[round(heap[2] / frames, 2), round(heap[3] / frames, 2)],
            family=[round(heap[4] / frames, 2), round(heap[5] / frames, 2)]))))


if __name__ == "__main__":
    main()
