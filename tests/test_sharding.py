"""N > 1 host logic on CPU: world_size-2 gloo process group; each rank maps its contiguous range of a chunk
(with the CPU emulation of the device code standing in for the GPU), rank 0 merges in input order; the index
blob broadcast is exercised with a byte tensor."""
import os
import socket
import subprocess
import sys

import numpy as np

from mapad_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, pickle
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from compare import compare_results
from helpers import product_params, random_genome, simulate_reads
from ref_cases import cli_params
from mapad_b200 import abi, api, sharding
from emu import emu
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
genome = random_genome(40000, seed=5)
index = api.Index.build([("chr1", genome[:15000]), ("chr2", genome[15000:])])
params = product_params(cli_params("single_stranded"))
# "index broadcast": rank 0 owns the blob bytes, everybody ends up with the same tensor
blob = torch.arange(1 << 16, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(1 << 16, dtype=torch.uint8)
meta = sharding.broadcast_index_blob(dist, b"meta-pod" if rank == 0 else None, blob, src=0)
assert meta == b"meta-pod" and int(blob[259]) == 3
seqs, quals = simulate_reads(genome, 101, (25, 60), seed=9)
packed = abi.pack_reads(seqs, quals)
seeds = np.arange(101, dtype=np.uint32) * 3
merged = sharding.map_sharded(dist, lambda sh, sd: emu.map_batch(index, params, seeds=sd, packed=sh), packed, seeds)
if rank == 0:
    whole = emu.map_batch(index, params, seeds=seeds, packed=packed)
    compare_results(whole, merged, label_a="single", label_b="sharded")
    print("SHARDED_OK", len(merged))
dist.barrier()
dist.destroy_process_group()
"""


def test_shard_ranges():
    for n in (0, 1, 7, 100, 250000):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, ROOT, str(port), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_OK 101" in outs[0], outs[0]
