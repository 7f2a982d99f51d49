"""`-m gpu`, needs at least two GPUs (skipped otherwise): the multi-GPU product paths.
  * torchrun-style: two ranks, NCCL, index blob broadcast from rank 0, every chunk sharded with sharding.map_sharded and
    merged in input order on rank 0 — identical to the single-GPU result (dispatcher.rs:341-379 replaced for one box)
  * in-process: Mapper.clone(device=1) (one peer copy of the blob) and `mapad_b200.cli map --gpus 2` — BAM byte-identical
    to the --gpus 1 run apart from the @PG command line"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from compare import compare_results
from helpers import product_params, random_genome, simulate_reads
from ref_cases import cli_params
from mapad_b200 import abi, api, sharding
rank = int(sys.argv[3])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2, device_id=torch.device("cuda", rank))
genome = random_genome(400000, seed=5)
contigs = [("chr1", genome[:150000]), ("chr2", genome[150000:])]
params = product_params(cli_params("single_stranded"))
index = api.Index.build(contigs) if rank == 0 else None
if rank == 0:
    m0 = api.Mapper(index, params, device=0)
    meta, _, nbytes = m0.export_index()
    hdr = [meta, nbytes]
else:
    hdr = [None, None]
dist.broadcast_object_list(hdr, src=0)
meta, nbytes = hdr
blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
if rank == 0:
    m0.copy_index_to(blob.data_ptr(), nbytes)
    torch.cuda.synchronize()
dist.broadcast(blob, src=0)
torch.cuda.synchronize()
mapper = api.Mapper.from_device_blob(meta, blob.data_ptr(), nbytes, index, params, device=rank)
seqs, quals = simulate_reads(genome, 1501, (25, 70), seed=9)
packed = abi.pack_reads(seqs, quals)
seeds = np.arange(1501, dtype=np.uint32) * 3
merged = sharding.map_sharded(dist, lambda sh, sd: mapper.map_batch(seeds=sd, packed=sh, want_hits=True), packed, seeds)
if rank == 0:
    whole = m0.map_batch(seeds=seeds, packed=packed, want_hits=True)
    compare_results(whole, merged, label_a="single", label_b="sharded")
    print("SHARDED_OK", len(merged))
dist.barrier()
mapper.close()
dist.destroy_process_group()
"""


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_rank_nccl_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, ROOT, str(port), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_OK 1501" in outs[0], outs[0]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_cli_two_gpus_identical_bam(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bamio import read_bam
    from helpers import random_genome, simulate_reads
    genome = random_genome(300_000, seed=21)
    fa = tmp_path / "g.fa"
    with open(fa, "w") as f:
        for n, s in (("chrA", genome[:100_000]), ("chrB", genome[100_000:])):
            f.write(">%s\n" % n)
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    seqs, quals = simulate_reads(genome, 2000, (30, 70), seed=5)
    fq = tmp_path / "r.fastq"
    with open(fq, "w") as f:
        for i, (s, q) in enumerate(zip(seqs, quals)):
            f.write("@read%d\n%s\n+\n%s\n" % (i, s.decode(), bytes(c + 33 for c in q).decode()))
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("o%d.bam" % gpus)
        cmd = [sys.executable, "-m", "mapad_b200.cli", "map", "-r", str(fq), "-g", str(fa), "-o", str(out), "--library", "single_stranded",
               "-p", "0.03", "-f", "0.5", "-t", "0.5", "-d", "0.02", "-s", "1.0", "-i", "0.001", "-x", "0.5", "--batch_size", "450",
               "--seed", "7", "--gpus", str(gpus), "--inflight", "2"]
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(read_bam(str(out)))
    (_, refs1, recs1), (_, refs2, recs2) = outs
    assert refs1 == refs2 and len(recs1) == len(recs2) == 2000
    assert [x["name"] for x in recs2] == ["read%d" % i for i in range(2000)]
    for a, b in zip(recs1, recs2):
        assert a == b, a["name"]
