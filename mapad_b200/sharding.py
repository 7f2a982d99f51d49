"""Single-box multi-GPU plumbing of the hot path: one process per GPU, the index replicated
per GPU (re-laid-out once on rank 0, then ONE broadcast of the device blob — NCCL over NVLink on GPUs, gloo in
the CPU tests), every chunk of reads split into contiguous per-rank ranges, no data-path collective, results
merged back into input order on rank 0.  Replaces the reference's TCP dispatcher / worker farm
(/root/reference/src/distributed/dispatcher.rs:103-338, worker.rs:45-215) for single-box runs; like there, the
unit that travels is a chunk of reads out (TaskSheet, input_chunk_reader.rs:247-253) and per-read results back
(ResultSheet, distributed/mod.rs:22-26).
"""
import numpy as np

from . import abi


def shard_range(n_reads, rank, world):
    """Contiguous range of reads [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_reads), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_packed(packed, rank, world, seeds=None):
    """(seq, qual, offsets[, seeds]) of the whole chunk -> the same for this rank's range (offsets re-based)."""
    seq, qual, offsets = packed
    lo, hi = shard_range(len(offsets) - 1, rank, world)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    off = (offsets[lo : hi + 1] - offsets[lo]).astype(np.uint64)
    sd = None if seeds is None else np.ascontiguousarray(seeds[lo:hi], dtype=np.uint32)
    return (seq[b0:b1], qual[b0:b1], off), sd, (lo, hi)


def broadcast_index_blob(dist, meta_bytes, blob_tensor, src=0):
    """One collective at start-up: the opaque meta POD (as an object) and the index blob tensor."""
    box = [meta_bytes]
    dist.broadcast_object_list(box, src=src)
    dist.broadcast(blob_tensor, src=src)
    return box[0]


def _rebase(res, hit_base, op_base, cig_base, text_base):
    rec = res.records.copy()
    rec["cigar_off"] += np.uint32(cig_base)
    rec["md_off"] += np.uint32(text_base)
    rec["hit_off"] += np.uint32(hit_base)
    for a in range(2):
        rec["alts"][:, a]["cigar_off"] += np.uint32(cig_base)
        rec["alts"][:, a]["md_off"] += np.uint32(text_base)
    hits = res.hits.copy()
    if len(hits):
        hits["edit_off"] += np.uint32(op_base)
    return rec, hits


def merge_results(parts):
    """list[abi.BatchResult] in rank order -> one abi.BatchResult in input order (pools concatenated, offsets re-based)."""
    out = object.__new__(abi.BatchResult)
    recs, hits, ops, cig, text, xa = [], [], [], [], [], []
    hb = ob = cb = tb = 0
    for p in parts:
        r, h = _rebase(p, hb, ob, cb, tb)
        recs.append(r); hits.append(h); ops.append(p.edit_ops); cig.append(p.cigar); text.append(p.text)
        if p.xa is not None:
            xa.extend(p.xa)
        hb += len(p.hits); ob += len(p.edit_ops); cb += len(p.cigar); tb += len(p.text)
    out.records = np.concatenate(recs) if recs else np.zeros(0, abi.RECORD_DTYPE)
    out.hits = np.concatenate(hits) if hits else np.zeros(0, abi.HIT_DTYPE)
    out.edit_ops = np.concatenate(ops) if ops else np.zeros(0, abi.EDIT_OP_DTYPE)
    out.cigar = np.concatenate(cig) if cig else np.zeros(0, np.uint32)
    out.text = b"".join(text)
    out.xa = xa if xa else None
    out.timing = {}
    out.gpu_launches = sum(getattr(p, "gpu_launches", 0) for p in parts)
    return out


def map_sharded(dist, map_fn, packed, seeds=None, dst=0):
    """Every rank maps its contiguous range with `map_fn(packed_shard, seeds_shard) -> abi.BatchResult`; rank `dst`
    returns the merged result in input order, the others return None.  The only communication is the final
    gather of the per-rank results (host side, like the dispatcher's ResultSheet collection)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    shard, sd, _ = shard_packed(packed, rank, world, seeds)
    res = map_fn(shard, sd)
    payload = dict(records=res.records, hits=res.hits, edit_ops=res.edit_ops, cigar=res.cigar, text=res.text, xa=res.xa,
                   gpu_launches=getattr(res, "gpu_launches", 0))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    parts = []
    for g in gathered:
        p = object.__new__(abi.BatchResult)
        p.records, p.hits, p.edit_ops, p.cigar, p.text, p.xa = g["records"], g["hits"], g["edit_ops"], g["cigar"], g["text"], g["xa"]
        p.gpu_launches = g["gpu_launches"]
        parts.append(p)
    return merge_results(parts)
