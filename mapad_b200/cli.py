"""`mapad map`-compatible driver on top of the C ABI: FASTQ / FASTQ.GZ in, BAM out.

    python -m mapad_b200.cli map -r reads.fastq.gz -g genome.fa -o out.bam --library single_stranded \
        -p 0.03 -f 0.5 -t 0.5 -d 0.02 -s 1.0 -i 0.001 -x 0.5

    python -m mapad_b200.cli index -g genome.fa [--seed 1234]      # writes genome.fa.{tbw,tle,toc,trt,tsa,tpi,tos}

Flag names, defaults and validators follow /root/reference/src/main.rs:30-303 (`-p` or `-c`/`-e`, probabilities checked
to lie in [0, 1], `-v`, `--seed`; `--threads` / `--port` are accepted and unused).  `map` loads the seven index files next to the
FASTA when they exist (as the reference does) and otherwise indexes the FASTA in memory.  Reads come from FASTQ,
FASTQ.GZ or BAM (flags, auxiliary fields and the @PG/@RG/@CO header lines of a BAM input are carried over like in
create_bam_header / create_bam_record).  Differences: CRAM is not read, and the per-read XD:f timing tag is not written.  Several chunks are kept in flight (--inflight) so that the
straggler reads of one chunk overlap with the next; records are written in input order.
"""
import argparse
import ctypes as C
import os
import queue
import sys
import threading

import numpy as np

from . import abi, api


def read_fasta(path):
    contigs, name, parts = [], None, []
    opener = open
    if path.endswith(".gz"):
        import gzip
        opener = gzip.open
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    contigs.append((name, "".join(parts)))
                name, parts = line[1:].split()[0] if line[1:].split() else "", []
            else:
                parts.append(line.strip())
    if name is not None:
        contigs.append((name, "".join(parts)))
    return contigs


def prob(raw):
    """parse_validate_prob (src/main.rs:33-39): an f32 in [0, 1]."""
    try:
        v = float(raw)
    except ValueError:
        raise argparse.ArgumentTypeError("not a number: %r" % raw)
    if not 0.0 <= v <= 1.0:
        raise argparse.ArgumentTypeError("%r is not a probability (0 <= value <= 1)" % raw)
    return v


def build_parser():
    ap = argparse.ArgumentParser(prog="mapad_b200")
    # global options of the reference (src/main.rs:63-98); --threads and --port have no meaning for the GPU path and are
    # accepted so that existing command lines keep working
    glob = argparse.ArgumentParser(add_help=False)
    glob.add_argument("-v", action="count", default=0, help="Sets the level of verbosity")
    glob.add_argument("--threads", type=int, default=1, help="accepted for compatibility (host threads only compress BAM blocks)")
    glob.add_argument("--port", type=int, default=3130, help="accepted for compatibility (no TCP dispatcher: one box, --gpus N)")
    glob.add_argument("--seed", type=int, default=1234, help="Seed for the random number generator")
    sub = ap.add_subparsers(dest="cmd", required=True)
    m = sub.add_parser("map", help="Maps reads to a genome", parents=[glob])
    m.add_argument("-r", "--reads", required=True)
    m.add_argument("-g", "--reference", required=True, help="FASTA file of the genome")
    m.add_argument("-o", "--output", required=True)
    # -p (Discrete bound) and -c / -e (Continuous bound) exclude each other, one is required (src/main.rs:136-160, 456-476)
    m.add_argument("-p", dest="poisson_prob", type=prob, default=None,
                   help="Minimum probability of the number of mismatches under `-D` base error rate")
    m.add_argument("-c", dest="as_cutoff", type=float, default=None, help="Per-base average alignment score cutoff (-c > AS / read_len^e ?)")
    m.add_argument("-e", dest="as_cutoff_exponent", type=float, default=1.0, help="Exponent applied to the read length (ignored without -c)")
    m.add_argument("-l", "--library", choices=["single_stranded", "double_stranded"], required=True)
    m.add_argument("-f", dest="five_prime_overhang", type=prob, required=True)
    m.add_argument("-t", dest="three_prime_overhang", type=prob, default=None)
    m.add_argument("-d", dest="ds_deamination_rate", type=prob, required=True)
    m.add_argument("-s", dest="ss_deamination_rate", type=prob, required=True)
    m.add_argument("-D", dest="divergence", type=prob, default=0.02)
    m.add_argument("-i", dest="indel_rate", type=prob, required=True)
    m.add_argument("-x", dest="gap_extension_penalty", type=prob, default=1.0)
    m.add_argument("--batch_size", type=int, default=250000)
    m.add_argument("--ignore_base_quality", action="store_true")
    m.add_argument("--gap_dist_ends", type=int, default=5)
    m.add_argument("--max_num_gaps_open", type=int, default=2)
    m.add_argument("--no_search_limit_recovery", action="store_true")
    m.add_argument("--force_overwrite", action="store_true")
    m.add_argument("-R", "--read_group", default=None, help="read group ID added to every record")
    m.add_argument("--device", type=int, default=0, help="first CUDA device")
    m.add_argument("--gpus", type=int, default=1, help="shard every chunk of reads over this many GPUs of the box (devices --device ..)")
    m.add_argument("--inflight", type=int, default=8, help="chunks kept in flight per GPU")
    ix = sub.add_parser("index", help="Indexes a genome file", parents=[glob])
    ix.add_argument("-g", "--reference", required=True, help="FASTA file of the genome")
    ix.add_argument("--device", type=int, default=None, help="sort suffixes on this CUDA device (tie-free synthetic texts only; default: host SA-IS)")
    return ap


def build_index(path, seed, device):
    contigs = read_fasta(path)
    # device=None: host SA-IS.  The device suffix sorter (--device) only finishes tie-free texts (synthetic i.i.d. genomes);
    # it is never chosen automatically for FASTA input, because real genomes make it fall back to the host after a wasted pass.
    return api.Index.build(contigs, seed=seed, device=device)


def run_index(a):
    build_index(a.reference, a.seed, a.device).save(a.reference)
    return 0


def params_from_args(a):
    if a.library == "single_stranded" and a.three_prime_overhang is None:
        raise SystemExit("-t is required for --library single_stranded")
    if (a.poisson_prob is None) == (a.as_cutoff is None):
        raise SystemExit("exactly one of -p (Poisson mismatch bound) and -c (alignment-score cutoff) is required")
    P = api.params_from_cli(library=a.library, p=a.poisson_prob if a.poisson_prob is not None else 0.03, f=a.five_prime_overhang,
                            t=a.three_prime_overhang or 0.0, d=a.ds_deamination_rate, s=a.ss_deamination_rate, D=a.divergence,
                            i=a.indel_rate, x=a.gap_extension_penalty, gap_dist_ends=a.gap_dist_ends,
                            max_num_gaps_open=a.max_num_gaps_open, ignore_base_quality=a.ignore_base_quality,
                            no_search_limit_recovery=a.no_search_limit_recovery)
    if a.as_cutoff is not None:  # Continuous::new(-c * -1.0, -e, representative mismatch penalty) (src/main.rs:463-475)
        P.bound_kind = abi.BOUND_CONTINUOUS
        P.cutoff, P.exponent = -float(a.as_cutoff), float(a.as_cutoff_exponent)
    return P


class _Shard:
    """A contiguous range [lo, hi) of a chunk's reads as its own mapad_reads view (pointers advanced, nothing copied)."""

    def __init__(self, R, names, noff, flags, lo, hi, seeds):
        self.R = abi.Reads()
        self.R.n_reads = hi - lo
        self.R.seq, self.R.qual = R.seq, R.qual
        self.R.offsets = R.offsets + 8 * lo
        self.seeds = np.ascontiguousarray(seeds[lo:hi])
        self.R.seeds = self.seeds.ctypes.data
        self.names = names
        self.noff = C.c_void_p(noff.value + 8 * lo) if noff.value else noff
        self.flags = C.c_void_p(flags.value + 2 * lo) if flags.value else flags
        self.lo, self.hi = lo, hi


def run_map(a, argv):
    params = params_from_args(a)
    if os.path.exists(a.reference + ".tbw"):
        index = api.Index.load(a.reference)
    else:
        print("no index files next to %s: indexing in memory" % a.reference, file=sys.stderr)
        index = build_index(a.reference, a.seed, None)
    # one set of `inflight` handles per GPU; the index is re-laid-out once and replicated with one peer copy per further GPU
    devices = [a.device + d for d in range(max(1, a.gpus))]
    inflight = max(1, a.inflight)
    for d in devices:
        api.plan_handles(d, inflight)
    firsts = [api.Mapper(index, params, device=devices[0])]
    firsts += [firsts[0].clone(device=d) for d in devices[1:]]
    pools = [[f] + [f.clone() for _ in range(inflight - 1)] for f in firsts]
    chunks = api.ReadChunks(a.reads, a.batch_size)
    writer = api.BamWriter(a.output, index, command_line=" ".join(argv), read_group_id=a.read_group, force_overwrite=a.force_overwrite,
                           src_header_text=chunks.header_text)
    rng = np.random.default_rng(a.seed)
    state = dict(next=0, mapped=0, reads=0, error=None)
    lock = threading.Condition()
    queues = [queue.Queue(maxsize=inflight) for _ in devices]
    pending = {}  # chunk k -> [shards still to be written, chunk handle]

    def worker(mp, q):
        # every shard has a sequence number; results (valid until this handle's next call) are written in that order.
        # After a failure the workers keep draining their queues without mapping, so that nobody waits forever.
        while True:
            job = q.get()
            if job is None:
                return
            seqno, k, sh, aux, aoff = job
            try:
                res = None
                if state["error"] is None:
                    res = mp.map_raw(sh.R, 0)
                with lock:
                    while state["next"] != seqno and state["error"] is None:
                        lock.wait()
                    if state["error"] is None:
                        api._check(api.lib().mapad_bam_write_chunk_aux(writer.w, index.h, C.byref(sh.R), sh.names, sh.noff, sh.flags, aux, aoff, C.byref(res)))
                        recs = abi._as_array(res.records, res.n_reads, abi.RECORD_DTYPE)
                        state["mapped"] += int(recs["mapped"].sum())
                        state["reads"] += sh.hi - sh.lo
                        state["next"] = seqno + 1
                        lock.notify_all()
            except Exception as e:  # noqa: BLE001
                with lock:
                    if state["error"] is None:
                        state["error"] = e
                    lock.notify_all()
            finally:
                with lock:
                    pending[k][0] -= 1
                    if pending[k][0] == 0:
                        chunks.free(pending.pop(k)[1])

    threads = [threading.Thread(target=worker, args=(mp, queues[d])) for d in range(len(devices)) for mp in pools[d]]
    for t in threads:
        t.start()
    from .sharding import shard_range
    seqno = k = 0
    try:
        for R, names, noff, flags, n, ch in chunks:
            if state["error"] is not None:
                chunks.free(ch)
                break
            seeds = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
            aux, aoff = C.c_void_p(), C.c_void_p()
            api._check(api.lib().mapad_chunk_aux(ch, C.byref(aux), C.byref(aoff)))
            ranges = [shard_range(n, d, len(devices)) for d in range(len(devices))]
            ranges = [(d, lo, hi) for d, (lo, hi) in enumerate(ranges) if hi > lo]
            with lock:
                pending[k] = [len(ranges), ch]
            for d, lo, hi in ranges:
                sh = _Shard(R, names, noff, flags, lo, hi, seeds)
                sh_aoff = C.c_void_p(aoff.value + 8 * lo) if aoff.value else aoff
                queues[d].put((seqno, k, sh, aux, sh_aoff))
                seqno += 1
            k += 1
    finally:
        for d in range(len(devices)):
            for _ in pools[d]:
                queues[d].put(None)
        for t in threads:
            t.join()
    err = state["error"]
    try:
        writer.close()
    except Exception as e:  # noqa: BLE001
        err = err or e
    chunks.close()
    for pool in pools:
        for mp in pool[1:]:
            mp.close()
    for f in reversed(firsts):
        f.close()
    if err is not None:
        print("mapad_b200: mapping failed: %s" % err, file=sys.stderr)
        return 1
    if a.v:
        print("mapped %d of %d reads (%d skipped on input)" % (state["mapped"], state["reads"], chunks.skipped), file=sys.stderr)
    return 0


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    a = build_parser().parse_args(argv[1:])
    if a.cmd == "map":
        return run_map(a, argv)
    if a.cmd == "index":
        return run_index(a)
    return 2


if __name__ == "__main__":
    sys.exit(main())
