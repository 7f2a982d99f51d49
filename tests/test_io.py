"""FASTQ(.GZ) reader and BAM writer on the CPU: the 17 reads of the reference's integration test
(tests/integration_tests.rs) go FASTQ -> reader -> (emulated device logic) -> BAM writer -> BAM parser, and every field
the reference's `check_results` compares must match `shared_expectation()` (:464-868)."""
import ctypes as C
import gzip
import json
import os

import numpy as np

from bamio import read_bam, write_bam
from emu import emu
from helpers import product_params, revcomp
from mapad_b200 import abi, api
from ref_cases import INTEGRATION_PARAMS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fastq_to_bam_integration_fixture(tmp_path):
    data = json.load(open(os.path.join(GOLDEN, "ref_integration.json")))
    index = api.Index.build([(n, s) for n, s in data["contigs"]], draws="A")
    fq = tmp_path / "reads.fastq.gz"
    with gzip.open(fq, "wt") as f:
        for r in data["reads"]:
            s, q = r["seq"], r["qual"]
            if r["flag"] & 16:  # the BAM input of the reference stores flag-16 reads reverse-complemented (record.rs:159-162)
                s, q = revcomp(s), q[::-1]
            f.write("@%s some description\n%s\n+\n%s\n" % (r["name"], s.lower() if "A00123" in r["name"] else s, q))
        f.write("@broken\nACGT\n+\n]]\n")  # length mismatch: skipped like input_chunk_reader.rs:206-214
    chunks = api.FastqChunks(str(fq), batch_size=10)
    out = tmp_path / "out.bam"
    w = api.BamWriter(str(out), index, command_line="mapad map test", read_group_id=None)
    params = product_params(INTEGRATION_PARAMS)
    total = 0
    for R, names, noff, flags, n, ch in chunks:
        seq = np.ctypeslib.as_array(C.cast(R.seq, C.POINTER(C.c_uint8)), shape=(int(C.cast(R.offsets, C.POINTER(C.c_uint64))[n]),)).copy()
        qual = np.ctypeslib.as_array(C.cast(R.qual, C.POINTER(C.c_uint8)), shape=(len(seq),)).copy()
        off = np.ctypeslib.as_array(C.cast(R.offsets, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        res = emu.map_batch(index, params, seeds=np.arange(n, dtype=np.uint32) + 100, packed=(seq, qual, off))
        rs, keep = abi.results_struct(res)
        w.write_chunk(R, names, noff, flags, rs)
        total += n
        chunks.free(ch)
    w.close()
    chunks.close()
    assert total == 17 and chunks.skipped == 1
    text, refs, recs = read_bam(str(out))
    assert text.startswith("@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:600\n@SQ\tSN:Chromosome_02\tLN:600\n@SQ\tSN:Chromosome_03\tLN:84\n@SQ\tSN:Chromosome_04\tLN:46\n")
    assert "@PG\tID:mapAD\tPN:mapAD" in text
    assert refs == [("chr1", 600), ("Chromosome_02", 600), ("Chromosome_03", 84), ("Chromosome_04", 46)]
    exp = {e["name"]: e for e in data["expectation"]}
    assert [r["name"] for r in recs] == [r["name"] for r in data["reads"]]  # input order
    for rec in recs:
        e = exp[rec["name"]]
        if e["tid"] is None:
            assert rec["flag"] & 4 and rec["ref_id"] == -1 and rec["mapq"] == 0 and rec["cigar"] == "" and "MD" not in rec["tags"]
            assert rec["seq"] == e["seq"]
            continue
        # FASTQ input carries no flags: only the strand bit can be set (the BAM-input flag cases are covered in test_bam_flags)
        assert rec["flag"] == (16 if e["flags"] & 16 else 0), rec["name"]
        assert (rec["ref_id"], rec["pos"] + 1, rec["mapq"], rec["cigar"], rec["seq"]) == (e["tid"], e["pos"], e["mq"], e["cigar"], e["seq"]), rec["name"]
        assert rec["qual"] == bytes(c - 33 for c in e_qual(data, rec["name"], bool(e["flags"] & 16)))
        t = rec["tags"]
        assert (t["MD"], t["X0"], t["X1"], t["XT"]) == (e["md"], e["x0"], e["x1"], e["xt"]), rec["name"]
        assert t.get("XA") == e["xa"], rec["name"]
        if e["xs"] is not None:
            assert np.float32(t["XS"]) == np.float32(e["xs"])
        else:
            assert "XS" not in t
        assert "AS" in t and "NM" in t


def e_qual(data, name, reverse_out):
    r = [x for x in data["reads"] if x["name"] == name][0]
    q = r["qual"].encode()
    if r["flag"] & 16:
        q = q[::-1]        # as read from the FASTQ
    return q[::-1] if reverse_out else q


def test_bam_flags(tmp_path):
    """Flag clean-up of create_bam_record (mapping.rs:748-776) with explicit input flags (e.g. 589 -> 577)."""
    data = json.load(open(os.path.join(GOLDEN, "ref_integration.json")))
    index = api.Index.build([(n, s) for n, s in data["contigs"]], draws="A")
    reads = [r for r in data["reads"] if r["name"].startswith(("A00234", "A00789_0130", "A00456"))]
    seqs, quals, in_flags = [], [], []
    for r in reads:
        s, q = r["seq"], bytes(c - 33 for c in r["qual"].encode())
        if r["flag"] & 16:
            s, q = revcomp(s), q[::-1]
        seqs.append(s.encode()); quals.append(q); in_flags.append(r["flag"])
    packed = abi.pack_reads(seqs, quals)
    res = emu.map_batch(index, product_params(INTEGRATION_PARAMS), seeds=[1, 2, 3], packed=packed)
    R, keep = api.make_reads(*packed)
    names = b"".join(r["name"].encode() for r in reads)
    noff = np.cumsum([0] + [len(r["name"]) for r in reads]).astype(np.uint64)
    fl = np.array(in_flags, dtype=np.uint16)
    out = tmp_path / "flags.bam"
    w = api.BamWriter(str(out), index, read_group_id="RG01", force_overwrite=True)
    rs, keep2 = abi.results_struct(res)
    w.write_chunk(R, C.c_char_p(names), noff.ctypes.data, fl.ctypes.data, rs)
    w.close()
    text, refs, recs = read_bam(str(out))
    assert "@RG\tID:RG01\n" in text
    exp = {e["name"]: e for e in data["expectation"]}
    for rec in recs:
        assert rec["flag"] == exp[rec["name"]]["flags"], rec["name"]
        assert rec["tags"]["RG"] == "RG01"


def test_bam_input_to_bam_output(tmp_path):
    """BAM in -> BAM out (record.rs:138-182, create_bam_header, create_bam_record): the integration reads as the
    reference's own BAM fixture stores them (flag-16 reads reverse-complemented, flags such as 589), with auxiliary
    fields and a header to carry over."""
    import struct
    data = json.load(open(os.path.join(GOLDEN, "ref_integration.json")))
    index = api.Index.build([(n, s) for n, s in data["contigs"]], draws="A")
    aux = (b"XYZkeep me\x00" + b"NMi" + struct.pack("<i", 99) + b"RGZold\x00" + b"abBs" + struct.pack("<Ihhh", 3, 1, -2, 3) +
           b"xff" + struct.pack("<f", 1.5) + b"ASC\x07" + b"xhH1AE301\x00" + b"XAZstale\x00" + b"zcc\xfe")
    recs_in = [dict(name=r["name"], flag=r["flag"], seq=r["seq"], qual=bytes(c - 33 for c in r["qual"].encode()), aux=aux if i % 2 == 0 else b"")
               for i, r in enumerate(data["reads"])]
    recs_in.insert(3, dict(name="noqual", flag=4, seq="ACGTACGTAC", qual=None))  # missing qualities: skipped
    header = ("@HD\tVN:1.0\tSO:queryname\n@SQ\tSN:stale\tLN:5\n@RG\tID:old\tSM:x\n@PG\tID:bwa\tPN:bwa\n"
              "@PG\tID:mapAD\tPN:mapAD\tPP:bwa\n@CO\tfirst comment\n")
    src = tmp_path / "in.bam"
    write_bam(str(src), header, [("stale", 5)], recs_in, block=700)  # records straddle BGZF blocks
    params = product_params(INTEGRATION_PARAMS)
    exp = {e["name"]: e for e in data["expectation"]}

    def run(read_group):
        chunks = api.ReadChunks(str(src), batch_size=7)
        assert chunks.is_bam and chunks.header_text == header.rstrip("\n") + "\n"
        out = tmp_path / ("out_%s.bam" % (read_group or "none"))
        w = api.BamWriter(str(out), index, command_line="mapad map bam", read_group_id=read_group, src_header_text=chunks.header_text)
        total = 0
        for R, names, noff, flags, n, ch in chunks:
            tb = int(C.cast(R.offsets, C.POINTER(C.c_uint64))[n])
            seq = np.ctypeslib.as_array(C.cast(R.seq, C.POINTER(C.c_uint8)), shape=(tb,)).copy()
            qual = np.ctypeslib.as_array(C.cast(R.qual, C.POINTER(C.c_uint8)), shape=(tb,)).copy()
            off = np.ctypeslib.as_array(C.cast(R.offsets, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
            res = emu.map_batch(index, params, seeds=np.arange(n, dtype=np.uint32) + 100, packed=(seq, qual, off))
            rs, keep = abi.results_struct(res)
            w.write_chunk(R, names, noff, flags, rs, chunk=ch)
            total += n
            chunks.free(ch)
        w.close()
        chunks.close()
        assert total == 17 and chunks.skipped == 1
        return read_bam(str(out))

    text, refs, recs = run(None)
    lines = text.rstrip("\n").split("\n")
    assert lines[0] == "@HD\tVN:1.6\tSO:unsorted" and lines[1:5] == ["@SQ\tSN:chr1\tLN:600", "@SQ\tSN:Chromosome_02\tLN:600",
                                                                    "@SQ\tSN:Chromosome_03\tLN:84", "@SQ\tSN:Chromosome_04\tLN:46"]
    assert "stale" not in text
    assert lines[5] == "@RG\tID:old\tSM:x" and lines[6] == "@PG\tID:bwa\tPN:bwa" and lines[7] == "@PG\tID:mapAD\tPN:mapAD\tPP:bwa"
    assert lines[8].startswith("@PG\tID:mapAD.1\tPN:mapAD\t") and lines[8].endswith("\tCL:mapad map bam\tPP:mapAD")
    assert lines[9] == "@CO\tfirst comment" and len(lines) == 10
    assert [r["name"] for r in recs] == [r["name"] for r in data["reads"]]
    for i, rec in enumerate(recs):
        e = exp[rec["name"]]
        assert rec["flag"] == e["flags"], rec["name"]  # e.g. 589 -> 577 (mapping.rs:748-776)
        t = rec["tags"]
        if i % 2 == 0:  # carried over, in input order, ahead of the new tags; filtered: NM AS XA of the input
            assert rec["tag_order"][:6] == ["XY", "RG", "ab", "xf", "xh", "zc"], rec["tag_order"]
            assert (t["XY"], t["RG"], t["ab"], t["xf"], t["xh"], t["zc"]) == ("keep me", "old", ("s", [1, -2, 3]), 1.5, ("H", "1AE301"), -2)
        else:
            assert "XY" not in t and "RG" not in t
        if e["tid"] is None:
            assert rec["flag"] & 4 and rec["ref_id"] == -1 and "NM" not in t and "AS" not in t and "XA" not in t
            continue
        assert (rec["ref_id"], rec["pos"] + 1, rec["mapq"], rec["cigar"], rec["seq"]) == (e["tid"], e["pos"], e["mq"], e["cigar"], e["seq"]), rec["name"]
        assert (t["MD"], t["X0"], t["X1"], t["XT"], t.get("XA")) == (e["md"], e["x0"], e["x1"], e["xt"], e["xa"]), rec["name"]
        assert t["NM"] != 99 and isinstance(t["AS"], float)

    text, refs, recs = run("NEW")
    assert "@RG\tID:NEW\n" in text and "ID:old" not in text
    for i, rec in enumerate(recs):
        assert rec["tags"]["RG"] == "NEW" and rec["tag_order"].count("RG") == 1
        if i % 2 == 0:
            assert rec["tag_order"][:6] == ["XY", "ab", "xf", "xh", "zc", "RG"]


def test_input_sniffing(tmp_path):
    plain = tmp_path / "r.fq"
    plain.write_text("@r1 desc\nacgtn\n+\nIIII!\n@r2\nGG\n+\n#$\n")
    ch = api.ReadChunks(str(plain), batch_size=10)
    assert not ch.is_bam and ch.header_text is None
    R, names, noff, flags, n, h = next(ch)
    assert n == 2 and bytes(C.cast(R.seq, C.POINTER(C.c_uint8))[0:7]) == b"ACGTNGG"
    assert list(C.cast(R.qual, C.POINTER(C.c_uint8))[0:7]) == [40, 40, 40, 40, 0, 2, 3]
    ch.free(h)
    ch.close()
    cram = tmp_path / "x.cram"
    cram.write_bytes(b"CRAM\x03\x00" + b"\x00" * 30)
    try:
        api.ReadChunks(str(cram))
        assert False, "CRAM must be refused"
    except api.MapadError as e:
        assert e.code == -1


def test_bam_writer_threads_are_deterministic(tmp_path, monkeypatch):
    """BGZF blocks are compressed by several host threads; the file must not depend on how many."""
    data = json.load(open(os.path.join(GOLDEN, "ref_integration.json")))
    index = api.Index.build([(n, s) for n, s in data["contigs"]], draws="A")
    seqs = [r["seq"].encode() for r in data["reads"]] * 400  # ~6 800 records: several BGZF blocks
    quals = [bytes(c - 33 for c in r["qual"].encode()) for r in data["reads"]] * 400
    packed = abi.pack_reads(seqs, quals)
    res = emu.map_batch(index, product_params(INTEGRATION_PARAMS), seeds=np.arange(len(seqs), dtype=np.uint32), packed=packed)
    rs, keep = abi.results_struct(res)
    R, keep2 = api.make_reads(*packed)
    names = b"".join(b"r%06d" % i for i in range(len(seqs)))
    noff = (np.arange(len(seqs) + 1, dtype=np.uint64) * 7)
    fl = np.zeros(len(seqs), dtype=np.uint16)
    blobs = []
    for nt in ("1", "5"):
        monkeypatch.setenv("MAPAD_BAM_THREADS", nt)
        out = tmp_path / ("t%s.bam" % nt)
        w = api.BamWriter(str(out), index, force_overwrite=True)
        for _ in range(3):
            w.write_chunk(R, C.c_char_p(names), noff.ctypes.data, fl.ctypes.data, rs)
        w.close()
        blobs.append(open(out, "rb").read())
    assert blobs[0] == blobs[1] and len(blobs[0]) > 100_000
    text, refs, recs = read_bam(str(tmp_path / "t5.bam"))
    assert len(recs) == 3 * len(seqs) and recs[-1]["name"] == "r%06d" % (len(seqs) - 1)


def test_cli_flags_follow_reference():
    """`mapad map` flag surface (src/main.rs:30-303): -p xor -c/-e, probability validators, -l, -v, global --seed/--threads."""
    import pytest
    from mapad_b200 import abi, cli
    ap = cli.build_parser()
    base = ["map", "-r", "r.fq", "-g", "g.fa", "-o", "o.bam", "-l", "single_stranded", "-f", "0.5", "-t", "0.5", "-d", "0.02", "-s", "1.0",
            "-i", "0.001", "-x", "0.5"]
    a = ap.parse_args(base + ["-p", "0.03", "-v", "-v", "--threads", "8", "--seed", "7"])
    assert a.v == 2 and a.seed == 7 and a.library == "single_stranded"
    P = cli.params_from_args(a)
    assert P.bound_kind == abi.BOUND_DISCRETE and abs(P.poisson_threshold - 0.03) < 1e-7
    a = ap.parse_args(base + ["-c", "0.8", "-e", "0.9"])
    P = cli.params_from_args(a)
    assert P.bound_kind == abi.BOUND_CONTINUOUS and abs(P.cutoff + 0.8) < 1e-6 and abs(P.exponent - 0.9) < 1e-6
    with pytest.raises(SystemExit):   # neither -p nor -c
        cli.params_from_args(ap.parse_args(base))
    with pytest.raises(SystemExit):   # both
        cli.params_from_args(ap.parse_args(base + ["-p", "0.03", "-c", "0.8"]))
    with pytest.raises(SystemExit):   # not a probability
        ap.parse_args(base + ["-p", "1.5"])
    with pytest.raises(SystemExit):
        ap.parse_args([x if x != "0.02" else "-0.1" for x in base] + ["-p", "0.03"])
    ix = ap.parse_args(["index", "-g", "g.fa", "--seed", "5"])
    assert ix.seed == 5
