"""FASTQ(.GZ) reader and BAM writer on the CPU: the 17 reads of the reference's integration test
(tests/integration_tests.rs) go FASTQ -> reader -> (emulated device logic) -> BAM writer -> BAM parser, and every field
the reference's `check_results` compares must match `shared_expectation()` (:464-868)."""
import ctypes as C
import gzip
import json
import os

import numpy as np

from bamio import read_bam
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
