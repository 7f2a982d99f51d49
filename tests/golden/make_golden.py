#!/usr/bin/env python3
"""Extracts the bulky known-answer data of the reference's own tests into tests/golden/*.json.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources:
  src/map/mapping.rs:2669-2956                 test_bench: 10 kbp reference, seven 100 bp reads, hit counts
  src/map/sequence_difference_models.rs:451-1276  SimpleAncientDnaModel value tables (2 x 400 values)
  tests/integration_tests.rs:58-172,464-868    FASTA, 17 reads and the per-record expectation
The small search known-answer cases are transcribed by hand in tests/test_oracle_golden.py.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def extract_bench():
    src = open(os.path.join(REF, "src/map/mapping.rs")).read()
    start = src.index("fn test_bench()")
    body = src[start:]
    m = re.search(r'let ref_seq = "(.*?)"\s*\.as_bytes\(\)', body, re.S)
    ref_seq = re.sub(r"[\\\s]", "", m.group(1))
    reads = []
    for mm in re.finditer(r'// (bench_\w+)\s*\{(.*?)assert_eq!\(intervals\.len\(\), (\d+)\);', body, re.S):
        name, blk, n = mm.group(1), mm.group(2), int(mm.group(3))
        pat = re.search(r'let pattern = "([ACGT]+)"', blk).group(1)
        reads.append(dict(name=name, pattern=pat, n_hits=n))
    assert len(ref_seq) == 10000 and len(reads) == 7, (len(ref_seq), len(reads))
    return dict(ref_seq=ref_seq, reads=reads)


def extract_sdm():
    src = open(os.path.join(REF, "src/map/sequence_difference_models.rs")).read()
    out = {}
    for fn in ("test_simple_adna_model", "test_simple_adna_model_ds"):
        start = src.index("fn %s()" % fn)
        end = src.index("#[test]", start)
        body = src[start:end]
        vals = []
        for mm in re.finditer(
            r"assert_approx_eq!\((-?[0-9._]+), adna_model\.get\((\d+), (\d+), b'(.)', b'(.)', (\d+)\)\);", body
        ):
            vals.append([float(mm.group(1).replace("_", "")), int(mm.group(2)), int(mm.group(3)), mm.group(4), mm.group(5), int(mm.group(6))])
        out[fn] = vals
        assert len(vals) == 400, (fn, len(vals))
    return out


def extract_integration():
    src = open(os.path.join(REF, "tests/integration_tests.rs")).read()
    fasta = re.search(r'let fasta_content = "(.*?)";', src, re.S).group(1)
    contigs = []
    for block in fasta.split(">")[1:]:
        lines = block.strip().split("\n")
        contigs.append([lines[0].strip(), "".join(l.strip() for l in lines[1:])])
    sam = re.search(r'let sam_content = b"\\\n(.*?)";', src, re.S).group(1)
    reads = []
    for line in sam.split("\n"):
        line = line.strip()
        if line.endswith("\\n\\"):
            line = line[:-3]
        line = line.replace("\\t", "\t").replace("\\\\", "\\")
        if not line or line.startswith("@"):
            continue
        f = line.split("\t")
        reads.append(dict(name=f[0], flag=int(f[1]), seq=f[9], qual=f[10]))
    assert len(reads) == 17, len(reads)
    exp = []
    body = src[src.index("fn shared_expectation()"):]
    for blk in body.split("BamFieldSubset {")[1:]:
        def g(pat, cast=str, default=None):
            m = re.search(pat, blk, re.S)
            return cast(m.group(1)) if m else default
        name = g(r'name: Some\(b"(.*?)"')
        flags = g(r"flags: (\d+)\.into", int)
        tid = g(r"tid: Some\((\d+)", int)
        pos = g(r"pos: Some\((\d+)", int)
        mq = g(r"mq: Some\((\d+)", int)
        cig = re.findall(r"cigar::Op::new\(cigar::op::Kind::(\w+), (\d+)\)", blk)
        cigar = "".join("%s%s" % (n, {"Match": "M", "Insertion": "I", "Deletion": "D"}[k]) for k, n in cig)
        seq = g(r'seq: b"(.*?)"')
        md = g(r'md: Some\("(.*?)"')
        x0 = g(r"x0: Some\((\d+)", int)
        x1 = g(r"x1: Some\((\d+)", int)
        xa = g(r'xa: Some\(\s*"(.*?)"', str)
        xs = g(r"xs: Some\((-?[0-9.]+)", float)
        xt = g(r"xt: Some\('(.)'", str)
        exp.append(dict(name=name, flags=flags, tid=tid, pos=pos, mq=mq, cigar=cigar, seq=seq, md=md, x0=x0, x1=x1, xa=xa, xs=xs, xt=xt))
    assert len(exp) == 17, len(exp)
    return dict(contigs=contigs, reads=reads, expectation=exp)


if __name__ == "__main__":
    json.dump(extract_bench(), open(os.path.join(OUT, "ref_test_bench.json"), "w"))
    json.dump(extract_sdm(), open(os.path.join(OUT, "ref_sdm_values.json"), "w"))
    json.dump(extract_integration(), open(os.path.join(OUT, "ref_integration.json"), "w"), indent=1)
    print("golden vectors written to", OUT)
