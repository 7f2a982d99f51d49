"""Non-GPU check of the PRODUCT's per-read device logic: the code of mapad_b200/csrc/*_core.cuh is
compiled as plain C++ (tests/emu) and compared bit for bit with the oracle on the reference's
known-answer cases and on simulated reads.  The real CUDA run is covered by tests/test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

from compare import compare_results
from helpers import oracle_params, product_params, ora, revcomp, random_genome, simulate_reads
from mapad_b200 import api
from ref_cases import BENCH_PARAMS, INTEGRATION_PARAMS, SEARCH_CASES, cli_params
from emu import emu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_index_from_product(index, occ_k=128):
    a = index.arrays()
    return ora.OracleIndex.from_arrays(a["bwt"], a["sa_sample"], a["sa_rate"], a["extra_rows"], a["contigs"], a["orig_pos"], a["orig_sym"],
                                       with_x=True, occ_k=occ_k)


@pytest.mark.parametrize("layout", [-1, 1], ids=["narrow", "wide"])
@pytest.mark.parametrize("case", SEARCH_CASES, ids=[c["name"] for c in SEARCH_CASES])
def test_reference_cases(case, layout):
    index = api.Index.build([("ref", case["ref"])])
    oix = oracle_index_from_product(index)
    pat = case["pattern"].encode()
    q = bytes([case["qual"]] * len(pat))
    want = ora.map_batch(oix, oracle_params(case), [pat], [q], seeds=[7], want_hits=True)
    got = emu.map_batch(index, product_params(case), [pat], [q], seeds=[7], layout=layout)
    compare_results(want, got)


def test_index_cross_check():
    """Product index (SA-IS) vs oracle index (prefix doubling) on a genome with N runs, and rank queries
    through the device block layout vs the oracle's Occ."""
    rng = np.random.default_rng(3)
    g = random_genome(3000, seed=11)
    g = g[:500] + "N" * 25 + g[500:900] + "NNN" + g[900:1500] + "RYK" + g[1500:] + "N" * 40
    contigs = [("c1", g[:1200]), ("c2", g[1200:])]
    draws = "ACGTACGTAC"
    index = api.Index.build(contigs, draws=draws)
    oix = ora.OracleIndex.build(contigs, draws=draws)
    a = index.arrays()
    assert a["n"] == oix.n
    assert np.array_equal(a["bwt"], oix.bwt())
    assert a["less"][: len(oix.less())] == oix.less()
    assert a["sentinel_rows"] == oix.sentinel_rows()
    assert np.array_equal(a["sa_sample"], oix.sa_samples())
    assert [tuple(x) for x in a["extra_rows"].tolist()] == oix.extra_rows()
    assert dict(zip(a["orig_pos"].tolist(), a["orig_sym"].tolist())) == oix.original_symbols()
    for layout in (-1, 1):
        for row in list(range(0, 200)) + [int(x) for x in rng.integers(0, a["n"], size=400)] + [a["n"] - 1]:
            c4, b = emu.occ4(index, row, layout)
            assert c4 == [oix.occ(row, k) for k in (1, 2, 3, 4)], (layout, row)
            assert b == int(a["bwt"][row])
            assert emu.sa_get(index, row, layout) == oix.sa_get(row), (layout, row)


def test_index_cross_check_5mbp():
    """The two independent indexers (product: SA-IS; oracle: its own suffix sorter) on a 5 Mbp text with 1 kbp tandem repeats,
    a palindromic stretch and N runs on both sides of the 20 bp threshold — a size and a repeat structure at which different
    suffix-sorting strategies would diverge if either were wrong."""
    g = random_genome(5_000_000, seed=21)
    unit = g[1000:2000]
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    pal = g[7000:7400]
    pal = pal + "".join(comp[c] for c in reversed(pal))
    g = g[:300_000] + unit * 6 + g[300_000:1_200_000] + "N" * 37 + g[1_200_000:2_500_000] + pal + g[2_500_000:4_000_000] + "N" * 7 + unit * 3 + g[4_000_000:]
    contigs = [("c1", g[:2_000_000]), ("c2", g[2_000_000:])]
    draws = "ACGTTGCA"
    index = api.Index.build(contigs, draws=draws)
    oix = ora.OracleIndex.build(contigs, draws=draws)
    a = index.arrays()
    assert a["n"] == oix.n == 2 * len(g) + 2
    assert np.array_equal(a["bwt"], oix.bwt())
    assert a["less"][: len(oix.less())] == oix.less()
    assert a["sentinel_rows"] == oix.sentinel_rows()
    assert np.array_equal(a["sa_sample"], oix.sa_samples())
    assert [tuple(x) for x in a["extra_rows"].tolist()] == oix.extra_rows()
    assert dict(zip(a["orig_pos"].tolist(), a["orig_sym"].tolist())) == oix.original_symbols()


def test_bench_reads_and_integration():
    data = json.load(open(os.path.join(GOLDEN, "ref_test_bench.json")))
    index = api.Index.build([("ref", data["ref_seq"])])
    oix = oracle_index_from_product(index)
    seqs = [r["pattern"].encode() for r in data["reads"]]
    quals = [bytes([40] * len(s)) for s in seqs]
    want = ora.map_batch(oix, oracle_params(BENCH_PARAMS), seqs, quals, want_hits=True)
    got = emu.map_batch(index, product_params(BENCH_PARAMS), seqs, quals)
    compare_results(want, got)
    assert [int(r["n_hits"]) for r in got.records] == [r["n_hits"] for r in data["reads"]]
    # integration fixture: product index with the 'N' drawn as 'A'
    data = json.load(open(os.path.join(GOLDEN, "ref_integration.json")))
    index = api.Index.build([(n, s) for n, s in data["contigs"]], draws="A")
    oix = ora.OracleIndex.build([(n, s) for n, s in data["contigs"]], draws="A")
    seqs, quals = [], []
    for r in data["reads"]:
        s, q = r["seq"], bytes(c - 33 for c in r["qual"].encode())
        if r["flag"] & 16:
            s, q = revcomp(s), q[::-1]
        seqs.append(s.encode())
        quals.append(q)
    seeds = list(range(100, 100 + len(seqs)))
    want = ora.map_batch(oix, oracle_params(INTEGRATION_PARAMS), seqs, quals, seeds=seeds, want_hits=True)
    for layout in (-1, 1):
        got = emu.map_batch(index, product_params(INTEGRATION_PARAMS), seqs, quals, seeds=seeds, layout=layout)
        compare_results(want, got)
    exp = {e["name"]: e for e in data["expectation"]}
    for i, r in enumerate(data["reads"]):
        e = exp[r["name"]]
        s = got.record_summary(i)
        if e["tid"] is None:
            assert not s["mapped"]
        else:
            assert (s["tid"], s["pos"] + 1, s["mapq"], s["cigar"], s["md"], s["x0"], s["x1"], s["xt"]) == \
                   (e["tid"], e["pos"], e["mq"], e["cigar"], e["md"], e["x0"], e["x1"], e["xt"]), r["name"]
            assert got.xa[i] == (e["xa"] or "")


@pytest.mark.parametrize("library", ["single_stranded", "double_stranded"])
def test_simulated_reads(library):
    genome = random_genome(60000, seed=42)
    index = api.Index.build([("chr1", genome[:25000]), ("chr2", genome[25000:])])
    oix = oracle_index_from_product(index)
    spec = cli_params(library)
    seqs, quals = simulate_reads(genome, 300, (25, 70), seed=1001, library=library)
    seqs[5] = seqs[5][:10] + b"N" + seqs[5][11:]
    seeds = np.arange(len(seqs), dtype=np.uint32) * 7919
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds)
    compare_results(want, got)
    assert sum(int(r["mapped"]) for r in got.records) > 150


def test_continuous_bound_and_custom_model():
    """Continuous mismatch bound (-c/-e, mismatch_bounds.rs:77-121) and a user-supplied SequenceDifferenceModel
    (callback across the C ABI, host-evaluated penalty table) against the oracle."""
    import ctypes as C
    from mapad_b200 import abi
    genome = random_genome(30000, seed=3)
    index = api.Index.build([("c", genome)])
    oix = oracle_index_from_product(index)
    seqs, quals = simulate_reads(genome, 120, (25, 60), seed=17)
    seeds = np.arange(len(seqs), dtype=np.uint32)
    # 1. continuous bound with the aDNA model
    spec = dict(cli_params("single_stranded"))
    spec["bound"] = ("continuous", -0.25, 1.0)
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds)
    compare_results(want, got)
    assert 0 < sum(int(r["mapped"]) for r in got.records) < len(seqs) + 1
    # 2. custom model = TestDifferenceModel semantics through the callback, bidirectional start (len / 2)
    spec2 = dict(model=("test", -1.0, -2.0, 0.0), bound=("test", -5.0, None), gaps=(-4.0, -1.0, 3, 2))
    P = product_params(spec2)
    P.model_kind = abi.MODEL_CUSTOM

    @abi.SDM_GET_FN
    def get(user, i, L, frm, to, q):
        if frm == ord("C") and to == ord("T"):
            return -1.0
        return 0.0 if frm == to else -2.0

    P.custom_get = get
    want = ora.map_batch(oix, oracle_params(spec2), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, P, seqs, quals, seeds=seeds)
    compare_results(want, got)


def test_device_logic_variants():
    """Build-time variants of the device logic that are candidates for the next tuning round must stay bit-exact too:
    the whole emulation suite is re-run against each variant build (MAPAD_EMU_DEFS)."""
    import subprocess
    import sys
    if os.environ.get("MAPAD_EMU_DEFS"):
        pytest.skip("already inside a variant run")
    here = os.path.dirname(os.path.abspath(__file__))
    for defs in ("-DMAPAD_COMPACT_CAND=1",):
        env = dict(os.environ, MAPAD_EMU_DEFS=defs)
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_emulated_kernels.py"), "-x", "-q", "-k", "not variants"],
                           env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, defs + "\n" + r.stdout[-3000:] + r.stderr[-2000:]
