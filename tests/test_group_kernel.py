"""Non-GPU check of the group search kernel (mapad_b200/csrc/search_group.cuh): its source is compiled as plain C++ and
run under the SIMT emulator of tests/emu (lanes of a group = coroutines, group collectives = rendezvous), for several
group sizes, both index layouts, tiny chunk pools (deferral + retry) and small search limits (pop_min eviction), and
compared bit for bit with the oracle.  The real CUDA run is covered by tests/test_gpu_parity.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from compare import compare_results
from helpers import oracle_params, product_params, ora, random_genome, simulate_reads
from mapad_b200 import api
from ref_cases import SEARCH_CASES, cli_params
from emu import emu
from test_emulated_kernels import oracle_index_from_product


@pytest.fixture(scope="module")
def small_world():
    genome = random_genome(60000, seed=42)
    index = api.Index.build([("chr1", genome[:25000]), ("chr2", genome[25000:])])
    return genome, index, oracle_index_from_product(index)


@pytest.mark.parametrize("G", [1, 4, 8, 32])
@pytest.mark.parametrize("layout", [-1, 1], ids=["narrow", "wide"])
def test_group_sizes_vs_oracle(small_world, G, layout):
    genome, index, oix = small_world
    spec = cli_params("single_stranded")
    seqs, quals = simulate_reads(genome, 150, (25, 70), seed=1001)
    seqs[5] = seqs[5][:10] + b"N" + seqs[5][11:]
    seqs[6] = b""
    quals[6] = b""
    seeds = np.arange(len(seqs), dtype=np.uint32) * 7919
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds, layout=layout, group=dict(G=G, n_groups=3))
    compare_results(want, got)


def test_reference_cases_group():
    for case in SEARCH_CASES:
        index = api.Index.build([("ref", case["ref"])])
        oix = oracle_index_from_product(index)
        pat = case["pattern"].encode()
        q = bytes([case["qual"]] * len(pat))
        want = ora.map_batch(oix, oracle_params(case), [pat], [q], seeds=[7], want_hits=True)
        for G in (1, 8):
            got = emu.map_batch(index, product_params(case), [pat], [q], seeds=[7], group=dict(G=G, n_groups=1))
            compare_results(want, got)


def test_double_stranded_and_bidirectional(small_world):
    genome, index, oix = small_world
    seqs, quals = simulate_reads(genome, 120, (25, 70), seed=5, library="double_stranded")
    seeds = np.arange(len(seqs), dtype=np.uint32)
    spec = cli_params("double_stranded")
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds, group=dict(G=8, n_groups=2))
    compare_results(want, got)
    # TestDifferenceModel starts in the middle of the read (find_alignment_start = len / 2): forward and backward steps
    spec2 = dict(model=("test", -1.0, -2.0, 0.0), bound=("test", -5.0, None), gaps=(-4.0, -1.0, 3, 2))
    want = ora.map_batch(oix, oracle_params(spec2), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    for G in (1, 8):
        got = emu.map_batch(index, product_params(spec2), seqs, quals, seeds=seeds, group=dict(G=G, n_groups=2))
        compare_results(want, got)
    # Continuous bound (mismatch_bounds.rs:77-121)
    spec3 = dict(cli_params("single_stranded"))
    spec3["bound"] = ("continuous", -0.25, 1.0)
    want = ora.map_batch(oix, oracle_params(spec3), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    got = emu.map_batch(index, product_params(spec3), seqs, quals, seeds=seeds, group=dict(G=8, n_groups=2))
    compare_results(want, got)


@pytest.mark.parametrize("abort", [False, True])
def test_limits_eviction_group(small_world, abort):
    """STACK_LIMIT / EDIT_TREE_LIMIT recovery with small limits: pop_min eviction and slab key reuse (mapping.rs:1358-1380)."""
    genome, index, oix = small_world
    spec = dict(cli_params("single_stranded"))
    spec["limits"] = (300, 700)
    spec["abort"] = abort
    seqs, quals = simulate_reads(genome, 150, (30, 90), seed=77)
    seeds = np.arange(len(seqs), dtype=np.uint32)
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    assert sum(1 for r in want.records if r["flags"] & 1) > 5
    for G in (1, 8):
        got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds, group=dict(G=G, n_groups=3))
        compare_results(want, got)


def test_lane_order_independence(small_world):
    """The emulator can run the lanes of a group in reverse order (lane G-1 first): results must not depend on which lane
    gets ahead, i.e. every read of shared state is fenced from the writes around it."""
    genome, index, oix = small_world
    spec = cli_params("single_stranded")
    seqs, quals = simulate_reads(genome, 80, (25, 70), seed=9)
    seeds = np.arange(len(seqs), dtype=np.uint32)
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    os.environ["MAPAD_SIMT_EMU_ORDER"] = "reverse"
    try:
        got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds, group=dict(G=8, n_groups=3))
    finally:
        del os.environ["MAPAD_SIMT_EMU_ORDER"]
    compare_results(want, got)


def test_small_chunks_variant():
    """4 KiB pool chunks (128 nodes / 64 heap lines per chunk) so that every read crosses many chunk boundaries, plus a pool
    that runs dry (reads are handed back and re-run with fewer groups in flight): separate build of the emulation."""
    if os.environ.get("MAPAD_EMU_DEFS"):
        pytest.skip("already inside a variant run")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MAPAD_EMU_DEFS="-DMAPAD_GCHUNK_SHIFT=12u")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_group_kernel.py"), "-x", "-q", "-k",
                        "group_sizes or limits or dry"], env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_pool_runs_dry(small_world):
    if "-DMAPAD_GCHUNK_SHIFT=12u" not in os.environ.get("MAPAD_EMU_DEFS", ""):
        pytest.skip("needs the small-chunk build (run through test_small_chunks_variant)")
    genome, index, oix = small_world
    spec = cli_params("single_stranded")
    seqs, quals = simulate_reads(genome, 120, (40, 90), seed=3)
    seeds = np.arange(len(seqs), dtype=np.uint32)
    want = ora.map_batch(oix, oracle_params(spec), seqs, quals, seeds=seeds, n_threads=4, want_hits=True)
    # the largest read alone needs ~0.0235 chunks of 4 KiB per popped frame (2.5 nodes + 1.5 heap entries); a pool that
    # just covers it runs dry while the other seven groups hold chunks too
    need = int(int(max(r["frames_popped"] for r in want.records)) * 0.0257)
    got = emu.map_batch(index, product_params(spec), seqs, quals, seeds=seeds, group=dict(G=8, n_groups=8, pool_chunks=16 + need))
    assert emu.map_batch.last_deferred > 0
    compare_results(want, got)
    assert sum(1 for r in got.records if r["flags"] & 2) > 0
