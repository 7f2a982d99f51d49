"""Known-answer cases of the reference's own unit tests for the hot path, transcribed as data.

Every case cites the reference test it restates (file:line under /root/reference).  The same table
drives the oracle tests (tests/test_oracle_golden.py, CPU) and the CUDA parity tests
(tests/test_gpu_parity.py, `-m gpu`), so both are held to the reference's expectations.
"""

# model: ("test", deam, mm, match) | ("vindija",) | ("simple", library, f, t, d, s, divergence, ignore_q)
# bound: ("test", threshold, repr_mm_bound|None=model repr mm) | ("discrete", p, rate)
# gaps:  (open, extend, gap_dist_ends, max_num_gaps_open); "repr" multiples are given as ("repr", factor)

TEST_MODEL_10 = ("test", -10.0, -10.0, 0.0)

SEARCH_CASES = [
    dict(  # src/map/mapping.rs:1401-1455
        name="test_inexact_search", model=("test", -0.5, -1.0, 0.0), bound=("test", -1.0, -1.0), gaps=(-2.0, -1.0, 0, 2),
        ref="ACGTACGTACGTACGT", pattern="GTTC", qual=0,
        expect=dict(scores_iter=[-1.0], positions_sorted=[2, 6, 10, 19, 23, 27]),
    ),
    dict(  # mapping.rs:1458-1509
        name="test_reverse_strand_search", model=TEST_MODEL_10, bound=("test", -1.0, -10.0), gaps=(-20.0, -10.0, 0, 2),
        ref="GAAAAG", pattern="TTTT", qual=0, expect=dict(positions_sorted=[8]),
    ),
    dict(  # mapping.rs:1512-1563
        name="test_gapped_alignment", model=TEST_MODEL_10, bound=("test", -3.0, -10.0), gaps=(-2.0, -1.0, 0, 2),
        ref="TAT", pattern="TT", qual=0, expect=dict(positions_sorted=[0, 2, 5]),
    ),
    dict(  # mapping.rs:1566-1616 (gap in the middle: allowed)
        name="test_gapped_alignment_read_end/middle", model=TEST_MODEL_10, bound=("test", -6.0, -10.0), gaps=(-2.0, -1.0, 5, 2),
        ref="AAAAAAGGGGAAAAAA", pattern="AAAAAAAAAAAA", qual=0, expect=dict(nonempty=True),
    ),
    dict(  # mapping.rs:1617-1639 (gap near read end: not allowed)
        name="test_gapped_alignment_read_end/end", model=TEST_MODEL_10, bound=("test", -6.0, -10.0), gaps=(-2.0, -1.0, 5, 2),
        ref="AAAAAAGGGGAAAAAA", pattern="AGGGAAAAAA", qual=0, expect=dict(positions_sorted=[]),
    ),
    dict(  # mapping.rs:1642-1695 (one gap: allowed)
        name="test_gap_open_limit/one", model=TEST_MODEL_10, bound=("test", -6.0, -10.0), gaps=(-2.0, -1.0, 5, 1),
        ref="CTAGCCAGCGATTTACATGCTCTCGGAATATCGACATGTA", pattern="CTAGCCAGCGAACATGCTCTCGGAATATCGACATGTA", qual=0,
        expect=dict(contains_position=0),
    ),
    dict(  # mapping.rs:1697-1720 (two gaps: not allowed)
        name="test_gap_open_limit/two", model=TEST_MODEL_10, bound=("test", -6.0, -10.0), gaps=(-2.0, -1.0, 5, 1),
        ref="CTAGCCAGCGATTTACATGCTCTCGGAATATCGACATGTA", pattern="CTAGCCAGCGATTACATGCTCTCGGAATTCGACATGTA", qual=0,
        expect=dict(positions_sorted=[]),
    ),
    dict(  # mapping.rs:1724-1774
        name="test_vindija_pwm_alignment/1", model=("vindija",), bound=("test", -30.0, None), gaps=(-200.0, -100.0, 0, 2),
        ref="CCCCCC", pattern="TTCCCT", qual=40, expect=dict(score0=-4.641691, positions_sorted=[0]),
    ),
    dict(  # mapping.rs:1776-1801
        name="test_vindija_pwm_alignment/2", model=("vindija",), bound=("test", -30.0, None), gaps=(-200.0, -100.0, 0, 2),
        ref="CCCCCC", pattern="CCCCCC", qual=0, expect=dict(score0=0.0, positions_sorted=[0]),
    ),
    dict(  # mapping.rs:1807-1831
        name="test_vindija_pwm_alignment/3", model=("vindija",), bound=("test", -30.0, None), gaps=(-200.0, -100.0, 0, 2),
        ref="AAAAAA", pattern="AAGAAA", qual=0, expect=dict(score0_approx=-10.965062),
    ),
    dict(  # mapping.rs:1874-1934
        name="test_corner_cases", model=("vindija",), bound=("discrete", 0.01, 0.02), gaps=(("repr", 3.0), ("repr", 0.6), 0, 2),
        ref="GTTGTATTTTTAGTAGAGACAGGGTTTCATCATGTTGGCCAGAAAAAAAAAAAAAAAAAAAATTTGTATTTTTAGTAGAGACAGGCTTTCATCATGTTGGCCAG",
        pattern="GTTGTATTTTTAGTAGAGACAGGCTTTCATCATGTTGGCCAG", qual=40,
        expect=dict(scores_iter=[-10.936638, -39.474224, -10.965062], positions_sorted=[0, 62, 63], peek_positions=[0]),
    ),
    # test_cigar_indels, mapping.rs:1937-2229
    dict(name="test_cigar_indels/deletion", model=TEST_MODEL_10, bound=("test", -4.0, -10.0), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTAGCA", pattern="ATTACA", qual=0, expect=dict(best_cigar="4M1D2M")),
    dict(name="test_cigar_indels/deletion2", model=TEST_MODEL_10, bound=("test", -4.0, -10.0), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACAG", pattern="GATCAG", qual=0, expect=dict(best_cigar="3M2D3M", best_score=-4.0)),
    dict(name="test_cigar_indels/insertion", model=TEST_MODEL_10, bound=("test", -4.0, -10.0), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTAGCA", qual=0, expect=dict(best_cigar="5M1I2M", best_score=-3.0)),
    dict(name="test_cigar_indels/insertion2", model=TEST_MODEL_10, bound=("test", -4.0, -10.0), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTAGGCA", qual=0, expect=dict(best_cigar="5M2I2M", best_score=-4.0)),
    dict(name="test_cigar_indels/insertion3", model=TEST_MODEL_10, bound=("test", -5.0, None), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTAGTGCA", qual=0, expect=dict(best_cigar="5M3I2M", best_score=-5.0)),
    # test_md_tag, mapping.rs:2232-2440
    dict(name="test_md_tag/mutation", model=("test", -1.0, -2.0, 0.0), bound=("test", -1.0, -2.0), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTATA", qual=40, expect=dict(best_md="5C1")),
    dict(name="test_md_tag/deletion", model=("test", -1.0, -2.0, 0.0), bound=("test", -4.0, None), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTAGCA", pattern="ATTACA", qual=0, expect=dict(best_md="4^G2")),
    dict(name="test_md_tag/deletion2", model=("test", -1.0, -2.0, 0.0), bound=("test", -4.0, None), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACAG", pattern="GATCAG", qual=0, expect=dict(best_md="3^TA3")),
    dict(name="test_md_tag/insertion", model=("test", -1.0, -2.0, 0.0), bound=("test", -4.0, None), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTAGCA", qual=0, expect=dict(best_md="7")),
    dict(name="test_md_tag/insertion2", model=("test", -1.0, -2.0, 0.0), bound=("test", -4.0, None), gaps=(-2.0, -1.0, 0, 2),
         ref="GATTACA", pattern="GATTAGGCA", qual=0, expect=dict(best_md="7")),
    dict(  # mapping.rs:2443-2513
        name="test_reverse_strand_search_2", model=("test", -1.0, -1.0, 0.0), bound=("test", 0.0, -1.0), gaps=(-3.0, -1.0, 0, 2),
        ref="AAAGCGTTTGCG", pattern="TTT", qual=0, expect=dict(best_strand_positions=[(6, "F"), (0, "B")]),
    ),
    dict(  # mapping.rs:2516-2590
        name="test_edit_operations_reverse_strand", model=("test", -1.0, -1.0, 0.0), bound=("test", -1.0, -1.0), gaps=(-3.0, -1.0, 0, 2),
        ref="GATTACA", pattern="TAGT", qual=0,
        expect=dict(peek_strand_positions=[(1, "B")], peek_md_backward="1T2", peek_nm_backward=1),
    ),
    dict(  # mapping.rs:2593-2646
        name="test_n/all_n", model=("simple", "single_stranded", 0.475, 0.475, 0.001, 0.9, 0.02 / 3.0, False),
        bound=("test", -14.0, None), gaps=(("log2", 0.001), ("repr", 1.0), 0, 2),
        ref="GATTACAGATTACAGATTACA", pattern="NNNNNNNNNN", qual=40, expect=dict(n_hits=0),
    ),
    dict(  # mapping.rs:2648-2665
        name="test_n/one_n", model=("simple", "single_stranded", 0.475, 0.475, 0.001, 0.9, 0.02 / 3.0, False),
        bound=("test", -14.0, None), gaps=(("log2", 0.001), ("repr", 1.0), 0, 2),
        ref="GATTACAGATTACAGATTACA", pattern="AGATNACAG", qual=40, expect=dict(n_hits=1),
    ),
]

# test_bench (mapping.rs:2669-2956): parameters; data in tests/golden/ref_test_bench.json
BENCH_PARAMS = dict(
    model=("simple", "single_stranded", 0.475, 0.475, 0.001, 0.9, 0.02 / 3.0, False),
    bound=("discrete", 0.04, 0.02), gaps=(("log2", 0.00001), ("repr", 1.0), 5, 2), qual=40,
)

# tests/integration_tests.rs:143-171
INTEGRATION_PARAMS = dict(
    model=("simple", "single_stranded", 0.6, 0.55, 0.01, 1.0, 0.02 / 3.0, False),
    bound=("discrete", 0.03, 0.02), gaps=(("repr", 1.5), ("repr", 0.5), 5, 2),
)

# src/map/bi_d_array.rs:243-309
D_ARRAY_CASE = dict(
    model=("test", -1.0, -1.0, 0.0), bound=("test", 0.0, None), gaps=(("log2", 0.00001), ("repr", 1.0), 0, 2),
    ref="GATTACA", pattern="CCCCCCC", quals=[10, 40, 40, 40, 40, 10, 40], split=3,
    d_composite=[0.0, 0.0, -1.0, 0.0, 0.0, -1.0, -1.0], get_2_3=-2.0, get_0_6=0.0,
)

# README "50 % deamination" parameters (Readme.md:143-155)
def cli_params(library="single_stranded"):
    if library == "single_stranded":
        model = ("simple", "single_stranded", 0.5, 0.5, 0.02, 1.0, 0.02 / 3.0, False)
    else:
        model = ("simple", "double_stranded", 0.5, 0.5, 0.02, 1.0, 0.02 / 3.0, False)
    return dict(model=model, bound=("discrete", 0.03, 0.02), gaps=(("log2", 0.001), ("repr", 0.5), 5, 2))
