"""Case specs -> C-ABI parameter PODs, with f32 arithmetic identical to the reference's.

A spec is dict(model=..., bound=..., gaps=...[, abort=..., limits=...]) as used by the test-suite (tests/ref_cases.py), the
bench and the measurement tools.  This module belongs to the product package and never touches the CPU oracle: derived
values come from the library itself (`mapad_sdm_representative_mismatch_penalty`) and from the host libm (`log2f`, what the
reference's `f32::log2` calls).
"""
import ctypes as _C
import ctypes.util as _util

import numpy as np

f32 = np.float32

_libm = _C.CDLL(_util.find_library("m") or "libm.so.6")
_libm.log2f.restype = _C.c_float
_libm.log2f.argtypes = [_C.c_float]


def log2f(x):
    return float(_libm.log2f(float(f32(x))))


def model_div(model):
    """`0.02 / 3.0` in the reference is an f32 division (src/main.rs:452)."""
    if model[0] == "simple":
        m = list(model)
        if isinstance(m[6], float) and abs(m[6] - 0.02 / 3.0) < 1e-12:
            m[6] = float(f32(0.02) / f32(3.0))
        return tuple(m)
    return model


def resolve_gap(v, repr_mm):
    if isinstance(v, tuple):
        if v[0] == "repr":
            return float(f32(v[1]) * f32(repr_mm)) if v[1] != 1.0 else float(f32(repr_mm))
        if v[0] == "log2":
            return log2f(v[1])
        raise ValueError(v)
    return float(v)


def cli_spec(library="single_stranded"):
    """The flag set of BASELINE.md: -p 0.03 -f 0.5 -t 0.5 -d 0.02 -s 1.0 -D 0.02 -i 0.001 -x 0.5 (gap_dist_ends 5, 2 gaps)."""
    model = ("simple", library, 0.5, 0.5, 0.02, 1.0, 0.02 / 3.0, False)
    return dict(model=model, bound=("discrete", 0.03, 0.02), gaps=(("log2", 0.001), ("repr", 0.5), 5, 2))


def product_params(spec):
    """spec -> mapad_b200.abi.Params (the C-ABI POD)."""
    from . import abi, api

    model = model_div(spec["model"])
    P = abi.Params()
    if model[0] == "test":
        P.model_kind = abi.MODEL_TEST
        P.test_deam_score, P.test_mm_score, P.test_match_score = model[1], model[2], model[3]
    elif model[0] == "vindija":
        P.model_kind = abi.MODEL_VINDIJA_PWM
    else:
        P.model_kind = abi.MODEL_SIMPLE_ADNA
        P.library = abi.LIB_SINGLE_STRANDED if model[1] == "single_stranded" else abi.LIB_DOUBLE_STRANDED
        P.five_prime_overhang, P.three_prime_overhang = model[2], model[3]
        P.ds_deamination_rate, P.ss_deamination_rate, P.divergence = model[4], model[5], model[6]
        P.ignore_base_quality = int(model[7])
    repr_mm = api.representative_mismatch_penalty(P)
    P.representative_mismatch_penalty = repr_mm
    b = spec["bound"]
    if b[0] == "test":
        P.bound_kind = abi.BOUND_TEST
        P.test_threshold = b[1]
        P.test_representative_mm = repr_mm if b[2] is None else b[2]
    elif b[0] == "discrete":
        P.bound_kind = abi.BOUND_DISCRETE
        P.poisson_threshold, P.base_error_rate = b[1], b[2]
    else:
        P.bound_kind = abi.BOUND_CONTINUOUS
        P.cutoff, P.exponent = b[1], b[2]
    g = spec["gaps"]
    P.penalty_gap_open = resolve_gap(g[0], repr_mm)
    P.penalty_gap_extend = resolve_gap(g[1], repr_mm)
    P.gap_dist_ends, P.max_num_gaps_open = g[2], g[3]
    P.stack_limit_abort = int(spec.get("abort", False))
    if "limits" in spec:
        P.stack_limit, P.edit_tree_limit = spec["limits"]
    return P
