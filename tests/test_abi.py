"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mapad_gpu.h declares, its
PODs have the sizes the Python mirror assumes, and the GPU entry points refuse to run without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mapad_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    header = open(os.path.join(ROOT, "include", "mapad_gpu.h")).read()
    declared = set(re.findall(r"\b(mapad_[a-z0-9_]+)\s*\(", header)) - {"mapad_sdm_get_fn", "mapad_sdm_start_fn"}
    assert declared == set(api.EXPORTED_SYMBOLS), declared ^ set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.mapad_abi_version() == 1


def test_pod_sizes():
    L = api.lib()
    mirrors = [abi.Params, abi.Reads, abi.EditOp, abi.Hit, abi.Alt, abi.Record, abi.Results, abi.IndexView]
    for what, cls in enumerate(mirrors):
        assert int(L.mapad_abi_sizeof(what)) == C.sizeof(cls), (cls.__name__, int(L.mapad_abi_sizeof(what)), C.sizeof(cls))
    assert C.sizeof(abi.EditOp) == 4 and C.sizeof(abi.Hit) == 40


def test_host_side_scoring_matches_oracle():
    from helpers import oracle_params, product_params
    from ref_cases import cli_params, INTEGRATION_PARAMS
    for spec in (cli_params("single_stranded"), cli_params("double_stranded"), INTEGRATION_PARAMS):
        P, O = product_params(spec), oracle_params(spec)
        assert np.float32(P.representative_mismatch_penalty) == np.float32(O.repr_mm)
        for L in (10, 17, 25, 50, 100, 150, 300):
            assert api.allowed_mismatches(P, L) == O.discrete_get(L)
        for i, L, f, t, q in [(0, 30, "C", "T", 40), (29, 30, "G", "A", 30), (5, 50, "C", "C", 2), (7, 50, "A", "G", 20), (3, 40, "T", "T", 0)]:
            assert np.float32(api.sdm_get(P, i, L, f, t, q)) == np.float32(O.sdm_get(i, L, f, t, q))
    P = api.params_from_cli()
    O = oracle_params(cli_params("single_stranded"))
    assert np.float32(P.representative_mismatch_penalty) == np.float32(O.repr_mm)
    assert P.gap_dist_ends == 5 and P.max_num_gaps_open == 2


def test_no_device_no_compute():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    index = api.Index.build([("c", "ACGTACGTTTGACC")])
    with pytest.raises(api.MapadError) as e:
        api.Mapper(index, api.params_from_cli())
    assert e.value.code == -2  # MAPAD_ENODEV: no CPU fallback


def _c_struct_fields(header, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "float a, b, c" / "const uint8_t* seq" / "uint64_t less[8]" / "mapad_alt alts[2]"
        first, *rest = [x.strip() for x in decl.split(",")]
        fields.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?$", first)[0])
        fields += [re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?$", r)[0] for r in rest]
    return fields


def test_rust_binding_covers_header():
    """bindings/rust/mapad-gpu-sys (the thin FFI crate of the north star; not compilable here: no Rust toolchain) declares
    every function of include/mapad_gpu.h and mirrors every POD field for field, in order."""
    header = open(os.path.join(ROOT, "include", "mapad_gpu.h")).read()
    rust = open(os.path.join(ROOT, "bindings", "rust", "mapad-gpu-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"\b(mapad_[a-z0-9_]+)\s*\(", header)) - {"mapad_sdm_get_fn", "mapad_sdm_start_fn"}
    bound = set(re.findall(r"pub fn (mapad_[a-z0-9_]+)\s*\(", rust))
    assert declared == bound, declared ^ bound
    for name in ("mapad_params", "mapad_index_view", "mapad_reads", "mapad_edit_op", "mapad_hit", "mapad_alt", "mapad_record", "mapad_results"):
        want = _c_struct_fields(header, name)
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rust, re.S).group(1)
        got = re.findall(r"pub ([a-z0-9_]+):", body)
        assert want == got, (name, want, got)
    for const, value in re.findall(r"(MAPAD_[A-Z0-9_]+) = (-?\d+)u?", header):
        m = re.search(r"pub const %s: \w+ = (-?\d+);" % const, rust)
        assert m and int(m.group(1)) == int(value), const
