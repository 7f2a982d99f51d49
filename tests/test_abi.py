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


def test_cli_spec_equals_params_from_cli():
    """specs.product_params(specs.cli_spec(lib)) — what bench.py and the GPU tests use — is byte for byte what the C-ABI's own
    mapad_params_from_cli derives from the same flags (src/main.rs:418-499)."""
    from mapad_b200 import specs
    for lib in ("single_stranded", "double_stranded"):
        a = specs.product_params(specs.cli_spec(lib))
        b = api.params_from_cli(library=lib, p=0.03, f=0.5, t=0.5, d=0.02, s=1.0, D=0.02, i=0.001, x=0.5)
        assert bytes(a) == bytes(b), lib


def test_product_modules_do_not_load_the_oracle():
    """The oracle is test infrastructure: importing the product package, its spec helpers, the measurement tools' imports and
    bench.py itself must not load it (bench.py only does inside its CPU legs)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests'); sys.argv = ['bench.py']\n"
            "import mapad_b200.api, mapad_b200.specs, mapad_b200.cli, mapad_b200.sharding, mapad_b200.workloads, compare\n"
            "import importlib.util as u; s = u.spec_from_file_location('bench_mod', %r + '/bench.py'); m = u.module_from_spec(s); s.loader.exec_module(m)\n"
            "mapad_b200.specs.product_params(mapad_b200.specs.cli_spec('single_stranded'))\n"
            "bad = [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.') or k == 'helpers']\n"
            "assert not bad, bad\n" % (ROOT, ROOT, ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
