"""The C-ABI library loads without a GPU and exports every symbol include/cgpe.h declares; the
ctypes structures have the C layout.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from coarse_grain_polyelectrolyte_b200 import _abi, build as cuda_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cgpe.h")


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CGPE_API\s+[\w\s\*]+?\b(cgpe_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    return cuda_build.build()


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 28
    assert sorted("cgpe_" + n for n in _abi.SIGNATURES) == names


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    b = _abi.Binding(lib, "cgpe_")
    assert b.abi_version() == _abi.ABI_VERSION


def test_oracle_mirrors_the_abi(oracle_mod):
    lib = oracle_mod.binding().lib
    for name in _abi.SIGNATURES:
        assert hasattr(lib, "orc_" + name), name


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "cgpe.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(cgpe_config), sizeof(cgpe_reaction),"
        " sizeof(cgpe_sample_params), sizeof(cgpe_trial_record), offsetof(cgpe_config, bond_k), offsetof(cgpe_config, wca_eps),"
        " offsetof(cgpe_config, reactions), offsetof(cgpe_sample_params, p_burst1));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_abi.Config), C.sizeof(_abi.Reaction), C.sizeof(_abi.SampleParams), _abi.TRIAL_DTYPE.itemsize,
            _abi.Config.bond_k.offset, _abi.Config.wca_eps.offset, _abi.Config.reactions.offset, _abi.SampleParams.p_burst1.offset]
    assert got == want


def test_no_cpu_fallback_in_product(lib_path):
    """Without a CUDA device cgpe_create must fail loudly (status CGPE_ERR_CUDA), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from coarse_grain_polyelectrolyte_b200.engine import Engine
    cfg = _abi.make_config(n_replicas=1, n_beads=4, n_types=1, box_l=10.0, time_step=1e-3)
    with pytest.raises(_abi.CgpeError) as ei:
        Engine(cfg)
    assert ei.value.status == _abi.ERR_CUDA and "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    bad = []
    for base in ("coarse_grain_polyelectrolyte_b200", "src"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|orc_\w+\(", text, re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
