"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol that
include/pfetch.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fetching_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pfetch.h")).read()
    return sorted(set(re.findall(r"PF_API\s+[\w\s\*]+?\b(pf_\w+)\s*\(", text)))


def test_header_declares_what_python_binds():
    from fetching_b200 import _lib
    assert declared_symbols() == sorted(_lib.SYMBOLS)
    assert len(declared_symbols()) >= 16


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(os.path.join(ROOT, "fetching_b200", "lib", "libpfetch.so"))
    for name in declared_symbols():
        assert hasattr(raw, name), name
    assert lib.pf_abi_version() == 1


def test_struct_layout_matches_header():
    from fetching_b200._lib import ModelDesc
    # pf_model_desc: 4 x int32, u64, 2 x u32, 2 x u64, double, 2 pointers
    assert ctypes.sizeof(ModelDesc) == 16 + 8 + 8 + 16 + 8 + 16


def test_blasso_log_prior_host_function(py_models):
    from fetching_b200 import blasso_log_prior
    for theta, want in zip(py_models["bl_thetas"], py_models["bl_log_prior"]):
        got = blasso_log_prior(theta, float(py_models["bl_tau"]))
        assert abs(got - want) <= 1e-13 * abs(want)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fetching_b200 import FrontierEngine, PfetchError
    with pytest.raises(PfetchError, match="PF_ERR_NO_DEVICE"):
        FrontierEngine.normal(np.zeros((16, 2)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fetching_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "liboracle" not in text and "oracle/_ref" not in text, f


def test_normal_dataset_recipe_matches_oracle(orc):
    from fetching_b200 import models
    assert np.array_equal(models.normal_dataset(4097), orc.normal_dataset(4097))
