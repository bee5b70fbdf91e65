"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly
what include/bf_b200.h declares; the host-side mirror imports; the product never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from brainforge_b200 import _lib
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert _lib.lib.bf_version() >= 100
    assert _lib.last_error() == ""


def test_header_cites_reference_for_every_entry_point_group():
    text = open(os.path.join(ROOT, "include", "bf_b200.h")).read()
    for cite in ("atomic/core_op.py", "atomic/tensor_op.py", "atomic/activation_op.py", "atomic/recurrent_op.py",
                 "metrics/costs.py", "optimizers/", "evolution/_evolution.py", "learner/neuroevolution.py"):
        assert cite in text, cite


def test_error_convention_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from brainforge_b200 import _lib
    L = _lib.lib
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    # conv depth mismatch -> BF_ERR_VALUE -> ValueError with the reference's message (tensor_op.py:18-21)
    rc = L.bf_conv_forward(one, one, null, one, 1, 2, 5, 5, 1, 3, 3, 3, 0, 0, 0, null)
    assert rc == _lib.ERR_VALUE
    with pytest.raises(ValueError, match="input depth: 2 != 3 :filter depth"):
        _lib.check(rc)
    # pool with non-divisible input -> ValueError-class (layers/tensor.py:21-24 raises at connect)
    rc = L.bf_pool_forward(one, one, one, 1, 1, 5, 4, 2, null)
    assert rc == _lib.ERR_VALUE
    rc = L.bf_opt_step(99, one, one, null, null, 4, 0.1, 1.0, 0, 0, 0, 0, null)
    assert rc == _lib.ERR_ARG
    with pytest.raises(RuntimeError):
        _lib.check(rc)
    assert L.bf_pool_mask_bytes(2, 3, 8, 6, 2) == 2 * 3 * 4 * 3
    assert L.bf_pool_mask_bytes(1, 1, 9, 9, 3) == 9 * 8


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under brainforge_b200/ may reference it."""
    pkg = os.path.join(ROOT, "brainforge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(dirpath, f)
                assert "bf_oracle" not in text, os.path.join(dirpath, f)


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import brainforge_b200 as bf
    from brainforge_b200.layers import Dense
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bf.Backpropagation(input_shape=(4,), layerstack=[Dense(3)], cost="mse", optimizer="sgd")
