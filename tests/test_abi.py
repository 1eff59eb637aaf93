"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
that include/semiclassical_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from semiclassical_b200 import _native


def header_symbols():
    text = open(_native.HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = _native.build()
    assert os.path.exists(path)
    L = _native.lib()
    assert L.sc_abi_version() == 1


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_native.build())
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), "missing export: " + name


def test_binding_covers_header():
    assert set(header_symbols()) == set(_native.exported_symbols())


def test_invalid_arguments_are_rejected_without_gpu():
    L = _native.lib()
    assert L.sc_engine_num_trajectories(None) == -1
    rc = L.sc_engine_create(None, None)
    assert rc == _native.SC_ERR_INVALID
    with pytest.raises(AssertionError):
        _native.check(rc)


def test_no_cpu_fallback():
    """the product path refuses to run without a CUDA device instead of silently falling back"""
    import torch
    from semiclassical_b200 import propagators
    G = torch.eye(2, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        propagators.HermanKlukPropagator(G, G, device='cpu')


def test_product_does_not_import_oracle():
    root = os.path.dirname(_native.HERE)
    for dirpath, _, files in os.walk(_native.HERE):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/make_golden.py", ""), f"{f} mentions the oracle"


def test_dynamics_driver_host_logic(tmp_path):
    """`semi dynamics` driver (semiclassical_b200/dynamics.py): model-file parsing as cli.py:229-283, configuration
    errors, and no CPU path"""
    import numpy as np
    import pytest
    from semiclassical_b200 import dynamics, workloads
    rows = workloads._AS5_ROWS
    model_file = tmp_path / "AS_model.dat"
    np.savetxt(model_file, np.column_stack((rows, np.full(len(rows), 0.02))), fmt="%.10f", header="omega S nac chi")
    pot, q0, p0, G0, zpt = dynamics._as_model(str(model_file))
    m = workloads.as_5modes(0.02)
    assert np.allclose(q0.numpy(), m.q0, rtol=0, atol=1e-14) and float(abs(p0).max()) == 0.0
    assert np.allclose(np.diag(G0.numpy()), m.omega, rtol=0, atol=1e-18) and abs(zpt - m.en_zpt) < 1e-16
    assert pot.dimensions() == 5
    task = {"potential": {"type": "no such potential"}, "results": {}}
    with pytest.raises(dynamics.ConfigurationError):
        dynamics.run_semiclassical_dynamics(task, device="cuda:0")
    task = {"potential": {"type": "anharmonic AS", "model_file": str(model_file)}, "num_steps": 3, "time_step_fs": 0.01,
            "num_trajectories": 10, "batch_size": 10, "results": {"correlations": str(tmp_path / "c.npz")}}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dynamics.run_semiclassical_dynamics(task, device="cpu")
