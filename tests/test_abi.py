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


def test_coherent_state_helpers_match_oracle():
    """the plain-torch mirrors of CoherentStatesOverlap / CoherentStatesWavefunction (propagators.py:124-290) against
    the numpy restatement, including a rank-deficient width matrix (zero-mode invariance, test_propagators.py:73-113)"""
    import numpy as np
    import torch
    from oracle import oracle
    from semiclassical_b200 import propagators
    torch.set_default_dtype(torch.float64)
    rng = np.random.default_rng(11)
    d, ni, nj = 4, 5, 3
    V, _ = np.linalg.qr(rng.standard_normal((d, d)))
    for w in (np.array([0.5, 1.0, 2.0, 3.0]), np.array([0.0, 1.0, 2.0, 3.0])):
        G = (V * w[None, :]) @ V.T
        G = 0.5 * (G + G.T)
        qi, pi, qj, pj = (rng.standard_normal((d, n)) for n in (ni, ni, nj, nj))
        cs = propagators.CoherentStatesOverlap(torch.from_numpy(G), torch.from_numpy(G))
        O = cs(*(torch.from_numpy(x) for x in (qi, pi, qj, pj))).numpy()
        assert np.abs(O - oracle.overlap(G, G, qi, pi, qj, pj)).max() < 1e-13
        v = rng.standard_normal(ni) + 1j * rng.standard_normal(ni)
        x = rng.standard_normal((d, 7))
        csw = propagators.CoherentStatesWavefunction(torch.from_numpy(G))
        phi = csw(torch.from_numpy(qi), torch.from_numpy(pi), torch.from_numpy(v), torch.from_numpy(x)).numpy()
        y = np.concatenate((qi, pi, np.zeros((1, ni))), axis=0)
        assert np.abs(phi - oracle.hk_wavefunction(G, y, v, x)).max() < 1e-13
