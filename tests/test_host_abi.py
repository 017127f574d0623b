"""CPU: the C-ABI library loads and exports every symbol include/hfg.h declares; host-side mirrors (model init,
M-step, component count) agree with the oracle and the golden vectors; the compute entry points fail loudly without a
GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util
from flagger_b200 import _abi, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat(a):
    return a.view(np.float64).reshape(-1)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hfg.h")).read() + open(os.path.join(ROOT, "include", "hfg_io.h")).read()
    declared = set(re.findall(r"\b(hfg_[a-z_0-9]+)\s*\(", header))
    declared -= {"hfg_ctx", "hfg_cov_data", "hfg_io"}
    assert len(declared) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, "flagger_b200", "libhfg.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hfg.h but not exported"
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header():
    assert _abi.config_dtype.itemsize == 72
    assert _abi.chunk_desc_dtype.itemsize == 32
    assert _abi.region_params_dtype.itemsize == 8 * 219
    assert _abi.region_stats_dtype.itemsize == 8 * 402
    assert _abi.region_stats_dtype.fields["lambda_num"][1] == 128
    assert _abi.region_params_dtype.fields["trans"][1] == 8 * (2 + 3 * 64)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_host_model_init_and_mstep_match_golden(name):
    g, wl = golden_util.load(name)
    cfg = g["cfg"]
    assert api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages) == int(g["K"])
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    assert np.array_equal(_flat(p0), _flat(g["params0"]))
    p1, conv = api.mstep(cfg, p0, g["stats"], tol=1e-3)
    assert np.array_equal(_flat(p1), _flat(g["params1"])) and conv == bool(g["converged1"])


def test_mstep_gates_and_convergence(orc):
    """MIN_COUNT_FOR_PARAMETER_UPDATE gating and the convergence flag, against the oracle on crafted statistics."""
    cfg = _abi.make_config(n_regions=2, n_col_comps=3)
    p = api.model_init(cfg, [40, 30], 4000)
    rng = np.random.default_rng(0)
    for scale in (0.0, 1.0, 1e3):
        stats = np.zeros(2, dtype=_abi.region_stats_dtype)
        flat = _flat(stats)
        flat[:] = rng.random(flat.shape) * scale
        a, ca = api.mstep(cfg, p, stats, tol=1e-3)
        b, cb = orc.mstep(cfg, p, stats, tol=1e-3)
        assert np.array_equal(_flat(a), _flat(b)) and ca == cb
    # zero statistics: emissions keep their values (counts <= 10), transition rows fall back to the pseudo-counts
    stats = np.zeros(2, dtype=_abi.region_stats_dtype)
    a, conv = api.mstep(cfg, p, stats, tol=1e-3)
    assert not conv and np.array_equal(a["mean"], p["mean"]) and np.allclose(a["trans"][:, :4, :4], 0.25 * (1 - 1e-4))
    # a second M-step from there changes nothing => converged
    a2, conv2 = api.mstep(cfg, a, stats, tol=1e-3)
    assert conv2 and np.array_equal(_flat(a2), _flat(a))


def test_start_only_mode_matches_reference(orc, ref):
    """start-only mode rescales the baseline by windowLen/avgAlignmentLen, and the region scales by its inverse
    (src/hmm_flagger.c:186-199): the Hap means come out as the region coverages again -- a reference quirk we mirror."""
    cfg = _abi.make_config(n_regions=2, n_col_comps=4, mean_read_length=20000)
    a = api.model_init(cfg, [40, 52], 4000, start_only=True)
    b = orc.model_init(cfg, [40, 52], 4000, start_only=True)
    assert np.array_equal(_flat(a), _flat(b))
    assert np.allclose(a["mean"][:, 2, 0], [40.0, 52.0])
    if ref is not None:
        assert np.array_equal(_flat(a), _flat(ref.model_init(cfg, [40, 52], 4000, start_only=True)))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.HfgError) as e:
        api.HmmFlaggerGPU(_abi.make_config())
    assert e.value.code == _abi.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_invalid_configs_rejected():
    lib = api.lib()
    h = ctypes.c_void_p()
    bad = _abi.make_config(n_regions=0)
    assert lib.hfg_create(ctypes.byref(h), _abi.ptr(bad)) == _abi.ERR_INVALID
    bad = _abi.make_config(n_col_comps=17)
    assert lib.hfg_create(ctypes.byref(h), _abi.ptr(bad)) == _abi.ERR_INVALID
    bad = _abi.make_config(model_type=3)  # 0 trunc_exp_gaussian, 1 gaussian, 2 negative_binomial
    assert lib.hfg_create(ctypes.byref(h), _abi.ptr(bad)) == _abi.ERR_INVALID
    assert b"model_type" in lib.hfg_last_error(None)
