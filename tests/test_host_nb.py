"""CPU: the product's host side of the negative-binomial model (flagger_b200/csrc/hfg_nb.c + the model-type branches of
hfg_host_model.c) against the oracle, which is pinned against the unmodified reference (tests/test_oracle_nb.py).
There is no device path for this model yet (hfg_create rejects it); what is checked here is everything around it."""
import os

import numpy as np
import pytest

import golden_util
from flagger_b200 import _abi, api, synth

NB = _abi.MODEL_NEGATIVE_BINOMIAL


def _flat(a):
    return a.view(np.float64).reshape(-1)


def test_product_digamma_matches_reference_bits():
    z = np.load(os.path.join(golden_util.GOLDEN_DIR, "digamma.nb.npz"))
    for x, hi, lo in zip(z["x"], z["hi"], z["lo"]):
        assert api.digamma(x) == (hi, lo), x


@pytest.mark.parametrize("name", golden_util.NB_NAMES)
def test_host_nb_model_init_mstep_squarem_match_golden(name):
    g, wl = golden_util.load(name, ".nb.npz")
    cfg = g["cfg"]
    p0 = api.model_init(cfg, wl.region_coverages, wl.window_len)
    assert np.array_equal(_flat(p0), _flat(g["params0"]))
    p1, conv = api.mstep(cfg, p0, g["stats"], tol=1e-3)
    assert np.array_equal(_flat(p1), _flat(g["params1"])) and conv == bool(g["converged1"])
    for n in range(len(g["cand_rates"])):
        cand, rate, feasible = api.squarem(cfg, g["params0"], g["params1"], g["params2"], n)
        assert rate == g["cand_rates"][n] and bool(feasible) == bool(g["cand_feasible"][n])
        assert np.array_equal(_flat(cand), _flat(g["cand_params"][n]))


def test_host_nb_feasibility_bounds():
    cfg = _abi.make_config(n_regions=1, n_col_comps=2, model_type=NB)
    p = api.model_init(cfg, np.array([40], np.int32), 4000)
    assert api.params_feasible(cfg, p)
    for field, value in (("mean", 1.0), ("mean", 0.0), ("var", 0.0), ("weight", 1.5), ("weight", -0.1)):
        q = p.copy()
        q[field][0][3][1] = value
        assert not api.params_feasible(cfg, q), (field, value)


def test_host_nb_emission_table_matches_oracle_forward(orc):
    """The table reproduces the oracle's emissions: a forward pass re-done in numpy with table look-ups gives the oracle's
    scales, hence its log-likelihood, to the last bits of the additions."""
    wl = synth.small_mixed(n_regions=2, seed=61)
    wl.cov[3:6] = 250
    cfg = _abi.make_config(n_regions=2, n_col_comps=3, model_type=NB, adjust_contig_ends=False)
    p = api.model_init(cfg, wl.region_coverages, wl.window_len)
    table = api.nb_emission_table(cfg, p)
    assert table.shape == (2, 4, 251) and np.all(table > 0) and np.all(table <= 1.0)
    assert np.allclose(table[:, 1:3, :].sum(axis=2), 1.0, atol=1e-9)  # Dup / Hap: all the mass lies below 251
    e = orc.estep(cfg, wl, np.zeros((4, 4)), p, want_fb=True)
    assert e["rc"] == 0
    ch = wl.chunks[0]
    o, L = int(ch["offset"]), int(ch["n_windows"])
    x, reg = wl.cov[o:o + L].astype(int), wl.region[o:o + L].astype(int)
    f = table[reg[0], :, x[0]] * p["trans"][reg[0]][4, :4]
    scale0 = f.sum()
    assert scale0 == e["scales"][o]
    assert np.array_equal(f / scale0, e["fwd"][o])


def test_host_nb_stats_from_histogram_match_oracle(orc):
    """One chunk: the statistics are exactly what the oracle (= the reference) makes of that chunk's histogram.  Several
    chunks: the reference updates per chunk and adds, the product updates once from the summed histogram -- the update is
    linear in the bin mass, so only the rounding differs."""
    for n_chunks_expected, wl, R in ((1, synth.config1(seed=71), 1), (None, synth.small_mixed(n_regions=3, seed=72), 3)):
        K = api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        cfg = _abi.make_config(n_regions=R, n_col_comps=K, model_type=NB, mean_read_length=wl.avg_alignment_len)
        p = api.model_init(cfg, wl.region_coverages, wl.window_len)
        for it in range(3):
            e, hist = orc.nb_histogram(cfg, wl, np.zeros((4, 4)), p)
            assert e["rc"] == 0
            got = api.nb_stats_from_histogram(cfg, p, hist)
            want = e["stats"]
            for key in ("mean_num", "mean_den", "var_num", "var_den", "weight_num", "weight_den"):
                if wl.n_chunks == 1:
                    assert np.array_equal(got[key], want[key]), (key, it)
                else:
                    assert np.allclose(got[key], want[key], rtol=1e-11, atol=1e-12), (key, it)
            # histogram mass of a state == transition mass into it
            assert np.allclose(hist.sum(axis=2), want["trans_count"].sum(axis=1), rtol=1e-12)
            got["trans_count"] = want["trans_count"]
            p2, _ = api.mstep(cfg, p, got, tol=1e-3)
            p_ref, _ = orc.mstep(cfg, p, want, tol=1e-3)
            if wl.n_chunks == 1:
                assert np.array_equal(_flat(p2), _flat(p_ref))
            else:
                assert np.allclose(_flat(p2), _flat(p_ref), rtol=1e-10, atol=1e-300)
            p = p_ref
        if n_chunks_expected:
            assert wl.n_chunks == n_chunks_expected


def test_create_without_a_gpu_fails_loudly_for_the_negative_binomial_model_too():
    """No CPU fallback of any kind: on a box without a usable device hfg_create says so (on a GPU box it succeeds)."""
    import ctypes as C
    cfg = _abi.make_config(model_type=NB)
    ctx = C.c_void_p()
    rc = api.lib().hfg_create(C.byref(ctx), api.ptr(cfg))
    api.lib().hfg_last_error.restype = C.c_char_p
    if rc == 0:
        api.lib().hfg_destroy(ctx)
    else:
        assert rc == 2 and not ctx.value and b"no CPU fallback" in api.lib().hfg_last_error(None)  # HFG_ERR_CUDA


def test_host_nb_table_reports_nan_like_the_reference_exit():
    """theta outside (0, 1) makes the pmf NaN: the reference exits with "prob is NAN"; the table builder says HFG_ERR_NAN."""
    cfg = _abi.make_config(n_regions=1, n_col_comps=2, model_type=NB)
    p = api.model_init(cfg, np.array([40], np.int32), 4000)
    p["mean"][0][2][0] = 1.5
    with pytest.raises(api.HfgError):
        api.nb_emission_table(cfg, p)
    with pytest.raises(api.HfgError):  # wrong model type for these entry points
        api.nb_emission_table(_abi.make_config(n_regions=1, n_col_comps=2), p)
