"""CPU: host-side data layout of the CUDA path (flagger_b200/csrc/hfg_layout.c): segmentation invariants, packed
observation words, the contig-end factor beta against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from flagger_b200 import _abi, api, synth
from flagger_b200._abi import ptr


@pytest.mark.parametrize("factory,n_regions,capacity", [
    (lambda: synth.small_mixed(n_regions=3, seed=1), 3, 512),
    (lambda: synth.small_mixed(n_regions=7, seed=2), 7, 2048),
    (lambda: synth.config1(), 1, 512),
    (lambda: synth.config3(n_contigs=300, seed=3), 1, 1024),
    (lambda: synth.config2(total_bp=200_000_000, seed=4), 1, 4096),
    (lambda: synth.config4(total_bp=100_000_000, seed=5), 7, 1536),
])
def test_layout_invariants(factory, n_regions, capacity):
    wl = factory()
    for adjust in (True, False):
        cfg = _abi.make_config(n_regions=n_regions, adjust_contig_ends=adjust, mean_read_length=wl.avg_alignment_len)
        ok, summary = api.layout_check(cfg, wl, capacity)
        assert ok
        n_seg, smax, n_edge, W, n_keys, n_tiles = (int(v) for v in summary)
        assert W == wl.n_windows and n_seg <= capacity and n_seg * smax >= W
        assert 1 <= n_keys <= W and n_tiles <= W // 16 + n_keys + 1
        assert (n_edge == 0) == (not adjust)


def test_layout_rejects_bad_input():
    wl = synth.small_mixed(n_regions=3, seed=1)
    cfg = _abi.make_config(n_regions=2)  # windows carry region 2
    ok, _ = api.layout_check(cfg, wl, 512)
    assert not ok


def test_more_region_runs_than_segment_slots_is_an_error_not_a_hang():
    """A segment never straddles a region change, so runs of equal region index bound the segment count from below:
    alternating regions with fewer slots than windows used to loop forever in the segment-length search."""
    wl = synth.small_mixed(n_regions=2, seed=3)
    wl.region[:] = (np.arange(wl.n_windows) % 2).astype(np.uint8)
    cfg = _abi.make_config(n_regions=2)
    ok, _ = api.layout_check(cfg, wl, 512)
    assert not ok
    ok, summary = api.layout_check(cfg, wl, 2048)  # enough slots: one window per segment
    assert ok and summary[0] == wl.n_windows and summary[1] == 1


def test_beta_matches_oracle(orc):
    f = orc.lib.orc_beta
    f.restype = C.c_double
    wl = synth.small_mixed(n_regions=1, seed=3)
    for frac, Lr in ((0.95, 15000), (1.0, 15000), (0.8, 40000), (0.95, 0), (0.5, 1000)):
        cfg = _abi.make_config(min_read_fraction_at_ends=frac, mean_read_length=Lr)
        for c in range(wl.n_chunks):
            ch = wl.chunks[c:c + 1]
            for i in range(int(ch["n_windows"][0])):
                a = api.beta(cfg, ch, i)
                b = f(ptr(cfg), ptr(ch), C.c_int(i))
                assert a == b or (np.isnan(a) and np.isnan(b))
