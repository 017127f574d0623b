"""ctypes / numpy mirrors of the plain-C layouts declared in include/hfg.h.

Nothing here computes: the structured dtypes below have exactly the size, field order and
alignment of the C structs so that numpy arrays can be handed to the C-ABI by pointer.
"""
import ctypes as C

import numpy as np

NUM_STATES = 4
MAX_COMPS = 16
MAX_REGIONS = 64

MODEL_TRUNC_EXP_GAUSSIAN = 0
MODEL_GAUSSIAN = 1
MODEL_NEGATIVE_BINOMIAL = 2  # host functions only so far (include/hfg.h)
NB_TABLE_X, NB_BINS = 251, 250  # HFG_NB_TABLE_X, HFG_NB_BINS

OK, ERR_INVALID, ERR_CUDA, ERR_SCALE_UNDERFLOW, ERR_NAN, ERR_NOMEM = range(6)

STATE_NAMES = ("Err", "Dup", "Hap", "Col")

config_dtype = np.dtype(
    [
        ("model_type", "<i4"),
        ("n_regions", "<i4"),
        ("n_comps", "<i4", (NUM_STATES,)),
        ("adjust_contig_ends", "<i4"),
        ("mean_read_length", "<i4"),
        ("min_read_fraction_at_ends", "<f8"),
        ("max_high_mapq_ratio", "<f8"),
        ("min_high_mapq_ratio", "<f8"),
        ("min_highly_clipped_ratio", "<f8"),
        ("device", "<i4"),
        ("reserved", "<i4"),
    ],
    align=True,
)

chunk_desc_dtype = np.dtype(
    [
        ("ctg_len", "<i4"),
        ("s", "<i4"),
        ("e", "<i4"),
        ("window_len", "<i4"),
        ("n_windows", "<i4"),
        ("reserved", "<i4"),
        ("offset", "<i8"),
    ],
    align=True,
)

region_params_dtype = np.dtype(
    [
        ("lambda", "<f8"),
        ("trunc_point", "<f8"),
        ("mean", "<f8", (NUM_STATES, MAX_COMPS)),
        ("var", "<f8", (NUM_STATES, MAX_COMPS)),
        ("weight", "<f8", (NUM_STATES, MAX_COMPS)),
        ("trans", "<f8", (NUM_STATES + 1, NUM_STATES + 1)),
    ],
    align=True,
)

region_stats_dtype = np.dtype(
    [
        ("trans_count", "<f8", (NUM_STATES, NUM_STATES)),
        ("lambda_num", "<f8"),
        ("lambda_den", "<f8"),
        ("mean_num", "<f8", (NUM_STATES, MAX_COMPS)),
        ("mean_den", "<f8", (NUM_STATES, MAX_COMPS)),
        ("var_num", "<f8", (NUM_STATES, MAX_COMPS)),
        ("var_den", "<f8", (NUM_STATES, MAX_COMPS)),
        ("weight_num", "<f8", (NUM_STATES, MAX_COMPS)),
        ("weight_den", "<f8", (NUM_STATES, MAX_COMPS)),
    ],
    align=True,
)

assert config_dtype.itemsize == 72
assert chunk_desc_dtype.itemsize == 32
assert region_params_dtype.itemsize == 8 * (2 + 3 * NUM_STATES * MAX_COMPS + 25)
assert region_stats_dtype.itemsize == 8 * (16 + 2 + 6 * NUM_STATES * MAX_COMPS)


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def make_config(
    n_regions=1,
    n_col_comps=4,
    model_type=MODEL_TRUNC_EXP_GAUSSIAN,
    adjust_contig_ends=True,
    mean_read_length=15000,
    min_read_fraction_at_ends=0.95,
    max_high_mapq_ratio=0.25,
    min_high_mapq_ratio=0.75,
    min_highly_clipped_ratio=1.0,
    device=0,
):
    """hfg_config with the reference CLI defaults (src/hmm_flagger.c:613-650; hifi preset :28)."""
    cfg = np.zeros(1, dtype=config_dtype)
    cfg["model_type"] = model_type
    cfg["n_regions"] = n_regions
    cfg["n_comps"][0] = (1, 1, 1, n_col_comps)
    cfg["adjust_contig_ends"] = 1 if adjust_contig_ends else 0
    cfg["mean_read_length"] = mean_read_length
    cfg["min_read_fraction_at_ends"] = min_read_fraction_at_ends
    cfg["max_high_mapq_ratio"] = max_high_mapq_ratio
    cfg["min_high_mapq_ratio"] = min_high_mapq_ratio
    cfg["min_highly_clipped_ratio"] = min_highly_clipped_ratio
    cfg["device"] = device
    return cfg


def stats_as_flat(stats):
    """View an array of hfg_region_stats as a flat float64 vector."""
    return stats.view(np.float64).reshape(-1)


def params_as_flat(params):
    return params.view(np.float64).reshape(-1)
