"""Synthetic HMM-Flagger workloads (window level) for tests and bench.py.

Shapes follow SURVEY.md section 8(d) / BASELINE.json `configs`; the chunk layout follows the reference's
index builder (submodules/chunk/chunk.c:240-294) and the window values what Chunk_addWindow would
produce (chunk.c:393-441: mean -> round -> clip to 250).  Data are generated directly at window
resolution; `write_cov` expands them to a run-length .cov(.gz) for the plumbing-size configs.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _abi

# T2T-CHM13 chromosome lengths (Mbp), chr1..22, X
_CHR_MBP = [248.4, 242.7, 201.1, 193.6, 182.0, 172.1, 160.6, 146.3, 150.6, 134.8, 135.1, 133.3,
            113.6, 101.2, 99.8, 96.3, 84.3, 80.5, 61.7, 66.2, 45.1, 51.3, 154.3]

HIFI_ALPHA = np.array(  # misc/alpha_tsv/HiFi_DC_1.2/*.tsv, rows = preState, cols = state
    [[0.753, 0.000, 0.236, 0.000],
     [0.000, 0.464, 0.440, 0.000],
     [0.527, 0.162, 0.010, 0.218],
     [0.000, 0.000, 0.041, 0.206]], dtype=np.float64)

ANNOTATION_NAMES = ["no_annotation", "whole_genome", "bsat", "hsat1A", "hsat1B", "hsat2", "hsat3", "ct"]


@dataclass
class Workload:
    name: str
    window_len: int
    chunk_len: int
    avg_alignment_len: int
    region_coverages: np.ndarray      # int32 [R]
    contig_names: list                # per chunk
    chunks: np.ndarray                # hfg_chunk_desc [C]
    cov: np.ndarray                   # uint16 [W]
    cov_high_mapq: np.ndarray         # uint16 [W]
    cov_high_clip: np.ndarray         # uint16 [W]
    region: np.ndarray                # uint8 [W]
    truth: np.ndarray                 # int8 [W] (-1 = none)
    annotation_names: list = field(default_factory=lambda: ["no_annotation", "whole_genome"])

    @property
    def n_windows(self):
        return int(self.cov.shape[0])

    @property
    def n_chunks(self):
        return int(self.chunks.shape[0])

    @property
    def n_regions(self):
        return int(self.region_coverages.shape[0])

    def total_bases(self):
        return int((self.chunks["e"].astype(np.int64) - self.chunks["s"] + 1).sum())

    def subset(self, chunk_indices, name=None):
        """A workload holding only the given chunks (used to shard across ranks / bound CPU samples)."""
        idx = np.asarray(chunk_indices, dtype=np.int64)
        ch = self.chunks[idx].copy()
        sel = np.concatenate([np.arange(c["offset"], c["offset"] + c["n_windows"]) for c in ch]) if len(ch) else \
            np.zeros(0, np.int64)
        off = 0
        for c in ch:
            c["offset"] = off
            off += int(c["n_windows"])
        return Workload(name or self.name, self.window_len, self.chunk_len, self.avg_alignment_len,
                        self.region_coverages, [self.contig_names[i] for i in idx], ch, self.cov[sel].copy(),
                        self.cov_high_mapq[sel].copy(), self.cov_high_clip[sel].copy(), self.region[sel].copy(),
                        self.truth[sel].copy(), self.annotation_names)


def chunk_layout(contig_lens, chunk_len):
    """(contig index, s, e) per chunk, as ChunksCreator_createCovIndex builds them (chunk.c:240-294)."""
    out = []
    for ci, L in enumerate(contig_lens):
        L = int(L)
        s = 0
        e = L - 1 if L < 2 * chunk_len else chunk_len - 1
        out.append((ci, s, e))
        while e < L - 1:
            s = e + 1
            e = L - 1 if L < (s - 1) + 2 * chunk_len else (s - 1) + chunk_len
            out.append((ci, s, e))
    return out


def _state_path(rng, n, event_every=200):
    """Mostly Hap with one 1-7 window Err/Dup/Col run per ~event_every windows."""
    st = np.full(n, 2, dtype=np.int8)
    pos = int(rng.integers(0, event_every))
    while pos < n:
        ln = int(rng.integers(1, 8))
        st[pos:pos + ln] = rng.choice(np.array([0, 1, 3], dtype=np.int8))
        pos += ln + int(rng.integers(event_every // 2, event_every * 3 // 2))
    return st


def make_workload(contig_lens, name="synthetic", window_len=4000, chunk_len=20_000_000, avg_alignment_len=15000,
                  region_coverages=(40,), region_fraction=0.0, seed=0, contig_prefix="ctg"):
    rng = np.random.default_rng(seed)
    region_coverages = np.asarray(region_coverages, dtype=np.int32)
    R = len(region_coverages)
    layout = chunk_layout(contig_lens, chunk_len)
    chunks = np.zeros(len(layout), dtype=_abi.chunk_desc_dtype)
    names, off = [], 0
    for k, (ci, s, e) in enumerate(layout):
        n = -(-(e - s + 1) // window_len)  # last window may be short (chunk.c:393-441)
        chunks[k] = (int(contig_lens[ci]), s, e, window_len, n, 0, off)
        names.append(f"{contig_prefix}{ci + 1}")
        off += n
    W = off
    states = _state_path(rng, W)
    region = np.zeros(W, dtype=np.uint8)
    if R > 1 and region_fraction > 0:
        pos = 0
        while pos < W:
            run = int(rng.integers(20, 401))
            if rng.random() < region_fraction:
                region[pos:pos + run] = rng.integers(1, R)
            pos += run
    base = region_coverages[region].astype(np.float64)
    mu = np.array([0.05, 0.5, 1.0, 2.0])[states] * base
    covf = rng.normal(mu, np.sqrt(1.5 * np.maximum(mu, 0.5)))
    cov = np.clip(np.rint(covf), 0, 250).astype(np.uint16)
    mapq = cov.copy()
    dup = states == 1
    mapq[dup] = np.rint(0.1 * cov[dup]).astype(np.uint16)
    clip = np.zeros(W, dtype=np.uint16)
    ann = ["no_annotation", "whole_genome"] if R == 1 else list(ANNOTATION_NAMES)
    return Workload(name, window_len, chunk_len, avg_alignment_len, region_coverages, names, chunks, cov, mapq, clip,
                    region, states.astype(np.int8), ann)


def config1(seed=1):
    """BASELINE.json configs[0]: single 1 Mbp contig, w=4000 -> 250 windows."""
    return make_workload([1_000_000], name="cfg1_1Mbp_250w", seed=seed)


def genome_contig_lens(total_bp=3_000_000_000):
    hap = np.array(_CHR_MBP, dtype=np.float64) * 1e6
    lens = np.concatenate([hap, hap * 0.985])  # two haplotypes, slightly different lengths
    lens = np.floor(lens * (total_bp / lens.sum())).astype(np.int64)
    return lens


def config2(total_bp=3_000_000_000, seed=2):
    """configs[1]: 3 Gbp diploid, 46 chr-sized contigs, 40x, w=4000 (the metric's workload)."""
    return make_workload(genome_contig_lens(total_bp), name=f"cfg2_{total_bp / 1e9:g}Gbp_46ctg_w4000", seed=seed,
                         contig_prefix="chr")


def config3(n_contigs=10_000, contig_len=300_000, seed=3):
    """configs[2]: 3 Gbp fragmented into 10 000 short contigs (75 windows each)."""
    return make_workload([contig_len] * n_contigs, name=f"cfg3_{n_contigs}x{contig_len}", seed=seed)


def config4(total_bp=3_000_000_000, seed=4):
    """configs[3]: 3 Gbp + per-region emission parameters (R=7, bias-detection output)."""
    return make_workload(genome_contig_lens(total_bp), name=f"cfg4_{total_bp / 1e9:g}Gbp_R7", seed=seed,
                         region_coverages=(40, 52, 30, 61, 25, 48, 36), region_fraction=0.3, contig_prefix="chr")


def small_mixed(seed=5, n_regions=3):
    """Test-size workload with every edge: 1/2/3-window contigs, ragged chunks, multi-chunk contig, regions."""
    lens = [4000, 7000, 9000, 61_000, 300_000, 1_250_000, 2_600_000, 123_457]
    cov = (40, 52, 30, 61, 25, 48, 36)[:n_regions]
    return make_workload(lens, name="small_mixed", chunk_len=1_000_000, region_coverages=cov,
                         region_fraction=0.4 if n_regions > 1 else 0.0, seed=seed)
