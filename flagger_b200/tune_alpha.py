#!/usr/bin/env python
"""tune_alpha.py -- tuning the alpha (previous-window dependency) matrix with the inputs resident on the GPU.

Mirrors the reference driver programs/src/tune_alpha_hmm_flagger.py: the same ten free entries of the 4x4 matrix
(:40-66), the same objective -- per input the mean of (overlap-based F1, base-level F1, contiguity) of the
HARMONIC_MEAN_NO_HAP / auN rows for one annotation and one size bin (:82-111), averaged over the training inputs
(:170-205) -- and the same bounds / start-point options.  What changes is the cost of one evaluation: the reference starts
one `hmm_flagger` process per input and candidate (parse the coverage file, run EM on the host, write and re-read the
benchmarking tables); here every input is parsed and uploaded ONCE, a candidate is one device-resident EM run
(hfg_run_em, ~10 ms for a 3 Gbp assembly) and its labels are scored in memory (hfg_benchmark_scores).  With --lanes N
(default 8) N candidates are fitted at a time on N sub-grids of the GPU (hfg_batch_run_em): the start points as one batch,
then rounds of N proposals around the best point -- 3x the candidates per second on a 300 Mbp input (tools/batch_bench.py).

The reference optimises with smt's EGO (Gaussian-process Bayesian optimisation).  That package is not part of this
repository's environment, so the search here is a seeded derivative-free one: the start points, then proposals drawn
around the best point with a shrinking radius, every fifth one uniform in the box.  With evaluations this cheap the number
of points, not the sample efficiency, is the budget to spend.

    python -m flagger_b200.tune_alpha --inputFilesTrain a.cov.gz,b.cov.gz --outputDir tune_alpha --iterations 500
"""
import argparse
import os
import sys
import time

import numpy as np

from . import _abi

DIMENSION = 10
# the free entries, in the order of convertAlphaMatrixToX (tune_alpha_hmm_flagger.py:40-52)
_FREE = ((0, 0), (0, 2), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2), (2, 3), (3, 2), (3, 3))


def x_to_alpha(x):
    a = np.zeros((4, 4))
    for v, (i, j) in zip(x, _FREE):
        a[i, j] = v
    return a


def alpha_to_x(alpha):
    alpha = np.asarray(alpha, float)
    return np.array([alpha[i, j] for i, j in _FREE])


class GpuEngine:
    """One input resident on one GPU: alpha matrix -> final labels of an EM run (hfg_run_em, device-resident loop)."""

    def __init__(self, cov, model_type="trunc_exp_gaussian", em_iterations=100, convergence_tol=0.001, device=0,
                 collapsed_comps=-1, lanes=1, **config):
        from . import api
        wl = cov.workload
        K = collapsed_comps if collapsed_comps > 0 else api.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
        # (negative_binomial has no alpha dependence: nothing to tune)
        mt = {"trunc_exp_gaussian": _abi.MODEL_TRUNC_EXP_GAUSSIAN, "gaussian": _abi.MODEL_GAUSSIAN}[model_type]
        self.cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, model_type=mt, mean_read_length=wl.avg_alignment_len,
                                    device=device, **config)
        self.params0 = api.model_init(self.cfg, wl.region_coverages, wl.window_len)
        self.gpu = api.HmmFlaggerGPU(self.cfg, wl)
        self.batch = api.HmmFlaggerBatch(self.cfg, wl, n_lanes=lanes) if lanes > 1 else None
        self.em_iterations, self.tol = int(em_iterations), float(convergence_tol)

    def __call__(self, alpha):
        _, _, labels = self.gpu.run_em(np.ascontiguousarray(alpha, np.float64), self.params0, self.em_iterations, tol=self.tol)
        return labels

    def many(self, alphas):
        """Labels of one EM run per alpha matrix, `lanes` runs at a time (hfg_batch_run_em)."""
        if self.batch is None or len(alphas) < 2:
            return [self(a) for a in alphas]
        _, _, labels = self.batch.run_em(np.asarray(alphas, np.float64), self.params0, self.em_iterations, tol=self.tol)
        return list(labels)

    def close(self):
        self.gpu.close()
        if self.batch is not None:
            self.batch.close()


class Objective:
    """score(x) = mean over the training inputs of (overlap F1 + base-level F1 + contiguity) / 3, as
    functionToMinimizeInternal (tune_alpha_hmm_flagger.py:118-222) computes it from the benchmarking files."""

    def __init__(self, train, validation=(), annotation_label="whole_genome", size_label="ALL_SIZES",
                 overlap_ratio_threshold=0.4, bin_array_file=None):
        self.train, self.validation = list(train), list(validation)  # lists of (NativeCov, engine)
        self.kw = dict(annotation_label=annotation_label, size_label=size_label, overlap_ratio_threshold=overlap_ratio_threshold,
                       bin_array_file=bin_array_file)
        self.history = []  # (kind, x, train score, validation score, per-input triples)

    def _scores(self, pairs, alpha):
        out = []
        for cov, engine in pairs:
            s = cov.benchmark_scores(engine(alpha), **self.kw)
            out.append(tuple(0.0 if v != v else v for v in s))  # "NA" rows count as 0
        return out

    def score(self, x, kind="iteration"):
        alpha = x_to_alpha(x)
        tr = self._scores(self.train, alpha)
        va = self._scores(self.validation, alpha)
        train = float(np.mean([sum(t) / 3.0 for t in tr]))
        valid = float(np.mean([sum(t) / 3.0 for t in va])) if va else None
        self.history.append((kind, np.array(x, float), train, valid, tr))
        return train

    def _scores_many(self, pairs, alphas):
        per_input = []
        for cov, engine in pairs:
            labels = engine.many(alphas) if hasattr(engine, "many") else [engine(a) for a in alphas]
            per_input.append([tuple(0.0 if v != v else v for v in cov.benchmark_scores(lab, **self.kw)) for lab in labels])
        return [[per_input[i][c] for i in range(len(pairs))] for c in range(len(alphas))]  # [candidate][input]

    def score_many(self, xs, kind="iteration"):
        """score() of several points, their EM runs batched per input; the history gets one entry per point, in order."""
        alphas = [x_to_alpha(x) for x in xs]
        tr_all = self._scores_many(self.train, alphas)
        va_all = self._scores_many(self.validation, alphas) if self.validation else [[] for _ in xs]
        out = []
        for x, tr, va in zip(xs, tr_all, va_all):
            train = float(np.mean([sum(t) / 3.0 for t in tr]))
            valid = float(np.mean([sum(t) / 3.0 for t in va])) if va else None
            self.history.append((kind, np.array(x, float), train, valid, tr))
            out.append(train)
        return out


def start_points(lower, upper, n, candidate_alpha=None, rng=None):
    """getStartPoints (tune_alpha_hmm_flagger.py:24-34): the candidate matrix (or a random point) first, then random ones."""
    rng = rng or np.random.default_rng(42)
    pts = [alpha_to_x(candidate_alpha) if candidate_alpha is not None else rng.uniform(lower, upper, DIMENSION)]
    pts += [rng.uniform(lower, upper, DIMENSION) for _ in range(max(n, 1) - 1)]
    return np.array(pts)


def optimise(objective, lower=0.0, upper=0.8, n_start=10, n_iter=50, candidate_alpha=None, seed=42, log=None, batch=1):
    """Maximises objective.score over the box [lower, upper]^10.  Returns (best x, best score).  batch > 1: the start points
    are scored as one batch and the proposals in rounds of `batch`, all of a round drawn around the same best point."""
    rng = np.random.default_rng(seed)
    best_x, best = None, -np.inf
    starts = start_points(lower, upper, n_start, candidate_alpha, rng)
    scores = objective.score_many(list(starts), "start") if batch > 1 else [objective.score(x, "start") for x in starts]
    for x, s in zip(starts, scores):
        if s > best:
            best_x, best = x.copy(), s
    radius0 = 0.25 * (upper - lower)
    it = 0
    while it < n_iter:
        xs = []
        for j in range(it, min(it + max(batch, 1), n_iter)):
            if j % 5 == 4:
                xs.append(rng.uniform(lower, upper, DIMENSION))
            else:
                radius = radius0 * (0.05 ** (j / max(n_iter - 1, 1)))  # shrinks to 5 % of the initial radius
                xs.append(np.clip(best_x + rng.normal(0.0, radius, DIMENSION) * (rng.random(DIMENSION) < 0.5), lower, upper))
        scores = objective.score_many(xs) if batch > 1 else [objective.score(xs[0])]
        for x, s in zip(xs, scores):
            it += 1
            if s > best:
                best_x, best = x.copy(), s
            if log:
                log(f"iteration {it}/{n_iter}: score {s:.3f}, best {best:.3f}")
    return best_x, best


def main(argv=None):
    ap = argparse.ArgumentParser(description="Tune the alpha matrix of hmm_flagger on coverage/bin files with truth labels; the "
                                             "inputs stay resident on the GPU and a candidate costs one device-resident EM run.")
    ap.add_argument("--inputFilesTrain", required=True, help="comma-separated .cov/.cov.gz/.bin files with truth labels")
    ap.add_argument("--inputFilesValidation", default="", help="(optional) files scored for every candidate but not optimised on")
    ap.add_argument("--outputDir", default="tune_alpha")
    ap.add_argument("--numberOfStartPoints", type=int, default=10)
    ap.add_argument("--lowerBound", type=float, default=0.0)
    ap.add_argument("--upperBound", type=float, default=0.8)
    ap.add_argument("--iterations", type=int, default=50, help="proposals after the start points")
    ap.add_argument("--modelType", default="gaussian", choices=["gaussian", "trunc_exp_gaussian"])
    ap.add_argument("--annotationLabel", default="whole_genome")
    ap.add_argument("--sizeLabel", default="ALL_SIZES")
    ap.add_argument("--binArrayFile", default="")
    ap.add_argument("--candidateAlphaTsv", default="")
    ap.add_argument("--emIterations", type=int, default=100, help="hmm_flagger --iterations of every run")
    ap.add_argument("--convergenceTol", type=float, default=0.001)
    ap.add_argument("--chunkLen", type=int, default=20_000_000)
    ap.add_argument("--windowLen", type=int, default=4000)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=8, help="candidates fitted at a time on sub-grids of the GPU (1 = one after the other)")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args(argv)
    from . import binfmt

    def load(paths):
        pairs = []
        for p in [q for q in paths.strip().split(",") if q]:
            cov = binfmt.NativeCov(p, args.chunkLen, args.windowLen)
            pairs.append((cov, GpuEngine(cov, args.modelType, args.emIterations, args.convergenceTol, args.device, lanes=args.lanes)))
        return pairs

    os.makedirs(args.outputDir, exist_ok=True)
    t0 = time.time()
    obj = Objective(load(args.inputFilesTrain), load(args.inputFilesValidation), args.annotationLabel, args.sizeLabel,
                    bin_array_file=args.binArrayFile or None)
    cand = np.loadtxt(args.candidateAlphaTsv) if args.candidateAlphaTsv else None
    log = lambda m: print(f"[tune_alpha] {m}", file=sys.stderr, flush=True)
    log(f"{len(obj.train)} training / {len(obj.validation)} validation inputs resident after {time.time() - t0:.1f} s")
    best_x, best = optimise(obj, args.lowerBound, args.upperBound, args.numberOfStartPoints, args.iterations, cand, args.seed, log,
                            batch=args.lanes)
    np.savetxt(os.path.join(args.outputDir, "alpha_optimum.tsv"), x_to_alpha(best_x), delimiter="\t", fmt="%.3f")
    with open(os.path.join(args.outputDir, "scores.tsv"), "w") as f:
        f.write("#point\tkind\ttrain_score\tvalidation_score\t" + "\t".join(f"x{i}" for i in range(DIMENSION)) + "\n")
        for i, (kind, x, tr, va, _) in enumerate(obj.history):
            f.write(f"{i + 1}\t{kind}\t{tr:.3f}\t{'NA' if va is None else f'{va:.3f}'}\t" + "\t".join(f"{v:.3f}" for v in x) + "\n")
    n = len(obj.history)
    log(f"best train score {best:.3f} after {n} evaluations in {time.time() - t0:.1f} s ({(time.time() - t0) / max(n, 1) * 1e3:.0f} ms each)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
