#!/usr/bin/env python
"""bench.py -- HMM windows/sec for EM iterations of the HMM-Flagger E-step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|...]

A *step* is one EM iteration over the whole synthetic workload: the E-step of every chunk (forward, backward, pair
statistics, posterior-argmax labels) plus the M-step -- what runHMMFlagger does once per iteration (reference
src/hmm_flagger.c:337-431).  The default K=51 is the job BASELINE.json names: 50 EM iterations + the
final decode pass, on configs[1] (3 Gbp diploid, 46 chr-sized contigs, 40x, w=4000; ~750k windows).

Prints ONE JSON line (rank 0).  `value` = windows x steps / time with the windows resident in HBM (device-resident EM
loop: one kernel per iteration, M-step in its tail, each iteration timed with its own CUDA-event pair); `e2e` = the same
through the blocking C-ABI call a host program makes (hfg_set_chunks once + hfg_em_iteration per step: host parameters
in, host statistics + labels out, + host M-step); `e2e_job` = the whole job as one hfg_run_em call, set-up included.  `--impl reference` times the reference's own CPU implementation
(oracle/_ref, all host threads) on the same workload/metric.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_WINDOW = 87  # 3 B obs x 2 sweeps + 40 B f^/scale written + 40 B read + 1 B label


_RESULT_LINE = []


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=51)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "small"])
    ap.add_argument("--total-bp", type=float, default=3e9)
    ap.add_argument("--model", default="trunc_exp_gaussian", choices=["trunc_exp_gaussian", "nb"],
                    help="emission model: the reference's default, or --modelType negative_binomial")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the 3 Gbp workload itself sharded over the ranks (BASELINE.json configs[4]); "
                         "weak = 3 Gbp PER GPU (an N x 3 Gbp assembly, chunks sharded over the ranks, one model).  A strong run "
                         "also reports the weak figure as the extra key `weak`")
    ap.add_argument("--no-binary", action="store_true", help="skip the whole-binary end-to-end leg (e2e_binary)")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1: sum of the EM statistics over ranks inside the E-step kernel through peer memory "
                         "(fused, default) or with one NCCL all-reduce per iteration after it (nccl)")
    return ap.parse_args()


def make_workload(name, total_bp):
    from flagger_b200 import synth
    if name == "cfg1":
        return synth.config1()
    if name == "cfg2":
        return synth.config2(total_bp=int(total_bp))
    if name == "cfg3":
        return synth.config3(n_contigs=int(total_bp // 300_000))
    if name == "cfg4":
        return synth.config4(total_bp=int(total_bp))
    return synth.small_mixed()


def shard_chunks(wl, rank, world):
    """Contiguous chunk ranges in list order, balanced by window count."""
    from flagger_b200 import dist as hdist
    return hdist.shard_chunks(wl, rank, world)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        # default: in-process NVML polling (no nvidia-smi process attaching to the driver and taking its locks every 100 ms
        # next to the job -- that showed up as sporadic 30-50 ms stalls of cudaMallocHost/cudaFree in the end-to-end legs);
        # BENCH_SAMPLER=smi forces the nvidia-smi loop, =none disables sampling
        mode = os.environ.get("BENCH_SAMPLER", "nvml")
        if mode == "none":
            return
        if mode == "nvml":
            try:
                return self._start_nvml()
            except Exception:
                pass  # no pynvml / NVML failure: fall back to nvidia-smi
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi attaches to the driver for tens of ms (and holds its locks): let that finish, first sample in
            # hand, before any timed region starts
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _start_nvml(self):
        """In-process NVML polling (nvidia_ml_py): no nvidia-smi process attaching to the driver next to the job."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        self.nvml_stop = threading.Event()

        def loop():
            while not self.nvml_stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(self.idx), str(sm), str(mx), "", ""] +
                                 ["Active" if r & b else "Not Active" for b in bits.values()])
                self.nvml_stop.wait(0.02)
        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()
        self.proc = "nvml"

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        if self.proc == "nvml":
            self.nvml_stop.set()
            self.t.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    """HBM peak (GB/s) for the roofline: the driver-written MEASURED_PEAKS.json when present (any numeric entry whose key
    mentions hbm; the sustained figure is preferred because the kernel is timed inside a long loop of iterations), else the
    6.65 TB/s fallback of B200_PROFILING.md.  Returns (GB/s, where it came from)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            found = []

            def walk(obj, prefix=""):
                if isinstance(obj, dict):
                    for k, v in obj.items():
                        walk(v, f"{prefix}.{k}" if prefix else str(k))
                elif isinstance(obj, (int, float)) and not isinstance(obj, bool) and "hbm" in prefix.lower():
                    found.append((prefix, float(obj)))

            walk(json.load(open(path)))
            if found:
                rank = lambda kv: (0 if "sustain" in kv[0].lower() else 1 if kv[0].lower().endswith("hbm_gbs") else
                                   2 if "burst" in kv[0].lower() else 3)
                key, val = sorted(found, key=rank)[0]
                if val < 100.0:  # given in TB/s
                    val *= 1000.0
                return val, f"measured (MEASURED_PEAKS.json {key})"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the E-step kernel per launch, from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(workload_name)
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def model_setup(wl, checker=None, model="trunc_exp_gaussian"):
    """Configuration and initial model of a workload.  `checker` (an oracle_lib checker: the reference build or the
    restatement) serves the reference arm, which must not load the product library; the product arm uses libhfg's own
    host mirrors (the two agree bit for bit, tests/test_oracle_golden.py)."""
    from flagger_b200 import _abi, synth
    if checker is None:
        from flagger_b200 import api as checker
    K = checker.best_num_collapsed_comps(int(wl.cov.max()), wl.region_coverages)
    if model == "nb":  # --modelType negative_binomial: no previous-window dependency, alpha is ignored
        cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, mean_read_length=wl.avg_alignment_len,
                               model_type=_abi.MODEL_NEGATIVE_BINOMIAL)
        alpha = np.zeros((4, 4))
    else:
        cfg = _abi.make_config(n_regions=wl.n_regions, n_col_comps=K, mean_read_length=wl.avg_alignment_len)
        alpha = synth.HIFI_ALPHA.copy()
    params = checker.model_init(cfg, wl.region_coverages, wl.window_len)
    return cfg, params, alpha, K


def workload_config(wl, K, model="trunc_exp_gaussian"):
    """The `config` object of the JSON line: the same keys in both arms."""
    return {"workload": wl.name, "windows": wl.n_windows, "chunks": wl.n_chunks, "regions": wl.n_regions, "col_components": K,
            "window_len": wl.window_len, "alpha": "HiFi_DC_1.2" if model != "nb" else "none (negative binomial)",
            "model_type": "negative_binomial" if model == "nb" else "trunc_exp_gaussian",
            "step": "one EM iteration = E-step of all chunks (fwd+bwd+statistics+labels) + M-step"}


def time_cpu(cfg, wl, alpha, params, steps, warmup, budget_s=90.0):
    """Reference CPU implementation (oracle/_ref when built, else the restatement): whole EM iterations with all host
    threads on a bounded sample of the workload (a prefix of its chunks sized to the time budget)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    cores = host_threads()
    # calibrate the rate on a small prefix
    n = wl.chunks["n_windows"].astype(np.int64)
    cum = np.cumsum(n)
    k0 = int(np.searchsorted(cum, min(40_000, int(cum[-1])), side="left")) + 1
    probe = wl.subset(range(min(k0, len(n))))
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)  # the reference prints two warnings per M-step
    try:
        run = oracle_lib.CpuRun(cfg, probe, alpha, params, cores)
        run.step()
        secs, _ = run.step()
        run.close()
        rate = probe.n_windows / max(secs, 1e-6)
        # the largest chunk prefix whose (steps + warmup) iterations fit the budget
        max_windows = int(rate * budget_s / max(steps + warmup, 1))
        ks = int(np.searchsorted(cum, max_windows, side="right"))
        ks = min(max(ks, 1), len(n))
        sample = wl.subset(range(ks)) if ks < len(n) else wl
        run = oracle_lib.CpuRun(cfg, sample, alpha, params, cores)
        for _ in range(warmup):
            run.step()
        total = 0.0
        for _ in range(steps):
            s, _ = run.step()
            total += s
        kind, used = run.kind, run.cores
        run.close()
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)
    value = sample.n_windows * steps / total
    desc = (f"{sample.n_chunks}/{wl.n_chunks} chunks ({sample.n_windows} of {wl.n_windows} windows) of {wl.name}, "
            f"{steps} EM iterations after {warmup} warm-up, E-step (pthread pool, {used} threads) + M-step")
    return {"value": value, "unit": "windows/s", "cores": used, "kind": kind, "sample": desc,
            "ms_per_step": 1e3 * total / steps}, sample


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    wl = make_workload(args.workload, args.total_bp * (max(args.gpus, 1) if args.scaling == "weak" else 1))
    checker = oracle_lib.reference(threads=host_threads()) or oracle_lib.oracle()  # never the product library
    cfg, params, alpha, K = model_setup(wl, checker, args.model)
    cpu, sample = time_cpu(cfg, wl, alpha, params, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "HMM windows/sec (EM iter + Viterbi), 3 Gbp @40x w=4000; achieved HBM GB/s",
        "value": cpu["value"], "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, K, args.model),
        "notes": {"sample_windows": sample.n_windows,
                  "note": "CPU reference path on a bounded sample; windows/s is size-independent (per-window cost is constant)"},
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cpu["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # shared objects of this repository mapped into the process: the checker's only, never the product library
        "native_libs": sorted({os.path.relpath(ln.split()[-1], ROOT) for ln in open("/proc/self/maps")
                               if ln.rstrip().endswith(".so") and ln.split()[-1].startswith(ROOT)}),
    }
    _RESULT_LINE.append(json.dumps(line))


def binary_e2e(wl, alpha, steps, cores, model="trunc_exp_gaussian"):
    """Whole-binary end to end: the stand-alone hmm_flagger_b200 and, beside it, the unmodified reference binary
    (oracle/_ref/hmm_flagger_ref, -@ all host threads) on the same .bin input, `steps - 1` EM iterations + final inference,
    timed from process start to exit.  Returns a dict for the JSON line (or None when the product binary is missing)."""
    import re
    import shutil
    import tempfile
    from flagger_b200 import binfmt
    ours = os.path.join(ROOT, "flagger_b200", "hmm_flagger_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "hmm_flagger_ref")
    if not os.path.exists(ours):
        return None
    tmp = tempfile.mkdtemp(prefix="hfg_bench_")
    try:
        inp, atsv = os.path.join(tmp, "in.bin"), os.path.join(tmp, "alpha.tsv")
        binfmt.write_bin(wl, inp)
        binfmt.write_alpha_tsv(alpha, atsv)
        out = {"input": f"{wl.name} as .bin ({os.path.getsize(inp) >> 20} MiB)",
               "command": f"-i in.bin -o out -A alpha.tsv -n {steps - 1} -t 1e-12 -@ {cores}" +
                          (" -m negative_binomial" if model == "nb" else "")}
        extra = ["--modelType", "negative_binomial"] if model == "nb" else []

        def one(binary, tag, repeats):
            best = None
            for rep in range(repeats):
                odir = os.path.join(tmp, f"out_{tag}_{rep}")
                os.makedirs(odir)
                t0 = time.perf_counter()
                r = subprocess.run([binary, "-i", inp, "-o", odir, "-A", atsv, "-n", str(steps - 1), "-t", "1e-12", "-@", str(cores)] + extra,
                                   capture_output=True, text=True)
                wall = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"error": r.stderr[-300:]}
                m = re.search(r"Real time:\s+([0-9.]+) sec", r.stderr)
                ph = re.search(r"Phases: read ([0-9.]+) s; GPU set-up ([0-9.]+) s; EM ([0-9.]+) s; summary tables ([0-9.]+) s; BED ([0-9.]+) s", r.stderr)
                rec = {"wall_s": wall, "real_time_line_s": float(m.group(1)) if m else None}
                if ph:
                    rec["phases_s"] = dict(zip(("read", "gpu_setup", "em", "summary_tables", "bed"), map(float, ph.groups())))
                rec["bed"] = open(os.path.join(odir, "final_flagger_prediction.bed"), "rb").read()
                if best is None or wall < best["wall_s"]:
                    best = rec
            return best

        o = one(ours, "ours", 3)  # the first process pays CUDA context creation on a cold driver: best of three
        out["hmm_flagger_b200"] = {k: v for k, v in o.items() if k != "bed"}
        if os.path.exists(ref) and "error" not in o:
            rr = one(ref, "ref", 1)
            out["hmm_flagger_ref"] = {k: v for k, v in rr.items() if k != "bed"}
            if "error" not in rr:
                out["bed_identical"] = rr["bed"] == o["bed"]
                out["speedup_wall"] = rr["wall_s"] / o["wall_s"]
        out["windows_per_s"] = wl.n_windows * steps / o["wall_s"] if "error" not in o else None
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from flagger_b200 import _abi, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: the per-GPU work is the metric's 3 Gbp workload; the job is an N x 3 Gbp assembly with ONE model, its
    # chunks sharded over the ranks and its EM statistics summed over them every iteration
    wl_full = make_workload(args.workload, args.total_bp * (world if args.scaling == "weak" else 1))
    cfg, params0, alpha, K = model_setup(wl_full, None, args.model)
    cfg["device"] = local_rank
    wl = shard_chunks(wl_full, rank, world)
    W_total = wl_full.n_windows
    R = wl_full.n_regions

    gpu = api.HmmFlaggerGPU(cfg, wl, timing=True)
    fused = world > 1 and args.allreduce == "fused"
    if fused:
        gpu.peer_connect(dist)  # CUDA-IPC mailboxes: the kernel itself sums the statistics over the ranks
    stats_bytes = gpu.stats_device_bytes()
    n_d = stats_bytes // 8
    stats_dev = torch.zeros(n_d, dtype=torch.float64, device=dev)
    stats_host = torch.zeros(n_d, dtype=torch.float64).pin_memory()
    stream = torch.cuda.Stream(device=dev)  # an explicit stream: its handle is what the C-ABI launches on
    torch.cuda.set_stream(stream)
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    SD = _abi.region_stats_dtype.itemsize // 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(params):
        """E-step on the device -> [all-reduce over ranks] -> statistics to the host -> host M-step."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gpu.em_iteration_device(alpha, params, stats_dev.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(stats_dev, op=dist.ReduceOp.SUM)  # one NCCL all-reduce of the EM statistics per iteration
        stats_host.copy_(stats_dev, non_blocking=True)
        e1.record(stream)
        e1.synchronize()
        t0 = time.perf_counter()
        flat = stats_host.numpy()
        if flat[n_d - 1] != 0:
            raise SystemExit(f"E-step reported error flags {flat[n_d - 1]}")
        stats = flat[: R * SD].view(_abi.region_stats_dtype)
        new_params, _ = api.mstep(cfg, params, stats, tol=1e-12)
        t_host = time.perf_counter() - t0
        return new_params, e0.elapsed_time(e1) * 1e-3 + t_host, float(flat[n_d - 2]), gpu.last_estep_kernel_ms()

    # ---- resident (value) ----
    if world == 1 or fused:
        # device-resident EM loop: the parameters stay in HBM, the M-step runs in the tail of the E-step kernel, and the
        # iterations are back-to-back launches on the library's stream -- the host only queues them.  Each iteration is
        # timed on that stream with its own pair of CUDA events; the L2 flush between iterations is outside the pairs.
        gpu.em_begin(alpha, params0, tol=1e-12, max_esteps=args.warmup + args.steps)
        for _ in range(args.warmup):
            gpu.em_enqueue()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = gpu.kernel_launches()
        barrier()
        for _ in range(args.steps):
            if flush is not None:
                gpu.l2_flush(256 << 20)  # 256 MiB > 126 MB L2: evicts the working set; outside the timed intervals
            if fused:
                gpu.peer_barrier()  # device-side rendezvous: every rank's timed interval starts together
            gpu.em_enqueue()
        barrier()
        params, ll_all, _, _ = gpu.em_finish(want_labels=False)
        if len(ll_all) != args.warmup + args.steps:
            raise SystemExit(f"device EM loop ran {len(ll_all)} E-steps, expected {args.warmup + args.steps}")
        kern_ms = [gpu.em_enqueued_ms(args.warmup + i) for i in range(args.steps)]
        step_s = [1e-3 * m for m in kern_ms]
        logliks = [float(v) for v in ll_all[args.warmup:]]
        timing_note = ("sum over steps of the CUDA-event interval around each iteration's kernel on the launching stream "
                       "(E-step of all chunks + statistics reduction" + (" + all-reduce over ranks" if world > 1 else "") +
                       " + M-step, all on the device; parameters stay in HBM), max over ranks")
    else:
        params = params0.copy()
        for _ in range(args.warmup):
            params, _, _, _ = resident_step(params)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = gpu.kernel_launches()
        barrier()
        step_s, kern_ms, logliks = [], [], []
        for _ in range(args.steps):
            if flush is not None:
                flush.fill_(1)  # 256 MiB > 126 MB L2: evicts the working set; outside the timed interval
                torch.cuda.synchronize()
            params, s, ll, kms = resident_step(params)
            step_s.append(s)
            kern_ms.append(kms)
            logliks.append(ll)
        barrier()
        timing_note = ("sum over steps of [CUDA-event interval on the launching stream (parameter upload, kernel, NCCL "
                       "all-reduce, statistics read-back) + host M-step], max over ranks")
    launches = gpu.kernel_launches() - launches0
    tail_clocks = gpu.debug_phase_clocks()[-1].copy() if (world > 1 and fused) else None  # of the last timed iteration's kernel
    t_total = torch.tensor([sum(step_s)], dtype=torch.float64, device=dev)
    k_total = torch.tensor([sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(k_total, op=dist.ReduceOp.MAX)
    t_total, k_total = float(t_total.item()), float(k_total.item())

    # ---- end to end through the blocking C-ABI call with host buffers ----
    # one "job" = hfg_create + hfg_set_chunks (host windows -> HBM, once) + K x [hfg_em_iteration with host parameters in,
    # host statistics and labels out, + host M-step]; the job is run E2E_REPEATS times and the median job time is reported
    E2E_REPEATS = 5
    labels_pinned = api.PinnedArray(wl.n_windows, np.int8)  # page-locked result buffer (hfg_host_alloc): no staging copy
    labels = labels_pinned.array
    jobs = []
    for rep in range(E2E_REPEATS):
        params = params0.copy()
        stats = np.zeros(R, dtype=_abi.region_stats_dtype)
        barrier()
        t0 = time.perf_counter()
        gpu2 = api.HmmFlaggerGPU(cfg)
        t_create = time.perf_counter() - t0
        gpu2.set_chunks(wl)  # host windows -> packed, segment-transposed words -> HBM (once per job)
        t_set = time.perf_counter() - t0
        if fused:
            gpu2.peer_connect(dist)
        t_upload = time.perf_counter() - t0
        if world > 1 and os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] rank {rank} e2e job {rep}: create {1e3 * t_create:.2f} ms, + set_chunks {1e3 * t_set:.2f} ms, + peer_connect "
                  f"{1e3 * t_upload:.2f} ms", file=sys.stderr)
        if world == 1 or fused:
            # the step loop runs in C, as the reference-side binding runs it (integration/hmm_estep_cuda.c: hfg_em_iteration with
            # host buffers, then the M-step on the host); the L2 flush before every step is outside the step's interval
            if fused:
                gpu2.peer_barrier()
            params, stats, _, labels, secs = gpu2.blocking_steps(alpha, params, args.steps, stats=stats, labels=labels,
                                                                 flush_bytes=(256 << 20) if flush is not None else 0, tol=1e-12)
            e2e_s = [float(v) for v in secs]
        else:
            e2e_s = []
            for i in range(args.steps):
                if flush is not None:
                    flush.fill_(1)
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                stats, ll, labels = gpu2.em_iteration(alpha, params, stats=stats, labels=labels)
                t = torch.from_numpy(_abi.stats_as_flat(stats)).to(dev)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                stats = t.cpu().numpy().view(_abi.region_stats_dtype)
                params, _ = api.mstep(cfg, params, stats, tol=1e-12)
                e2e_s.append(time.perf_counter() - t0)
        if rank == 0:
            print(f"[bench] e2e job {rep}: create {1e3 * t_create:.2f} ms, create+set_chunks {1e3 * t_upload:.2f} ms, steps "
                  f"mean {1e3 * np.mean(e2e_s):.3f} ms, max {1e3 * np.max(e2e_s):.3f} ms", file=sys.stderr)
        job = torch.tensor([sum(e2e_s) + t_upload], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(job, op=dist.ReduceOp.MAX)
        jobs.append(float(job.item()))
        gpu2.close()
    e2e_total = float(np.median(jobs))
    # the whole job as ONE call: hfg_create + hfg_set_chunks + hfg_run_em (host parameters in; parameters, log-likelihoods
    # and labels out; the EM loop itself is device-resident)
    run_jobs = []
    for rep in range(E2E_REPEATS):
        barrier()
        t0 = time.perf_counter()
        gpu3 = api.HmmFlaggerGPU(cfg)
        gpu3.set_chunks(wl)
        if fused:
            gpu3.peer_connect(dist)
        gpu3.run_em(alpha, params0, args.steps - 1, tol=1e-12)
        job = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(job, op=dist.ReduceOp.MAX)
        run_jobs.append(float(job.item()))
        if rank == 0:
            print(f"[bench] hfg_run_em job {rep}: {1e3 * run_jobs[-1]:.2f} ms", file=sys.stderr)
        gpu3.close()
    run_job_total = float(np.median(run_jobs))
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- multi-GPU: parity of the sharded run, cost of the exchange, the weak-scaling figure -------------------------------
    parity = exchange = weak = None
    if world > 1:
        # one E-step with the initial parameters: (i) the in-kernel all-reduce against per-rank sums added by NCCL,
        # (ii) every rank's labels against a single-GPU run of the whole workload on rank 0
        st_f, ll_f, lab = gpu.em_iteration(alpha, params0)
        g_local = api.HmmFlaggerGPU(cfg, wl)  # same shard, not connected to the peers
        st_l, ll_l, _ = g_local.em_iteration(alpha, params0)
        g_local.close()
        t = torch.from_numpy(np.concatenate([_abi.stats_as_flat(st_l).ravel(), [ll_l]])).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nccl_sum = t.cpu().numpy()
        mine = np.concatenate([_abi.stats_as_flat(st_f).ravel(), [ll_f]]) if fused else nccl_sum
        rel = float(np.max(np.abs(mine - nccl_sum)) / np.max(np.abs(nccl_sum)))
        gathered = [None] * world
        dist.all_gather_object(gathered, lab.tobytes())
        if rank == 0:
            g_full = api.HmmFlaggerGPU(cfg, wl_full)
            st_1, ll_1, lab_1 = g_full.em_iteration(alpha, params0)
            g_full.close()
            lab_all = np.frombuffer(b"".join(gathered), np.int8)
            one = np.concatenate([_abi.stats_as_flat(st_1).ravel(), [ll_1]])
            parity = {"fused_vs_nccl_max_rel": rel, "label_mismatches_vs_single_gpu": int((lab_all != lab_1).sum()),
                      "stats_vs_single_gpu_max_rel": float(np.max(np.abs(mine - one)) / np.max(np.abs(one))),
                      "windows": int(lab_1.size), "ok": bool(rel <= 1e-12 and np.array_equal(lab_all, lab_1))}
        if fused:
            # the in-kernel exchange of the last timed iteration, from the kernel's own clocks (tail row: 3 = first store into
            # the peers' mailboxes, 4 = all N vectors summed): what the all-reduce costs this rank, waiting included
            tail = tail_clocks
            ex = torch.tensor([float(tail[4] - tail[3]) / 1.965e3], dtype=torch.float64, device=dev)  # microseconds at 1965 MHz
            exs = [torch.zeros_like(ex) for _ in range(world)]
            dist.all_gather(exs, ex)
            exchange = {"per_rank_us": [round(float(e.item()), 2) for e in exs], "note": "in-kernel all-reduce of the last "
                        "iteration, kernel clocks at 1965 MHz: stores to the peers + wait for their vectors + sum"}
        if args.scaling == "strong":
            # the same loop on an N x 3 Gbp assembly (3 Gbp per GPU): the weak-scaling figure, 20 timed iterations
            wl_w = shard_chunks(make_workload(args.workload, args.total_bp * world), rank, world)
            gw = api.HmmFlaggerGPU(cfg, wl_w)
            if fused:
                gw.peer_connect(dist)
            ws, ww = min(20, args.steps), 3
            gw.em_begin(alpha, params0, tol=1e-12, max_esteps=ws + ww)
            for i in range(ws + ww):
                if flush is not None:
                    gw.l2_flush(256 << 20)
                if fused:
                    gw.peer_barrier()
                gw.em_enqueue()
            barrier()
            gw.em_finish(want_labels=False)
            tw = torch.tensor([sum(gw.em_enqueued_ms(ww + i) for i in range(ws)) * 1e-3], dtype=torch.float64, device=dev)
            nw_ = torch.tensor([float(wl_w.n_windows)], dtype=torch.float64, device=dev)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            dist.all_reduce(nw_, op=dist.ReduceOp.SUM)
            weak = {"value": float(nw_.item()) * ws / float(tw.item()), "unit": "windows/s", "windows": int(nw_.item()), "steps": ws,
                    "ms_per_step": 1e3 * float(tw.item()) / ws, "workload": f"{world} x {args.total_bp / 1e9:g} Gbp, one model"}
            gw.close()
        barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        value = W_total * args.steps / t_total
        kernel_s = k_total / args.steps * 1e-3
        achieved = ALGO_BYTES_PER_WINDOW * wl.n_windows / kernel_s / 1e9  # this rank's kernel over its own windows
        obs_bytes = wl.n_windows * 7  # u16 x3 + u8 region per window handed to hfg_set_chunks
        line = {
            "metric": "HMM windows/sec (EM iter + Viterbi), 3 Gbp @40x w=4000; achieved HBM GB/s",
            "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl_full, K, args.model),
            "notes": {"parallelism": ("1 GPU" if world == 1 else
                                      f"chunks sharded over {world} GPUs, one process per GPU; EM statistics summed over "
                                      "ranks " + ("inside the E-step kernel through NVLink peer memory (fused all-reduce)"
                                                  if fused else "with one NCCL all-reduce per iteration")),
                      "l2": "not flushed" if args.no_flush else "flushed between timed steps (256 MiB fill, untimed)",
                      "timing": timing_note},
            "e2e": {"value": W_total * args.steps / e2e_total, "unit": "windows/s",
                    "h2d_bytes_per_step": int(params.nbytes + obs_bytes / args.steps),
                    "d2h_bytes_per_step": int(stats.nbytes + 16 + wl.n_windows),
                    "includes": "hfg_create + hfg_set_chunks once, then per step hfg_em_iteration (host params in, host "
                                "statistics + labels out, labels into a page-locked buffer) + host hfg_mstep, the step loop driven "
                                "from C as the drop-in binding drives it (hfg_debug_blocking_steps); a single-region model takes the "
                                "call's one-launch path (parameters in the kernel arguments, completion word polled in pinned memory); "
                                "median of 5 such jobs",
                    "job_ms": [1e3 * j for j in jobs]},
            "e2e_job": {"value": W_total * args.steps / run_job_total, "unit": "windows/s",
                        "call": f"hfg_create + hfg_set_chunks + hfg_run_em({args.steps - 1} EM iterations + final inference) with "
                                "host buffers: windows and parameters in; parameters, log-likelihoods and labels out",
                        "job_ms": [1e3 * j for j in run_jobs]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(wl_full.name) if world == 1 else None, "kernel": "hfg_estep_v3_kernel", "kernel_ms": kernel_s * 1e3,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_WINDOW * wl.n_windows, "peak_source": peak_src},
            "clocks": clocks,
            "loglik_first_last": [logliks[0], logliks[-1]],
        }
        if parity is not None:
            line["parity_check"] = parity
        if exchange is not None:
            line["exchange"] = exchange
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_binary and args.workload != "small":
            line["e2e_binary"] = binary_e2e(wl_full, alpha, args.steps, host_threads(), args.model)
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = time_cpu(cfg, wl_full, alpha, params0, steps=5, warmup=1, budget_s=30.0)
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _RESULT_LINE.append(json.dumps(line))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that print banners there (NCCL_DEBUG=VERSION ...) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
