"""CPU: the parts of bench.py that run without a GPU -- the reference arm (`--impl reference`: the reference's own CPU
implementation through oracle/_ref, or the restatement when that was not built) prints ONE JSON line with the contract's
keys; `ours` refuses to run without a device (no CPU fallback); the roofline peak reader copes with the driver's file."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=e)


def test_reference_arm_prints_one_contract_line():
    r = _bench("--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in j, key
    assert j["impl"] == "reference" and j["unit"] == "windows/s" and j["higher_is_better"] is True and j["dtype"] == "f64"
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["vs_baseline"] is None and "workload" in j["config"]
    # the reference arm runs the checker only: the product library is not even mapped
    assert j["native_libs"] and not any("libhfg" in lib for lib in j["native_libs"]), j["native_libs"]
    assert set(j["config"]) == {"workload", "windows", "chunks", "regions", "col_components", "window_len", "alpha", "model_type", "step"}


def test_reference_arm_non_zero_ranks_stay_silent():
    r = _bench("--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ours_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _bench("--workload", "small", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_measured_peak_reader(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_peak()[0] == 6650.0
    for content, want in (({"hbm_gbs": 6540.8, "bf16_tflops": 1500}, 6540.8),
                          ({"hbm": {"burst_gbs": 7000, "sustained_gbs": 6500}}, 6500.0), ({"HBM_TBps": 6.54}, 6540.0)):
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(content))
        got, src = bench.measured_peak()
        assert abs(got - want) < 1e-9 and src.startswith("measured")
