"""bench.py's one-JSON-line contract: the CPU reference arm here, the GPU arm on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cg-its", "100", "--cpu-grid", "40"], 600)
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "newton_step_dof_per_s" and d["unit"] == "DOF/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    ms = d["cpu_baseline"]["measured_step"]                 # a genuinely timed full Newton step, not a model
    assert ms["measured_grid"] == [40, 40, 40] and ms["measured_cg_iterations"] > 10 and ms["measured_step_s"] > 0
    assert 0.3 < d["cpu_baseline"]["model_over_measured"] < 3.0


def test_both_arms_describe_the_same_config():
    """`config` must be the same object in both arms (the driver compares them)."""
    sys.path.insert(0, ROOT)
    import bench
    a = bench.bench_config(256, 256, 256, 1, False)
    assert set(a) == {"workload", "grid", "ndof", "matrix", "parallelism", "ksp", "l2"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(') == 2          # both arms build it with the same call


def test_bench_does_not_write_into_the_repo():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "json.dump(" not in src and "open(KNOWN" not in src


@pytest.mark.gpu
def test_gpu_arm_line_small_grid():
    d = run_bench(["--grid", "48", "--steps", "2", "--warmup", "3", "--no-cpu"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "matrix_free"} <= set(d)
    assert d["metric"] == "newton_step_dof_per_s" and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 3 * 48 ** 3 and d["e2e"]["d2h_bytes_per_step"] >= 8 * 3 * 48 ** 3
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["achieved"] > 0 and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["newton_its_per_step"] == [1, 1]
    assert d["matrix_free"]["cg_iterations"] == pytest.approx(d["cg_iterations_per_step"], abs=2)
    other = d["assembled_full"] if "symmetric" in d["operator"] else d["assembled_sym"]
    assert other["cg_iterations"] == pytest.approx(d["cg_iterations_per_step"], abs=2)
    assert d["cg_iteration_ms"] > 0 and d["fp64"]["dfma_tflops_measured"] > 1.0
    assert "jacobian_per_element_per_gp" in d["kernels_ms"]
