"""bench.py contract guards that need no GPU: the committed bench lines under profiles/ carry every key the driver reads, their
roofline / e2e objects are self-consistent, and the reference arm does no work on ranks other than 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = load("r01_bench_coarse_fine.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 4 * 1000.0 / d["ms_per_step"]) <= 1e-6 * d["value"]          # 4 clips per step on one GPU
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9 and 0 < r["frac"] <= 1.0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                                           "sw_thermal_slowdown"}
    b = d["cpu_baseline"]
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_committed_reference_and_scaling_lines():
    ref = load("r01_bench_reference.json")
    ours = load("r01_bench_coarse_fine.json")
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"]["workload"] == ours["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    two = load("r01_bench_2gpu.json")
    assert two["n_gpus"] == 2 and two["scaling"] == "weak" and two["value"] > 1.8 * ours["value"] * 0.9


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun rank 0 alone runs the CPU reference; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_does_not_import_the_product():
    """`bench.py --impl reference` must run the reference (baseline/_ref, else the oracle port) on the host cores without
    importing the product package or loading libcfnet_b200.so (the gridpool workload is the quick one)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "gridpool", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["product_imported"] is False and d["value"] > 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    ref_fn = src[src.index("def run_reference"):src.index("def main")]
    assert "import coarse_fine_networks_b200" not in ref_fn and "from coarse_fine_networks_b200" not in ref_fn
