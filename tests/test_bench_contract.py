"""bench.py's CPU reference arm prints one JSON line with the contract's keys (runs on CPU in a few seconds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "8", "--config", "c2"],
                         capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-500:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "mpc_solves_per_sec" and line["unit"] == "solves/s"
    from oracle import ref_runner
    want = "reference" if ref_runner.available() else "port"       # oracle/_ref: the unmodified reference, byte-compiled
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]
    # `config` carries the same keys in both arms (the driver compares them)
    assert set(line["config"]) == {"workload", "batch_per_gpu", "control_steps", "opt_tolerance", "cold_start",
                                   "footprint_mode", "costmap_mode", "l2", "parallelism"}


def test_reference_arm_port_flag():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--port", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "8", "--config", "c2"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-500:]
    assert json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_committed_bench_line_has_the_contract_keys():
    """profiles/bench_r1_c3_1gpu.json is a line bench.py printed on the B200 box: it must carry every key of the
    measurement contract (task statement section 4 and the base bench contract)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "profiles", "bench_r1_c3_1gpu.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "mpc_solves_per_sec" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "solves/s" and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "solves/s" and e["h2d_bytes_per_step"] == 64 * d["config"]["batch_per_gpu"]
    assert e["d2h_bytes_per_step"] == 32 * d["config"]["batch_per_gpu"]
    assert e["value"] < d["value"]                       # copies are inside the e2e region
    assert d["gpu_launches"] == d["steps"]               # one solve kernel per timed step
    k = d["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["cost_residual"]["median"] <= 0.0 and d["cost_residual"]["max"] <= 5e-3


def test_committed_round2_bench_line():
    """profiles/bench_r2_c3_1gpu.json is the line bench.py printed on the B200 box in round 2: the contract keys, the e2e
    through the twists entry (12 B per problem back), the reference arm on the unmodified reference, the sustained record."""
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_c3_1gpu.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
              "sustained", "cost_residual", "plugin_tick_latency"):
        assert k in d, k
    assert d["metric"] == "mpc_solves_per_sec" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload", "batch_per_gpu", "control_steps", "opt_tolerance", "cold_start",
                                "footprint_mode", "costmap_mode", "l2", "parallelism"} and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and (r["traffic"] is None or r["traffic"] > 0)
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["oracle_port_value"] > c["value"]
    e = d["e2e"]
    n = d["config"]["batch_per_gpu"]
    assert e["h2d_bytes_per_step"] == 64 * n and e["d2h_bytes_per_step"] == 12 * n and 0.85 * d["value"] < e["value"] < d["value"]
    assert d["gpu_launches"] == d["steps"]
    s = d["sustained"]
    assert s["seconds"] >= 1.9 and not s["reasons"] and s["sm_mhz_median"] >= 0.95 * d["clocks"]["sm_max_mhz"]
    cr = d["cost_residual"]
    assert cr["max"] <= 0.0 and cr["p99"] <= 0.0 and cr["frac_worse_than_opt_tol"] == 0.0
    assert cr["first_control_vs_tight_scipy_p90"] <= 1e-2 and cr["first_control_vs_tight_scipy_p99"] <= 3e-2
