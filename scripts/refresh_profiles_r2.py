#!/usr/bin/env python3
"""After scripts/gpu_evidence_r2.sh ran on the GPU box: turn what came back in gpurun_out/ into the tracked files under
profiles/ (ncu summaries per config, traffic.json, bench lines, launch list, logs).  Run here, needs `ncu` for reading only.

    python scripts/refresh_profiles_r2.py
"""
import csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return {h: v for h, v in zip(rows[0], rows[2])}


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise ValueError(path)


traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path))
for cfg in ("c3", "c2", "c4", "c5"):
    rep = os.path.join(O, f"prof_r2_{cfg}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep); continue
    md = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    md = md.replace(os.path.join(O, ""), "gpurun_out/").replace("(config C3: 65536 problems, control_steps 10)", f"--config {cfg}")
    open(os.path.join(P, f"solve_kernel_r2_{cfg}_summary.md"), "w").write(md)
    m = raw_metrics(rep)
    line = last_json(os.path.join(O, f"r2_bench_{cfg}.json"))
    f = lambda k: float(m[k].replace(",", ""))
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    units = dict(zip(rows[0], rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = rd * scale.get(units["dram__bytes_read.sum"], 1) + wr * scale.get(units["dram__bytes_write.sum"], 1)
    traffic[cfg] = {
        "kernel": line["roofline"]["kernel"],
        "kernel_captured": m.get("Kernel Name", "")[:40],
        "batch": line["config"]["batch_per_gpu"],
        "dram_bytes_per_launch": int(round(dram)),
        "issue_active_pct": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "inst_executed": int(f("smsp__inst_executed.sum")),
        "registers": int(f("launch__registers_per_thread")),
        "ncu_time_us": round(f("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units["gpu__time_duration.sum"], 1), 1),
        "source": f"profiles/solve_kernel_r2_{cfg}_summary.md: one ncu --set full capture around bench.py --config {cfg} (round 2, final build)",
    }
json.dump(traffic, open(traffic_path, "w"), indent=1)
copies = {"r2_bench_c3.json": "bench_r2_c3_1gpu.json", "r2_bench_c2.json": "bench_r2_c2_1gpu.json", "r2_bench_c4.json": "bench_r2_c4_1gpu.json",
          "r2_bench_c5.json": "bench_r2_c5_1gpu.json", "r2_bench_c3_unguided.json": "bench_r2_c3_unguided_1gpu.json",
          "r2_bench_ref_c3.json": "bench_r2_c3_reference_arm.json", "launches_r2.csv": "launches_r2.csv",
          "plugin_latency_r2.txt": "plugin_latency_r2.txt", "sanitizer_r2.txt": "sanitizer_r2.txt", "r2_gpu_tests.log": "gpu_tests_r2.log",
          "r2_smoke.log": "smoke_r2.log"}
for src, dst in copies.items():
    s = os.path.join(O, src)
    if not os.path.exists(s):
        print("missing", s); continue
    if src.endswith(".json"):
        open(os.path.join(P, dst), "w").write(json.dumps(last_json(s)) + "\n")
    else:
        shutil.copy(s, os.path.join(P, dst))
print("profiles refreshed")
