"""Summarise gpurun_out/prof_<kernel>.ncu-rep captures into profiles/ (tracked).

    python scripts/ncu_summary.py r01 k_step_a k_step_b ...

Writes profiles/<round>_<kernel>.txt (headline metrics + hottest source lines) and
profiles/traffic.json (dram bytes per launch, read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd, kernels = sys.argv[1], sys.argv[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tp = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tp)) if os.path.exists(tp) else {}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
for k in kernels:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{k}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none, {len(body)} launch(es) of {body[0][col['Kernel Name']]}",
             f"# source: gpurun_out/prof_{k}.ncu-rep (cold caches, serialised replays: shares, not absolutes)"]
    dram = 0.0
    for key in KEYS:
        if key in col:
            vals = [r[col[key]] for r in body]
            lines.append(f"{key:70s} {units[col[key]]:10s} {' '.join(vals)}")
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if key in col:
            dram += sum(float(r[col[key]].replace(",", "")) for r in body) / len(body) * UNIT.get(units[col[key]], 1.0)
    traffic[k] = {"dram_bytes_per_launch": dram, "round": rnd}
    try:
        top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, k, "25"], capture_output=True, text=True).stdout
        lines += ["", "# hottest source lines (warp-stall samples / executed warp instructions)", top]
    except Exception as e:  # noqa: BLE001
        lines.append(f"# source attribution failed: {e}")
    open(os.path.join(ROOT, "profiles", f"{rnd}_{k}.txt"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/{rnd}_{k}.txt", "dram bytes/launch", dram)
json.dump(traffic, open(tp, "w"), indent=1)
