"""Turn the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01

  gpurun_out/launches_<round>.csv            -> profiles/<round>_launches_bench.md (+ the per-kernel table as csv)
  gpurun_out/prof_<round>_<name>.ncu-rep     -> profiles/<round>_<name>_ncu.md (key raw metrics per launch)
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg"]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|rslo::|void ", "", name)
    return re.sub(r"\(.*", "", name)[:100]


def launches(rnd):
    path = os.path.join(SRC, f"launches_{rnd}.csv")
    if not os.path.exists(path):
        return
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines[start:]):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg[short(row["Kernel Name"])]
        a[0] += 1
        a[1] += v
        tot += v
    own = sum(a[1] for k, a in agg.items() if k.startswith("k_"))
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    with open(os.path.join(OUT, f"{rnd}_launches_bench.md"), "w") as f:
        cmd = "bench.py --steps 1 --warmup 3 --pairs-per-gpu 1" if rnd.startswith("r01") else "bench.py --steps 1 --warmup 3 (the bench's train step: 2 pairs per GPU)"
        f.write(f"# ncu launch list — one timed step of `{cmd}` ({rnd})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` (cold-cache, "
                "serialised: compare SHARES, not absolutes).\n\n")
        f.write(f"launches: {sum(a[0] for a in agg.values())}, summed kernel time: {tot / 1e3:.2f} ms; "
                f"own kernels (librslo_b200, `k_*`): {own / 1e3:.2f} ms = {100 * own / tot:.1f}% of the step\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, a in rows[:60]:
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |\n")
    with open(os.path.join(OUT, f"{rnd}_launches_bench.csv"), "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, a in rows:
            f.write(f"\"{k}\",{a[0]},{a[1]:.2f},{a[1] / tot:.5f}\n")
    print("launch list:", len(rows), "kernels,", f"{tot / 1e3:.2f} ms")


def reports(rnd):
    """prof_<round>_<name>.ncu-rep (read with ncu here) or prof_<round>_<name>.raw.csv (already exported on the GPU
    box with `ncu -i ... --page raw --csv`: the reports themselves exceed what gpurun brings back)."""
    for fn in sorted(os.listdir(SRC)):
        m = re.match(rf"prof_{rnd}_(.+?)(\.ncu-rep|\.raw\.csv)$", fn)
        if not m:
            continue
        if fn.endswith(".ncu-rep"):
            raw = subprocess.run(["ncu", "-i", os.path.join(SRC, fn), "--page", "raw", "--csv"], capture_output=True,
                                 text=True).stdout
        else:
            raw = open(os.path.join(SRC, fn)).read()
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        with open(os.path.join(OUT, f"{rnd}_{m.group(1)}_ncu.md"), "w") as f:
            f.write(f"# ncu --set full: `{m.group(1)}` ({rnd})\n\n")
            f.write("Captured inside `bench.py --steps 1 --warmup 3` (the bench's train step, 2 pairs per GPU; "
                    "`--clock-control none --import-source on --profile-from-start off`); one column per captured launch.\n\n")
            f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
            f.write("|---|---|" + "---:|" * len(data) + "\n")
            kn = hdr.index("Kernel Name")
            f.write("| kernel | | " + " | ".join(short(r[kn]) for r in data) + " |\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        print("report:", fn)
        # DRAM traffic per launch (bench.py reads it back into roofline.traffic)
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = [float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) +
                   float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0) for r in data]
            name = {"conv2d_wgrad_tc": "conv2d_tc_wgrad", "spconv_fwd": "spconv_forward"}.get(m.group(1), m.group(1))
            TRAFFIC[name] = {"dram_bytes_per_launch_avg": sum(tot) / len(tot), "launches_captured": len(tot),
                             "source": f"profiles/{rnd}_{m.group(1)}_ncu.md (ncu --set full inside bench.py --steps 1)"}
        except ValueError:
            pass
    if TRAFFIC:
        import json
        with open(os.path.join(OUT, f"{rnd}_traffic.json"), "w") as f:
            json.dump(TRAFFIC, f, indent=1)


TRAFFIC = {}


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(rnd)
    reports(rnd)
