"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
    python tools/summarize_ncu.py gpurun_out/r01_attn.ncu-rep profiles/r01_attn_ncu.txt
    python tools/summarize_ncu.py --launches gpurun_out/r01_launches.csv profiles/r01_launches_by_kernel.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on   ({path})\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"\nkernel: {d.get('Kernel Name')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"  {k:72s} {d[k]:>16s} {u.get(k, '')}\n")
            try:
                t = float(d["gpu__time_duration.sum"])
                tu = u["gpu__time_duration.sum"]
                sec = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu.replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1e-9)
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                bu = u["dram__bytes_read.sum"]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(bu, 1)
                f.write(f"  -> dram traffic {(rd + wr) * mult / 1e6:.1f} MB, {(rd + wr) * mult / sec / 1e9:.1f} GB/s over {sec * 1e6:.1f} us\n")
            except (KeyError, ValueError):
                pass


def launches(path, out):
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:110]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        agg[name][0] += 1
        agg[name][1] += ns
        total += ns
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# {sum(v[0] for v in agg.values())} launches, {total / 1e6:.2f} ms summed\n")
        f.write(f"{'share':>7s} {'ms':>9s} {'launches':>8s}  kernel\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
            f.write(f"{100 * ns / total:6.2f}% {ns / 1e6:9.3f} {n:8d}  {name}\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2])
