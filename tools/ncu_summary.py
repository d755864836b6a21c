"""Summarise ncu output into the text files kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> "<header line>"      # --metrics gpu__time_duration.sum --csv log
  python tools/ncu_summary.py full <report.ncu-rep> "<header line>"        # --set full capture, one entry per distinct kernel

Read-only helper for the measurement rows of DESIGN.md; nothing on the product path imports it.
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def launches(path, header):
    rows = []
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else v * (1.0 if unit in ("us", "usecond") else 1000.0)
        rows.append((r["Kernel Name"], r.get("Grid Size", ""), v_us))
    total = sum(r[2] for r in rows)
    agg = collections.OrderedDict()
    for name, grid, us in rows:
        k = (name.split("(")[0].replace("void ", "").replace("mocr::", ""), grid)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    print(f"# {header}")
    print("# per-launch times under ncu are cold-cache and serialised (no programmatic-dependent-launch overlap): compare SHARES")
    print(f"# total {total:.1f} us over {len(rows)} launches")
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:9.1f} us {100 * us / total:5.1f}%  n={n:4d} avg={us / n:7.2f} us  {name} grid={grid}")


def _raw(path):
    """Rows of `ncu -i <report> --page raw --csv` (or of that CSV saved on the GPU box, when the report itself was too big to bring back)."""
    if path.endswith(".csv"):
        with open(path, newline="") as f:
            out = f.read()
        out = out[out.find('"ID"'):]
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(path, header):
    rd = _raw(path)
    names, units = rd[0], rd[1]
    col = {n: i for i, n in enumerate(names)}
    seen = set()
    print(f"# {header}")
    print("# one entry per distinct kernel (first captured launch); values as reported by ncu")
    for row in rd[2:]:
        kname = row[col["Kernel Name"]]
        key = (kname, row[col["launch__grid_size"]] if "launch__grid_size" in col else "")
        if key in seen:
            continue
        seen.add(key)
        print("-" * 100)
        print(f"{'Kernel Name':<88} {kname}")
        for m in KEEP:
            if m in col:
                print(f"{m:<88} {row[col[m]]} {units[col[m]]}")


def traffic(path, command):
    """JSON for bench.py's roofline.traffic: DRAM bytes (read + write) per launch, averaged per kernel over the capture."""
    import json
    rd = _raw(path)
    names, units = rd[0], rd[1]
    col = {n: i for i, n in enumerate(names)}

    def scaled(row, metric, to):
        v = float(row[col[metric]].replace(",", ""))
        u = units[col[metric]].lower()
        f = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
             "msecond": 1e3}[u]
        return v * f

    agg = collections.OrderedDict()
    for row in rd[2:]:
        k = row[col["Kernel Name"]].replace("mocr::", "")
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += scaled(row, "dram__bytes_read.sum", "byte") + scaled(row, "dram__bytes_write.sum", "byte")
        a[2] += scaled(row, "gpu__time_duration.sum", "us")
    doc = {"command": command,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches of each kernel (cold cache, serialised)",
           "kernels": [{"kernel": k, "launches_captured": n, "dram_bytes_per_launch": b / n, "gpu_time_us_per_launch_cold": t / n}
                       for k, (n, b, t) in agg.items()]}
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
