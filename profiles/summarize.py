"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_x.ncu-rep   > profiles/r1_x.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"# launch list: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:")
    print("# compare SHARES, not absolutes)")
    print(f"{'kernel':80s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:80]:80s} {n:8d} {us:10.1f} {100 * us / total:6.1f}%")
    print(f"{'TOTAL':80s} {sum(a[0] for a in agg.values()):8d} {total:10.1f}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none (one launch)")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name", "?")[:120])
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print(f"  {h:75s} {v:>16s} {u}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hi is None:
        return
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    agg = sorted(((sum(int(r[ix[h]]) for r in data), h) for h in stalls), reverse=True)[:6]
    print(f"  SASS instructions: {len(data)}; warp-stall samples: {tot}")
    print("  stall reasons:", ", ".join(f"{h[6:]} {100 * n / tot:.0f}%" for n, h in agg))
    print("  hottest instructions:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:8]:
        s = int(r[ix["# Samples"]])
        print(f"    {100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
