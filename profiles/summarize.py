"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_x.ncu-rep   > profiles/r1_x.txt
    python profiles/summarize.py traffic  gpurun_out/traffic.csv "note" > profiles/r1_traffic_infer.json
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
        if "spin_kernel" in name:      # torch.cuda._sleep: parks the GPU before bench.py's instrumented pass, not work
            continue
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


def traffic(path, note=""):
    """launch list with dram__bytes_read/write -> JSON (per kernel: launches, us, bytes) incl. the `gemm_tc_all` entry
    bench.py reads for roofline.traffic (mean DRAM bytes per tcgen05 GEMM launch of one forward)."""
    import json
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = OrderedDict()
    seen = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        k = per.setdefault(name, {"launches": 0, "us": 0.0, "read": 0.0, "write": 0.0})
        metric, unit = r[ix["Metric Name"]], r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        if metric == "gpu__time_duration.sum":
            k["us"] += v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
            if seen.get(r[ix["ID"]]) is None:
                seen[r[ix["ID"]]] = 1
                k["launches"] += 1
        elif metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            k["read" if "read" in metric else "write"] += v * scale
    gem = [v for n, v in per.items() if "gemm_tc" in n or "posconv_tc" in n or "gemm_ln" in n or "enc_block" in n]
    tot = sum(v["read"] + v["write"] for v in gem)
    n = sum(v["launches"] for v in gem)
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; "
                     + note, "kernels": per,
           "gemm_tc_all": {"launches": n, "dram_bytes": tot, "dram_bytes_per_launch": tot / max(n, 1)}}
    print(json.dumps(out, indent=1))


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
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    # a report with several kernels repeats the header row per kernel: keep the data rows only
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].strip().isdigit()]
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
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
