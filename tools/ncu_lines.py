"""Per-source-line view of an ncu capture taken with `--set full --import-source on` (kernels built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [top_n]

Prints the CUDA source lines with the most warp-stall samples and each line's dominant stall reasons: the tool used to
decide what to restructure in the latency-bound kernels (decoder rollout / BPTT), where whole-kernel counters say
little.  Summaries are committed under profiles/."""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    data, hdr, fname = [], None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr is not None and r[0].strip().isdigit() and len(r) == len(hdr):
            data.append((fname, r))
    if hdr is None:
        print(out[:2000])
        return
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, h):
        try:
            return int(float(r[ix[h]]))
        except (ValueError, KeyError):
            return 0

    tot = sum(num(r, "# Samples") for _, r in data) or 1
    print(f"# {path}: {len(data)} source lines with code, {tot} warp-stall samples")
    agg = sorted(((sum(num(r, h) for _, r in data), h) for h in stalls), reverse=True)[:8]
    print("# stall reasons:", ", ".join(f"{h[6:]} {100 * n / tot:.0f}%" for n, h in agg))
    for f, r in sorted(data, key=lambda fr: -num(fr[1], "# Samples"))[:top]:
        s = num(r, "# Samples")
        if s == 0:
            break
        why = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:3]
        why = ", ".join(f"{n_}:{100 * c / s:.0f}%" for c, n_ in why if c)
        print(f"{100 * s / tot:5.1f}%  {f}:{r[0]:>4s}  {r[1].strip()[:90]:90s}  [{why}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
