"""Turn ncu outputs into the markdown summaries kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md> [title]
    python tools/ncu_summary.py full <prof.ncu-rep> <out.md> [title]
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out, title):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        n = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").strip()
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                f"(cold-cache, serialised: compare SHARES, not absolutes).  Total {tot:.0f} us over "
                f"{sum(a[0] for a in agg.values())} launches.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n[:100]}` | {c} | {t:.1f} | {100 * t / tot:.1f} % |\n")


WANT = [("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, lab) for k, lab in WANT if k in idx]
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, one row per captured launch.\n\n")
        f.write("| kernel | " + " | ".join(f"{lab} [{units[idx[k]]}]" for k, lab in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "").strip()
            f.write(f"| `{name[:70]}` | " + " | ".join(r[idx[k]][:12] for k, _ in cols) + " |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
