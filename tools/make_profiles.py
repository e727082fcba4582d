"""Build the tracked summaries under profiles/ from the raw ncu CSVs / bench outputs in gpurun_out/.

    python tools/make_profiles.py [round-tag]        (default r01)

Inputs (written by tools/run_profiles.sh / tools/run_check.sh on the GPU box):
  gpurun_out/launches_{train,infer}_*.csv   ncu --metrics gpu__time_duration.sum launch lists
  gpurun_out/prof_{conv,wgrad,bw*}_raw.csv  `ncu -i <rep> --page raw --csv` of the --set full captures
  gpurun_out/bw_bench*.txt                  tools/bw_bench.py output
Outputs: profiles/<tag>_*.md and profiles/roofline_traffic.json (read by bench.py for roofline.traffic).
"""
import collections
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def clean(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("mau::<unnamed>::", "").strip()


def latest(pattern):
    f = sorted(glob.glob(os.path.join(G, pattern)), key=os.path.getmtime)
    return f[-1] if f else None


def launches(path, out, title, first_kernel="nchw_to_nhwc"):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    names = [clean(r["Kernel Name"]) for r in rows]
    starts = [i for i, n in enumerate(names) if n.startswith(first_kernel)]
    # one whole step: the LAST one (profiling runs end right after the timed step: --no-e2e --no-kernel-pass; the step before it
    # is the warm-up step, which also creates the optimizer state)
    a, b = (starts[-1], len(rows)) if starts else (0, len(rows))
    while a > 0 and names[a - 1].startswith(("pack_w_fwd", "mlp_fwd", "lstm_fwd", "linear_fwd")):
        a -= 1                                   # this step's weight packs / encoders are launched ahead of the layout kernel
    agg, tot = collections.OrderedDict(), 0.0
    for i in range(a, b):
        v = float(rows[i]["Metric Value"].replace(",", ""))
        v = v / 1000 if rows[i]["Metric Unit"] == "ns" else (v * 1000 if rows[i]["Metric Unit"] == "ms" else v)
        e = agg.setdefault(names[i], [0, 0.0])
        e[0] += 1
        e[1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(path, ROOT)}` -- `ncu --metrics gpu__time_duration.sum --clock-control none`, "
                f"the last whole step (launches {a}..{b - 1} of the capture).  Launches under ncu are serialised and cold-cache: compare "
                f"SHARES, not absolutes.  Total {tot:.0f} us over {b - a} launches.\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n[:90]}` | {c} | {t:.1f} | {100 * t / tot:.1f} % |\n")
    return agg, tot


WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("dram__bytes_read.sum", "DRAM rd MB"),
        ("dram__bytes_write.sum", "DRAM wr MB"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs")]


def full(path, out, title, note=""):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, lab) for k, lab in WANT if k in idx]
    # ncu picks one unit per column and per file: normalise bytes to MB and durations to us
    SC = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6, "ns": 1e-3, "us": 1.0, "usecond": 1.0,
          "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}
    scale = {k: SC.get(units[idx[k]], 1.0) for k, _ in cols}
    recs = []
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(path, ROOT)}` = `ncu -i <report> --page raw --csv` of an "
                "`ncu --set full --clock-control none --import-source on` capture (the .ncu-rep itself exceeds the gpurun_out size cap "
                f"and is converted on the GPU box).  One row per captured launch.  {note}\n\n")
        f.write("| kernel | grid | " + " | ".join(lab for _, lab in cols) + " |\n|---|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = clean(r[idx["Kernel Name"]])
            vals = []
            for k, lab in cols:
                v = r[idx[k]].replace(",", "")
                try:
                    x = float(v) * scale[k]
                    vals.append(f"{x:.1f}" if lab != "regs" else f"{int(x)}")
                except ValueError:
                    vals.append(v)
            f.write(f"| `{name[-48:]}` | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |\n")
            rec = {"kernel": name}
            for k, lab in cols:
                try:
                    rec[lab] = float(r[idx[k]].replace(",", "")) * scale[k]
                except ValueError:
                    pass
            recs.append(rec)
    return recs


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(P, exist_ok=True)
    traffic = {}
    lt, li = latest("launches_train_*.csv"), latest("launches_infer_*.csv")
    if lt:
        launches(lt, os.path.join(P, f"{tag}_launches_train.md"), f"{tag}: training step (config 3, B=16), every launch")
    if li:
        launches(li, os.path.join(P, f"{tag}_launches_infer.md"), f"{tag}: inference step (config 2, B=16), every launch")
    pc, pw, pi = latest("prof_conv_raw.csv"), latest("prof_wgrad_raw.csv"), latest("prof_conv_infer_raw.csv")

    def per_launch(recs):
        n = len(recs)
        by = sum(r.get("DRAM rd MB", 0) + r.get("DRAM wr MB", 0) for r in recs) * 1e6
        return {"launches": n, "dram_bytes_per_launch": by / max(n, 1)}
    if pi:
        recs = full(pi, os.path.join(P, f"{tag}_ncu_conv_v2_infer.md"), f"{tag}: conv3x3_tc_v2 / conv3x3_tc_col3 kernels, the 18 launches of one inference forward (config 2, B=16)",
                    "Template arguments v2 <BN, MT, NBUF, NA, NB, NSTG, BT, EM, BRES[, STATS]>, col3 <NA, NB, EM[, STATS]>; ")
        traffic["conv_infer"] = per_launch(recs)
    if pc:
        recs = full(pc, os.path.join(P, f"{tag}_ncu_conv_v2.md"), f"{tag}: conv3x3_tc_v2 / conv3x3_tc_col3 kernels (forward + dgrad launches of one training step; STATS = 1: forward with the statistics warps, 384 threads)",
                    "Template arguments v2 <BN, MT, NBUF, NA, NB, NSTG, BT, EM, BRES[, STATS]>, col3 <NA, NB, EM[, STATS]>; ")
        traffic["conv_fwd_dgrad"] = per_launch(recs)
    if pw:
        recs = full(pw, os.path.join(P, f"{tag}_ncu_wgrad_v2.md"), f"{tag}: weight-gradient kernels (the 18 launches of one backward: wgrad3x3_tc_v2 split-K / stream-K, wgrad3x3_tc_pair for the <= 64-channel sides)",
                    "Template arguments v2 <BN, SWAP, STAGES>, pair <BN, STAGES>; ")
        traffic["wgrad"] = per_launch(recs)
    for src in sorted(glob.glob(os.path.join(G, "prof_bw*_raw.csv"))):
        base = os.path.basename(src).replace("_raw.csv", "")
        what = base
        if base == "prof_bw3":
            what = ("touched at the end of round 2 (prof_bw3: bilinear_bwd_lean at 127 registers -- the 80-register build came one call "
                    "later --, head_bwd and bn_bwd_apply with whole-wave grids)")
        full(src, os.path.join(P, f"{tag}_ncu_{base}.md"), f"{tag}: bandwidth-bound kernels {what if base == 'prof_bw3' else '(' + what + ')'}, tools/bw_bench.py shapes at B=16")
    bw = latest("bw_bench*.txt")
    if bw:
        with open(os.path.join(P, f"{tag}_bw_bench.md"), "w") as f:
            f.write(f"# {tag}: achieved HBM bandwidth per bandwidth-bound kernel\n\nSource: `{os.path.relpath(bw, ROOT)}` (`python tools/bw_bench.py`): "
                    "algorithmic bytes / CUDA-event time, tensors rotated over > L2-sized copies, B = 16, against the measured copy "
                    "bandwidth in MEASURED_PEAKS.json.\n\n```\n" + open(bw).read() + "```\n")
            later = os.path.join(G, "r02o_bw_occ3.txt")
            if tag == "r02" and os.path.exists(later):
                f.write("\nThe table above is from the call before the last change to `bilinear_bwd_lean_kernel` (80 registers, three resident "
                        "blocks per SM instead of 127 / two).  The final kernel, measured in the next call (`tools/r02_occ.sh`):\n\n```\n"
                        + open(later).read() + "```\n")
    if traffic:
        # bench.py's roofline.traffic: mean DRAM bytes per timed conv-family launch (config 2 times the forward convs,
        # config 3 forward + dgrad + wgrad).  The conv capture is of training-mode launches (same shapes as inference).
        out = {"source": f"profiles/{tag}_ncu_conv_v2.md, profiles/{tag}_ncu_wgrad_v2.md (dram__bytes_read.sum + dram__bytes_write.sum)"}
        c, w, ci = traffic.get("conv_fwd_dgrad"), traffic.get("wgrad"), traffic.get("conv_infer")
        if ci:
            out["config2"] = {"dram_bytes_per_launch": ci["dram_bytes_per_launch"], "launches": ci["launches"]}
        elif c:
            out["config2"] = {"dram_bytes_per_launch": c["dram_bytes_per_launch"], "launches": c["launches"]}
        if c and w:
            n = c["launches"] + w["launches"]
            out["config3"] = {"dram_bytes_per_launch": (c["dram_bytes_per_launch"] * c["launches"] + w["dram_bytes_per_launch"] * w["launches"]) / n,
                              "launches": n}
        json.dump(out, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(P)))


if __name__ == "__main__":
    main()
