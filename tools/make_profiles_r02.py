#!/usr/bin/env python
"""Round-2 summaries under profiles/ from the raw outputs of the gpurun calls in gpurun_out/ (scratch, not tracked):

    python tools/make_profiles_r02.py

profiles/r02_training_step.md   the A/B log of every change to the training step (bench lines of tools/r02_call*.sh)
profiles/r02_timeline.md        CUPTI kernel timeline of one training step at 1 / 2 / 8 GPUs (tools/timeline.py)
profiles/r02_multi_gpu.md       data-parallel parity, NCCL CTA / wire-format sweep, H2D concurrency (tools/r02_n2.sh, r02_n8.sh)
profiles/r02_ssim.md            SSIM kernels on hardware (tools/ssim_gpu_check.py) and their cost in the step
and copies of the bench JSON lines they quote (profiles/r02_bench_*.json).
"""
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def bench(name):
    p = os.path.join(G, name + ".json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def row(label, name, extra=""):
    d = bench(name)
    if d is None:
        return f"| {label} | — | — | — | — | `{name}.json` missing |\n"
    shutil.copyfile(os.path.join(G, name + ".json"), os.path.join(P, name.replace("r02c", "r02_bench_c").replace("r02n", "r02_bench_n").replace("r02g_", "r02_bench_final_").replace("r02o_", "r02_bench_occ_").replace("r02p_", "r02_bench_last_") + ".json"))
    s = d.get("sustained") or {}
    e = d.get("e2e") or {}
    return (f"| {label} | {d['ms_per_step']:.3f} | {d['value']:.0f} | {('%.3f' % s['ms_per_step']) if s.get('ms_per_step') else '—'} | "
            f"{('%.0f' % e['value']) if e.get('value') else '—'} | {extra} |\n")


def main():
    os.makedirs(P, exist_ok=True)
    with open(os.path.join(P, "r02_training_step.md"), "w") as f:
        f.write("# r02: the training step (config 3: U-Net + metadata, B = 16, 23x250x250, bf16, fwd + L1 + bwd + fused AdamW), change by change\n\n"
                "Every row is one `python bench.py --config 3 --no-cpu-baseline` line measured on a B200 through `gpurun` (the JSON lines are kept\n"
                "next to this file as `r02_bench_*.json`); A/B pairs were taken back to back in ONE call on ONE box, boxes differ by up to +-3 %\n"
                "(compare within a pair).  ms/step = CUDA events around 20 steps after 5 warm-up steps (boost clocks); sustained = the same loop\n"
                "for >= 1-2 s.  Round 1 ended at 8.10 ms (`profiles/r01_bench_c3.json`).\n\n"
                "| change (A/B switch) | ms/step | tiles/s | sustained ms/step | e2e tiles/s | note |\n|---|---:|---:|---:|---:|---|\n")
        f.write(row("round-1 kernels, new bench", "r02c1_bench_c3_bnunfused", "baseline of the round (call 1)"))
        f.write(row("BatchNorm as ONE cooperative launch per direction (stats, grid barrier, apply in reverse for L2 reuse)", "r02c1_bench_c3",
                    "slower: every launch +5..20 us (cooperative launch cannot overlap its neighbours' tails; 128 MB tensors do not stay in L2) -> dropped"))
        f.write(row("`MAU_WGRAD_PAIR=0` (round-1 weight-gradient kernel everywhere)", "r02c3_bench_c3_nopair", "call 3"))
        f.write(row("tap-pair weight-gradient kernel for the <= 64-channel sides", "r02c3_bench_c3_pair",
                    "wgrad 64->64 @250: 163 -> 93 us, 192->64: 324 -> 195 us, 23->64: 153 -> 84 us"))
        f.write(row("`MAU_FLAGS=8192` (weight gradients on the caller's stream)", "r02c4_bench_c3_nooverlap", "call 4; the 20-step figure of this run hit a host hiccup, compare the sustained column"))
        f.write(row("weight gradients on a second stream (overlap BatchNorm / pool / bilinear backward)", "r02c4_bench_c3", ""))
        f.write(row("`MAU_FLAGS=16384` (separate bn_stats pass)", "r02c7_bench_c3_nostats", "call 7"))
        f.write(row("BatchNorm statistics inside the convolution kernels (idle warps) + weight packs ahead on the second stream", "r02c7_bench_c3", ""))
        f.write(row("call 9 (a slower box), late fork: `MAU_FLAGS=16384` (separate bn_stats pass)", "r02c9_bench_c3_nostats", "statistics on FOUR extra warps from here on (384-thread CTAs)"))
        f.write(row("call 9: `MAU_WGRAD_FORK_LATE=1` (second stream forks behind the data gradient), run 1", "r02c9_bench_c3", "the weight gradients pile up and leave a tail: slower"))
        f.write(row("call 9: the same, run 2", "r02c9_bench_c3_b", ""))
        f.write(row("call 9: second stream forks right after the BatchNorm backward, data gradient enqueued first (final)", "r02c9_bench_c3_forkearly", "+ one memset for all BatchNorm sums"))
        f.write(row("the same with the reference's default criterion `l1-gradient-ssim` (separable SSIM kernels)", "r02c4_bench_c3_ssim",
                    "first version of the SSIM kernels (121 taps per window): +0.54 ms per step (`r02_bench_c1_bench_c3_ssim.json`)"))
        f.write(row("first SSIM kernels (direct 121-tap form), for reference", "r02c1_bench_c3_ssim", "call 1, on the round-1 step"))
        f.write(row("call 10 (a faster box): the step as of call 9 (`MAU_BILINEAR_BWD=stream MAU_WHOLE_WAVES=0` at that commit), run 1", "r02c10_c3_old_a", "boxes differ: compare within calls 10 / final / occupancy"))
        f.write(row("call 10: the same, run 2", "r02c10_c3_old_b", ""))
        f.write(row("call 10: rows-first bilinear backward (one thread per OUTPUT column, columns meet in shared memory, one barrier per input row) + whole-wave grids, run 1", "r02c10_c3_new_a",
                    "bilinear_bwd alone 96.4 us either way; in the step +0.1 ms (the barrier serialises the load latency of a block) -> the rows-first kernel was dropped"))
        f.write(row("call 10: the same, run 2", "r02c10_c3_new_b", ""))
        f.write(row("call 10: rows-first bilinear backward without the whole-wave grids", "r02c10_c3_vh_only",
                    "whole-wave grids: head_bwd 97.9 -> 74.3 us, bn_bwd_apply @62^2 28.8 -> 23.3 us alone (`tools/bw_bench.py`); below the noise of the step"))
        f.write(row("final call: first-generation streaming bilinear backward (`MAU_BILINEAR_BWD=stream` at that commit), run 1", "r02g_c3_first_a", ""))
        f.write(row("final call: the same, run 2", "r02g_c3_first_b", ""))
        f.write(row("final call: lean streaming bilinear backward (warp-uniform 2/4/6 unrolling, no per-lane predicates, pointer + 32-bit offset), run 1", "r02g_c3_lean_a",
                    "bilinear_bwd alone: 96.4 -> 84.2 us (16 x 250^2 x 128), 56.0 -> 51.4 us (16 x 124^2 x 256)"))
        f.write(row("final call: the same, run 2", "r02g_c3_lean_b", ""))
        f.write(row("occupancy call: lean kernel, 127 registers / two resident blocks (`MAU_BILINEAR_OCC=2` at that commit)", "r02o_c3_occ2", ""))
        f.write(row("occupancy call: lean kernel compiled for three resident blocks (80 registers; only the rare 6-contribution path spills) (final)", "r02o_c3_occ3",
                    "bilinear_bwd alone: 84.0 -> 75.5 us, 49.8 -> 41.1 us"))
        f.write(row("last call: default build (= final)", "r02p_c3_default", ""))
        f.write(row("last call: BatchNorm backward kernels compiled for three resident blocks (80 registers, spills in the loop), head backward for four", "r02p_c3_occ",
                    "bn_bwd_reduce @250^2 51.1 -> 83.8 us, bn_bwd_apply 70.9 -> 105.0 us, head_bwd 73.6 -> 72.2 us: spilled coefficients cost more than the extra warps hide -> not adopted"))
        f.write("\nU-Net++ (config 4, B = 16):\n\n| change | ms/step | tiles/s | sustained | e2e | note |\n|---|---:|---:|---:|---:|---|\n")
        f.write(row("`MAU_FLAGS=8192`", "r02c4_bench_c4_nooverlap", "call 4"))
        f.write(row("second-stream weight gradients", "r02c4_bench_c4", ""))
        f.write(row("+ statistics in the convolution kernels", "r02c6_bench_c4", "call 6"))
        lay = os.path.join(G, "r02c3_layers_c3_pair.txt")
        if os.path.exists(lay):
            f.write("\nPer-launch times of the weight-gradient kernels with the tap-pair kernel (CUDA events around each launch, call 3):\n\n```\n")
            f.write("".join(l for l in open(lay) if "wgrad" in l and l.startswith("k:")))
            f.write("```\n")
    # the default bench line of the final call (the driver's form) next to the CPU reference arm of the previous final run
    for src, dst in (("r02g_bench_default.json", "r02_bench_default.json"),):
        if os.path.exists(os.path.join(G, src)):
            shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
    # timelines
    with open(os.path.join(P, "r02_timeline.md"), "w") as f:
        f.write("# r02: kernel timeline of one training step (CUPTI through torch.profiler; nsys is not in the image)\n\n"
                "`python tools/timeline.py --config 3` (under torchrun for N > 1, rank 0 reports).  `span` = first kernel start to last kernel end of\n"
                "the middle step of three; `summed` = sum of all kernel durations; `union busy` = time at least one kernel runs; `overlapped` =\n"
                "summed - union (work that ran concurrently on the second stream / the NCCL stream).\n\n")
        for tag, title in (("r02c5_timeline.txt", "1 GPU, before the statistics fusion (call 5)"), ("r02c7_timeline.txt", "1 GPU, final kernels (call 7)"),
                           ("r02n2_timeline.txt", "2 GPUs, NCCL capped at 4 CTAs"), ("r02n8_timeline.txt", "8 GPUs, NCCL capped at 8 CTAs")):
            p = os.path.join(G, tag)
            if not os.path.exists(p):
                continue
            txt = open(p).read()
            m = re.search(r"(step span.*?largest idle gaps[^\n]*)", txt, re.S)
            if m:
                f.write(f"## {title}\n\n```\n{m.group(1)}\n```\n\n")
            nccl = sorted(set(re.findall(r"NCCL INFO (NVLS multicast support[^\n]*|.*(?:Ring|Tree|NVLS|algo)[^\n]*)", txt)))[:6]
            if nccl:
                f.write("NCCL_DEBUG=INFO excerpts: " + "; ".join(f"`{x.strip()[:110]}`" for x in nccl) + "\n\n")
            j = os.path.join(G, tag.replace("_timeline.txt", "") + "_timeline_c3_n%s.json" % ("1" if "c" in tag[:6] else tag[4]))
            if os.path.exists(j):
                recs = json.load(open(j))
                names = sorted({r["name"][:100] for r in recs if "nccl" in r["name"].lower()})
                if names:
                    f.write("NCCL kernels in the step: " + "; ".join(f"`{n}`" for n in names) + "\n\n")
    # multi-GPU
    with open(os.path.join(P, "r02_multi_gpu.md"), "w") as f:
        f.write("# r02: data-parallel training and multi-GPU measurements\n\n## Parity (tools/dp_parity.py, SyncBN, fp32 mode)\n\n```\n")
        for n in ("r02n2_dp_parity_fp32.txt", "r02n2_dp_parity_bf16.txt", "r02n8_dp_parity.txt"):
            p = os.path.join(G, n)
            if os.path.exists(p):
                f.write("".join(l for l in open(p) if "[dp_parity]" in l))
        f.write("```\n\n## Config 3 at 2 GPUs: NCCL CTA cap and wire format of the gradient all-reduce\n\n"
                "| wire / NCCL CTAs (= SMs the persistent backward kernels leave free) | ms/step | tiles/s | sustained ms/step | e2e tiles/s | |\n|---|---:|---:|---:|---:|---|\n")
        for w, c in (("fp32", 8), ("fp32", 4), ("bf16", 4), ("bf16", 2)):
            f.write(row(f"{w}, {c} CTAs", f"r02n2_bench_c3_{w}_ctas{c}"))
        f.write("\nOne GPU, same kernels: 7.07 ms/step (`r02_bench_c7_bench_c3.json`).  The bf16 wire format halves the payload but adds two cast launches per\n"
                "bucket on the communication stream and loses; 8 CTAs beat 4 -- the all-reduce finishes earlier and the tail after the last\n"
                "backward kernel shrinks.  Default: fp32 wire, 8 CTAs.\n\n## 8 GPUs\n\n| run | ms/step | tiles/s | sustained | e2e | |\n|---|---:|---:|---:|---:|---|\n")
        f.write(row("default bench line of that call (config 3 + riders, 8 NCCL CTAs), 8 GPUs", "r02n8_bench_default"))
        for c in (12, 16, 24, 32):
            f.write(row(f"config 3, {c} NCCL CTAs (second call, another box)", f"r02n8_bench_c3_ctas{c}"))
        for nm, lab in (("base", "12 CTAs, third call"), ("simple", "12 CTAs, NCCL_PROTO=Simple"), ("bucket32", "12 CTAs, 32 MB buckets")):
            f.write(row(f"config 3, {lab}", f"r02n8_bench_c3_{nm}"))
        f.write("\n`NCCL_ALGO=allreduce:nvls` is refused by NCCL here (\"no algorithm/protocol available for function AllReduce with datatype ncclFloat32\", with 4, 8 and 12 CTAs; fourth call, base 7.44 ms).  Neither the Simple protocol nor larger buckets change the step: the\n"
                "all-reduce is hidden; what it costs is the SMs it takes from the persistent kernels.\n")
        f.write("\n## Inference (config 2) through host buffers at 4 / 8 GPUs: the reference's fp32 NCHW contract vs tiles staged as bf16 NHWC\n\n"
                "| GPUs | resident tiles/s | e2e, fp32 NCHW (92 MB per step and rank) | e2e, staged bf16 NHWC (48 MB) |\n|---:|---:|---:|---:|\n")
        for n, nm in ((1, "r02c9_bench_c2"), (4, "r02n4_bench_c2"), (8, "r02n8_bench_c2")):
            d2 = bench(nm)
            if d2 and d2.get("e2e"):
                shutil.copyfile(os.path.join(G, nm + ".json"), os.path.join(P, nm.replace("r02c", "r02_bench_c").replace("r02n", "r02_bench_n") + ".json"))
                st = d2["e2e"].get("staged_bf16_nhwc", {}).get("value")
                f.write(f"| {n} | {d2['value']:.0f} | {d2['e2e']['value']:.0f} | {st:.0f} |\n" if st else f"| {n} | {d2['value']:.0f} | {d2['e2e']['value']:.0f} | — |\n")
        d = bench("r02n8_bench_default")
        if d and "inference" in d:
            f.write(f"\nRiders of the 8-GPU line: inference {d['inference']['value']:.0f} tiles/s (e2e {d['inference']['e2e']['value']:.0f}); "
                    + ", ".join(f"{k} {v['value']:.0f}" for k, v in d["riders"].items()) + " tiles/s.\n")
        h = os.path.join(G, "r02n8_h2d_diag.jsonl")
        if os.path.exists(h):
            f.write("\n## Concurrent host-to-device bandwidth per GPU (tools/h2d_diag.py): the limiter of the inference e2e number at N >= 4\n\n"
                    "Every rank copies its pinned 92 MB batch-16 input block to its GPU back to back for 1 s.\n\n| ranks | per-GPU GB/s (min / mean) | sum GB/s |\n|---:|---:|---:|\n")
            for l in open(h):
                try:
                    r = json.loads(l)
                    v = r["h2d"]["fp32 NCHW (the reference's contract)"]
                    f.write(f"| {r['n_gpus']} | {v['per_gpu_gbs_min']:.1f} / {v['per_gpu_gbs_mean']:.1f} | {v['sum_gbs']:.1f} |\n")
                except Exception:
                    pass
    s = os.path.join(G, "r02c5_ssim_check.json")
    if os.path.exists(s):
        d = json.load(open(s))
        with open(os.path.join(P, "r02_ssim.md"), "w") as f:
            f.write("# r02: SSIM loss kernels on a B200\n\n`python tools/ssim_gpu_check.py` (device kernels vs the torch restatement `oracle/ssim_oracle.py`; parity with `piq` itself stays\n"
                    "UNPINNED: piq is absent from the image and `pip download piq` has no index to reach).  Shapes with min(H, W) >= 384 run piq's\n"
                    "average-pool path (f = round(min / 256)).\n\n| shape | SSIM abs err | total-loss abs err | gradient rel err (L1+grad+SSIM) | SSIM gradient rel err |\n|---|---:|---:|---:|---:|\n")
            for c in d["cases"]:
                f.write(f"| {c['shape']} | {c['ssim_abs_err']:.1e} | {c['total_abs_err']:.1e} | {c['grad_rel_err']:.1e} | {c['ssim_grad_rel_err']:.1e} |\n")
            f.write(f"\nidentical maps -> loss {d['identical_maps_loss']:.1e}; all cases ok = {d['ok']}.\n\nCost in the training step (config 3, B = 16): see `r02_training_step.md` -- "
                    "+0.54 ms per step with the first (121 taps per window) kernels, +0.14 ms with the separable shared-memory kernels.\n")
    print("wrote", sorted(x for x in os.listdir(P) if x.startswith("r02")))


if __name__ == "__main__":
    main()
