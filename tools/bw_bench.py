"""Achieved HBM bandwidth of the bandwidth-bound kernels of the path, one kernel at a time.

    python tools/bw_bench.py [--json out.json]

Each kernel runs at the shapes it sees in the U-Net at B=16, 250x250 tiles; bytes are the
ALGORITHMIC bytes of the op (every tensor element read or written once), time is CUDA events around
`iters` launches that rotate over enough copies of the tensors to exceed the 126 MB L2.  Peak =
MEASURED_PEAKS.json hbm_gbs (else the B200_PROFILING.md fallback 6650).
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
from mau_b200 import engine  # noqa: E402

KINDS = {  # name: (kind, bytes(B,H,W,C,es) -> algorithmic bytes)
    "bn_stats": (0, lambda B, H, W, Cc, es: B * H * W * Cc * es),
    "bn_apply_relu": (1, lambda B, H, W, Cc, es: 2 * B * H * W * Cc * es),
    "bn_bwd_reduce": (2, lambda B, H, W, Cc, es: 2 * B * H * W * Cc * es),
    "bn_bwd_apply": (3, lambda B, H, W, Cc, es: 3 * B * H * W * Cc * es),
    "maxpool": (4, lambda B, H, W, Cc, es: (B * H * W + B * (H // 2) * (W // 2)) * Cc * es),
    "maxpool_bwd": (5, lambda B, H, W, Cc, es: (2 * B * H * W + B * (H // 2) * (W // 2)) * Cc * es),
    "bilinear": (6, lambda B, H, W, Cc, es: (B * H * W + B * (H // 2) * (W // 2)) * Cc * es),
    "bilinear_bwd": (7, lambda B, H, W, Cc, es: (B * H * W + B * (H // 2) * (W // 2)) * Cc * es),
    "head": (8, lambda B, H, W, Cc, es: B * H * W * (Cc * es + 2 * 4)),
    "head_bwd": (9, lambda B, H, W, Cc, es: B * H * W * (2 * Cc * es + 3 * 4)),
    "nchw_to_nhwc": (10, lambda B, H, W, Cc, es: B * H * W * (23 * 4 + 24 * es)),
    "embed_broadcast": (11, lambda B, H, W, Cc, es: B * H * W * Cc * es),
    "loss_l1_grad": (12, lambda B, H, W, Cc, es: B * H * W * 2 * 4 * 3),
    "slice_copy": (13, lambda B, H, W, Cc, es: 2 * B * H * W * Cc * es),
}
# (kernel, H=W, C): where each kernel is heaviest in the U-Net (B = 16)
CASES = [
    ("slice_copy", 250, 64), ("bn_stats", 250, 64), ("bn_apply_relu", 250, 64), ("bn_bwd_reduce", 250, 64),
    ("bn_bwd_apply", 250, 64), ("bn_stats", 125, 128), ("bn_bwd_reduce", 125, 128), ("bn_bwd_apply", 125, 128),
    ("bn_stats", 62, 256), ("bn_bwd_apply", 62, 256), ("bn_stats", 31, 512), ("bn_bwd_apply", 31, 512),
    ("maxpool", 250, 64), ("maxpool_bwd", 250, 64), ("bilinear", 250, 128), ("bilinear_bwd", 250, 128),
    ("bilinear", 124, 256), ("bilinear_bwd", 124, 256), ("head", 250, 64), ("head_bwd", 250, 64),
    ("nchw_to_nhwc", 250, 24), ("embed_broadcast", 250, 128), ("loss_l1_grad", 250, 2),
]


def main():
    L = engine.lib()
    B = int(os.environ.get("B", "16"))
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    rows = []
    iters = int(os.environ.get("BW_ITERS", "24"))
    only = os.environ.get("BW_ONLY", "")
    only = set(only.split(",")) if only else None
    for name, H, Cc in CASES:
        if only and name not in only:
            continue
        kind, fbytes = KINDS[name]
        by = fbytes(B, H, H, Cc, 2)
        sets = max(2, int(300e6 // max(by, 1)) + 1)
        sets = min(sets, 12)
        sets = int(os.environ.get("BW_SETS", sets))
        ms = C.c_float()
        rc = L.mau_op_bw_bench(kind, 0, B, H, H, Cc, iters, sets, C.byref(ms))
        if rc:
            print(f"{name:16s} {H:4d} C={Cc:4d}  ERR {L.mau_last_error().decode()[:80]}", flush=True)
            continue
        gbs = by / (ms.value * 1e-3) / 1e9
        rows.append({"kernel": name, "B": B, "H": H, "C": Cc, "bytes": by, "us": ms.value * 1e3, "GB/s": gbs,
                     "frac_of_hbm_peak": gbs / peak})
        print(f"{name:16s} {H:4d} C={Cc:4d}  {by/1e6:8.1f} MB  {ms.value*1e3:8.1f} us  {gbs:7.0f} GB/s  {100*gbs/peak:5.1f} % of {peak:.0f}",
              flush=True)
    if "--json" in sys.argv:
        json.dump({"peak_gbs": peak, "rows": rows}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
