"""Per-layer micro-benchmark of the tcgen05 conv kernel variants (U-Net layer shapes at B=16).
    python tools/conv_bench.py            # table of TFLOP/s per layer x config
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mau_b200 import engine  # noqa: E402

LAYERS = [  # name, H, Cin, Cout   (B = 16)
    ("conv0_0.2", 250, 64, 64), ("conv0_1.1", 250, 192, 64), ("conv1_0.2", 125, 128, 128), ("conv1_1.1", 125, 384, 128),
    ("conv2_0.2", 62, 256, 256), ("conv2_1.1", 62, 768, 256), ("conv3_0.2", 31, 512, 512), ("conv3_1.1", 31, 1536, 512),
    ("conv4_0.2", 15, 1024, 1024), ("dgrad0_1.1", 250, 64, 192),
]
CFGS = {64: ["64,4,2", "64,2,2"], 128: ["128,2,2", "128,1,2", "256,1,2"], 256: ["256,1,2", "256,2,1", "128,2,2"],
        192: ["64,4,2", "256,1,2", "128,2,2"]}


def main():
    L = engine.lib()
    B = int(os.environ.get("B", "16"))
    only = sys.argv[1:]
    for name, H, Cin, Cout in LAYERS:
        if only and name not in only:
            continue
        x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
        w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
        y = torch.zeros(B, H, H, Cout, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * 9 * Cin * Cout * H * H * B
        res = []
        key = Cout if Cout in CFGS else 256
        for cfg in CFGS[key] + ["tap", "row3"]:
            impl = 0
            if cfg == "tap":
                impl = 1
            elif cfg == "row3":
                impl = 3
            else:
                os.environ["MAU_CONV_CFG"] = cfg
            ms = C.c_float()
            rc = L.mau_op_conv3x3_bench(impl, x.data_ptr(), B, H, H, Cin, Cin, w.data_ptr(), Cout, y.data_ptr(), Cout, 10, C.byref(ms))
            os.environ.pop("MAU_CONV_CFG", None)
            if rc:
                res.append(f"{cfg}: ERR {L.mau_last_error().decode()[:60]}")
            else:
                res.append(f"{cfg}: {ms.value*1e3:7.1f}us {flops/ms.value/1e9:6.0f}TF")
        print(f"{name:11s} {H:3d} {Cin:4d}->{Cout:4d} | " + " | ".join(res), flush=True)


def wgrad():
    """weight gradient: v2 auto / forced orientations vs the first-generation atomics kernel"""
    L = engine.lib()
    B = int(os.environ.get("B", "16"))
    only = [a for a in sys.argv[1:] if a != "wgrad"]
    for name, H, Cin, Cout in LAYERS + [("conv0_0.1", 250, 24, 64), ("conv1_0.1", 125, 64, 128), ("conv4_0.1", 15, 576, 1024)]:
        if name.startswith("dgrad") or (only and name not in only):
            continue
        x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
        dy = torch.randn(B, H, H, Cout, device="cuda").bfloat16()
        dw = torch.zeros(Cout, Cin, 3, 3, device="cuda")
        flops = 2.0 * 9 * Cin * Cout * H * H * B
        res = []
        for label, impl in (("v2", 0), ("v2 M=co", 4), ("v2 M=ci", 5), ("v1", 1)):
            ms = C.c_float()
            rc = L.mau_op_conv3x3_wgrad_bench(impl, x.data_ptr(), dy.data_ptr(), B, H, H, Cin, Cin, Cout, Cout, dw.data_ptr(), 10, C.byref(ms))
            if rc:
                res.append(f"{label}: ERR {L.mau_last_error().decode()[:60]}")
            else:
                res.append(f"{label}: {ms.value*1e3:7.1f}us {flops/ms.value/1e9:6.0f}TF")
        print(f"wgrad {name:11s} {H:3d} {Cin:4d}->{Cout:4d} | " + " | ".join(res), flush=True)


if __name__ == "__main__":
    if "wgrad" in sys.argv[1:]:
        wgrad()
    else:
        main()
