"""Developer tool: per-layer timing of smot_conv2d, warm (back-to-back) vs cold (L2 flushed before each launch)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from siammot_b200 import ops, _lib

dev = "cuda"
dt = torch.float16
LAYERS = [  # name, B, Cin, H, W, Cout, k, stride
    ("level2 3x3 64", 1, 64, 176, 320, 64, 3, 1),
    ("level3 3x3 128", 1, 128, 88, 160, 128, 3, 1),
    ("level4 3x3 256", 1, 256, 44, 80, 256, 3, 1),
    ("level5 3x3 512", 1, 512, 22, 40, 512, 3, 1),
    ("level5 root 1280", 1, 1280, 22, 40, 512, 1, 1),
    ("rpn conv P2", 1, 128, 176, 320, 128, 3, 1),
    ("fpn out P3", 1, 128, 88, 160, 128, 3, 1),
    ("fc6 300", 1, 6272, 1, 300, 1024, 1, 1),
    ("towers 30", 30, 128, 16, 16, 256, 3, 1),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
if os.environ.get("SMOT_LAYERS"):   # comma-separated substrings of layer names
    LAYERS = [l for l in LAYERS if any(k in l[0] for k in os.environ["SMOT_LAYERS"].split(","))]
for name, B, Cin, H, W, Cout, k, s in LAYERS:
    x = torch.randn(B, H, W, Cin, device=dev).to(dt)
    w = (torch.randn(Cout, k, k, Cin, device=dev) / math.sqrt(Cin * k * k)).to(dt)
    sc = torch.rand(Cout, device=dev) + 0.5
    bi = torch.randn(Cout, device=dev)
    out = torch.empty(B, H // s, W // s, Cout, device=dev, dtype=dt)
    for _ in range(3):
        ops.conv2d(x, w, sc, bi, None, s, k // 2, True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        ops.conv2d(x, w, sc, bi, None, s, k // 2, True, out=out)
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / n * 1e3
    cold = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.conv2d(x, w, sc, bi, None, s, k // 2, True, out=out)
        b.record()
        torch.cuda.synchronize()
        cold.append(a.elapsed_time(b) * 1e3)
    cold.sort()
    flops = 2.0 * B * (H // s) * (W // s) * Cout * Cin * k * k
    print("%-18s warm %6.1f us (%6.1f TFLOP/s)   cold median %6.1f us" % (name, warm, flops / warm / 1e6, cold[len(cold) // 2]))
