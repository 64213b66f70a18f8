#!/usr/bin/env python
"""Upper bound for a DFT-as-GEMM (tensor-core) formulation of the config-A frame transform.

BASELINE.json's north_star allows tensor cores "only if ... a DFT-as-GEMM variant actually beats the
CUDA-core path".  This probe times the GEMM such a variant would have to run, with the library's
best kernel (torch.matmul -> cuBLASLt, tcgen05 on sm_100), so the number is an UPPER bound for any
hand-written version of the same contraction; nothing here is on the product path.

    frames x [401 -> 416 samples incl. pre-emphasis halo]  @  [416 x 512]  (re bins 0..256, im bins 1..255)

fp32-equivalent accuracy needs split operands: fp16 pieces (11 bits each) give
hi*hi + hi*lo + lo*hi = 3 products (4 with lo*lo), bf16 pieces (8 bits) need 6.  The probe reports
frames/s for 1, 3, 4 and 6 products of the measured GEMM time and prints JSON lines.
"""
import json
import sys

import torch


def time_mm(a, b, iters=20):
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda", 0)
    print(json.dumps({"device": torch.cuda.get_device_name(0), "frames": frames}))
    for name, dt, k, n in (("fp16 full DFT 416x512", torch.float16, 416, 512),
                           ("bf16 full DFT 416x512", torch.bfloat16, 416, 512),
                           ("tf32 full DFT 416x512", torch.float32, 416, 512)):
        if dt == torch.float32:
            torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(frames, k, device=dev, dtype=dt)
        b = torch.randn(k, n, device=dev, dtype=dt)
        t = time_mm(a, b)
        tf = 2.0 * frames * k * n / t / 1e12
        row = {"gemm": name, "ms": t * 1e3, "tflops": tf}
        for prod in (1, 3, 4, 6):
            row[f"frames_per_s_{prod}_products"] = frames / (t * prod)
        print(json.dumps(row), flush=True)
    # the two-stage (32 x 16) factorisation as library GEMMs: K = N = 32, operands streamed from HBM
    for name, rows in (("fp16 stage 1: (frames*16) x 32 @ 32x32", 16), ("fp16 stage 2: (frames*17) x 32 @ 32x32", 17)):
        a = torch.randn(frames * rows, 32, device=dev, dtype=torch.float16)
        b = torch.randn(32, 32, device=dev, dtype=torch.float16)
        t = time_mm(a, b)
        print(json.dumps({"gemm": name, "ms": t * 1e3, "tflops": 2.0 * frames * rows * 32 * 32 / t / 1e12,
                          "frames_per_s_3_products": frames / (3 * t),
                          "note": "HBM-bound in a library; on chip the operand would have to sit in TMEM (DESIGN.md §5.4)"}),
              flush=True)


if __name__ == "__main__":
    main()
