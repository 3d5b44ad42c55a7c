"""Micro-benchmark of attention_kernel on the self- / cross-attention shapes of the BASELINE configs:
python tools/bench_attention.py [--only SUBSTR]. Prints us, TFLOP/s (4 N_q N_kv d per head) and the exponential rate
(N_q N_kv per head / time) next to the MUFU ceiling (16 ex2 / clk / SM). One eager launch per shape runs inside the NVTX
range "final" (ncu --nvtx --nvtx-include "final/")."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import ops  # noqa: E402

BF = torch.bfloat16
ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
# (name, batch (x2 for cond+uncond), heads, nq, nkv, d)
CASES = [
    ("self 768x768 b4 level1 N=9216 d=40", 8, 8, 9216, 9216, 40),
    ("self 768x768 b4 level2 N=2304 d=80", 8, 8, 2304, 2304, 80),
    ("self 768x768 b4 level3 N=576 d=160", 8, 8, 576, 576, 160),
    ("cross 768x768 b4 level1 Nq=9216 Nkv=77 d=40", 8, 8, 9216, 77, 40),
    ("self 512x512 b1 level1 N=4096 d=40", 2, 8, 4096, 4096, 40),
    ("self 256x384 b1 level1 N=1536 d=40", 2, 8, 1536, 1536, 40),
    ("self 256x384 b1 level2 N=384 d=80", 2, 8, 384, 384, 80),
    ("cross 256x384 b1 level1 Nq=1536 Nkv=77 d=40", 2, 8, 1536, 77, 40),
]
for name, b, heads, nq, nkv, d in CASES:
    if args.only and args.only not in name:
        continue
    ldv = (nkv + 7) // 8 * 8
    q = torch.randn((b * heads, nq, d), device=dev).to(BF)
    k = torch.randn((b * heads, nkv, d), device=dev).to(BF)
    vt = torch.randn((b * heads, d, ldv), device=dev).to(BF)
    run = lambda: ops.attention(q, k, vt, b, heads, nq, nkv, d, ldv, 1.0 / math.sqrt(d))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("final")
    run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(args.iters):
            run()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1000 / args.iters
    flops = 4.0 * b * heads * nq * nkv * d
    exps = float(b * heads) * nq * nkv
    mufu_floor_us = exps / (148 * 16 * 1.965e3)
    print(f"{name:46s} {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s   {exps / us / 1e3:7.1f} G exp/s  (MUFU floor {mufu_floor_us:7.1f} us = "
          f"{100 * mufu_floor_us / us:4.1f}% of the time)")
