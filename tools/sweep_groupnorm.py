"""Knob sweep of the streamed GroupNorm (SDEO_GN_F16_BUFS / _TILE_KB / _LAG / _HINTS) on one shape:
python tools/sweep_groupnorm.py N C H W [--iters K]. Graph-captured launches over rotating inputs, as tools/bench_groupnorm.py."""
import argparse
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablediffusioneo_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shape", type=int, nargs=4)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--bufs", default="4")
ap.add_argument("--tile-kb", default="1024")
ap.add_argument("--lag", default="-1")
ap.add_argument("--hints", default="-1")
ap.add_argument("--groups", default="2")
args = ap.parse_args()
n, c, h, w = args.shape
dev = torch.device("cuda:0")
os.environ["SDEO_GN_F16_VARIANT"] = "stream"  # the streamed kernel is opt-in since round 2 (the default beyond a cluster is two launches)
x = (torch.randn((n, h, w, c), device=dev) * 1.5).half()
gamma, beta = torch.rand((c,), device=dev) + 0.5, torch.randn((c,), device=dev) * 0.1
nbytes = x.numel() * 4
k = max(1, min(64, -(-(512 << 20) // nbytes)))
copies = [x] + [x.clone() for _ in range(k - 1)]
for bufs, kb, lag, hints, grps in itertools.product(args.bufs.split(","), args.tile_kb.split(","), args.lag.split(","), args.hints.split(","),
                                                   args.groups.split(",")):
    os.environ["SDEO_GN_F16_BUFS"], os.environ["SDEO_GN_F16_TILE_KB"], os.environ["SDEO_GN_F16_GROUPS"] = bufs, kb, grps
    for key, v in (("SDEO_GN_F16_LAG", lag), ("SDEO_GN_F16_HINTS", hints)):
        os.environ.pop(key, None)
        if v != "-1":
            os.environ[key] = v
    outs = [ops.groupnorm_f16(cp, gamma, beta, 1e-5, True) for cp in copies]
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = [ops.groupnorm_f16(cp, gamma, beta, 1e-5, True) for cp in copies]
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / k)
    t = sum(ts) / len(ts)
    print(f"groups {grps} bufs {bufs} tile_kb {kb:>4s} lag {lag:>5s} hints {hints:>2s}: {t:8.1f} us  {nbytes / t / 1e3:6.0f} GB/s", flush=True)
    del graph, outs
