"""Experiment: does alternating between conv_gemm_kernel instantiations (each ~65 KB of SASS) cost instruction-cache
misses? Times graphs of 24 launches: one instantiation repeated vs. a round-robin over several."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import ops
BF = torch.bfloat16
dev = torch.device("cuda:0")
n, h, w, c = 2, 32, 48, 320
x = torch.randn((n, h, w, c), device=dev).to(BF)
wt = torch.randn((c, c, 1, 1), device=dev) / math.sqrt(c)
pw = ops.pack_conv_weight(wt)
bias = torch.randn((c,), device=dev)
res32 = torch.randn((n, h, w, c), device=dev)
res16 = res32.to(BF)
emb = torch.randn((n, c), device=dev)
variants = {
    "stream(f32 res, twin)": dict(bias=bias, residual=res32, out_fp32=True, twin=True),
    "plain bf16": dict(bias=bias),
    "bf16 res": dict(bias=bias, residual=res16),
    "f32 out": dict(bias=bias, out_fp32=True),
    "f32 out + emb": dict(bias=bias, emb=emb, out_fp32=True),
    "f32 res": dict(bias=bias, residual=res32, out_fp32=True),
    "gn stats": dict(bias=bias, residual=res32, out_fp32=True, twin=True, gn_stats=True),
    "silu": dict(bias=bias, act=ops.SDEO_ACT_SILU) if hasattr(ops, "SDEO_ACT_SILU") else dict(bias=bias),
}
os.environ["SDEO_PAIR"] = "0"
os.environ["SDEO_HALO"] = "0"
os.environ["SDEO_FORCE_BN"] = "64"
def time_graph(seq, reps=24):
    for kw in seq:
        ops.conv2d(x, pw, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            ops.conv2d(x, pw, **seq[i % len(seq)])
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1000 / reps)
    return best
single = {}
for name, kw in variants.items():
    single[name] = time_graph([kw])
    print(f"{name:24s} alone: {single[name]:6.2f} us per launch")
names = list(variants)
for k in (2, 4, len(names)):
    seq = [variants[nm] for nm in names[:k]]
    t = time_graph(seq)
    exp = sum(single[nm] for nm in names[:k]) / k
    print(f"round-robin over {k} instantiations: {t:6.2f} us per launch (mean of the singles {exp:6.2f})")
