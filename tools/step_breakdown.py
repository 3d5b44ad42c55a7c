"""Per-op / per-shape time breakdown of ONE denoising step (256x384, cond+uncond batch 2), eager launches on a single
back-logged stream with CUDA events around every libsdeo call: python tools/step_breakdown.py [--latent H W] [--top N].
Kernel durations only (a spin kernel keeps the GPU busy while the host enqueues); the graph replay overlaps two streams
and hides launch gaps, so the SUM here is an upper bound of the step, the SHARES are what matter."""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import ops, synth  # noqa: E402
from stablediffusioneo_b200.cldm.cldm import ControlLDM  # noqa: E402
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--latent", type=int, nargs=2, default=[32, 48])
ap.add_argument("--top", type=int, default=60)
args = ap.parse_args()
h, w = args.latent
dev = torch.device("cuda:0")
with torch.device(dev):
    model = ControlLDM().eval()
synth.randomize_(model)
x_T = torch.randn((1, 4, h, w), device=dev)
ctx = lambda s: torch.randn((1, 77, 768), generator=torch.Generator().manual_seed(s)).to(dev)
hint = (torch.rand((1, 1, 8 * h, 8 * w)) > 0.9).float().expand(-1, 3, -1, -1).contiguous().to(dev)
cond = {"c_concat": [hint], "c_crossattn": [ctx(1)]}
uncond = {"c_concat": [hint], "c_crossattn": [ctx(2)]}
sampler = DDIMSampler(model)
sampler.use_cuda_graph = False
sampler.sample(2, 1, (4, h, w), cond, verbose=False, eta=0.0, x_T=x_T, unconditional_guidance_scale=9.0,
               unconditional_conditioning=uncond)
torch.cuda.synchronize()
eng = sampler._engine
eng.side_stream = torch.cuda.current_stream()  # one stream: durations are not inflated by a concurrent branch

rec = []


def wrap(name, keyfn):
    orig = getattr(ops, name)

    def timed(*a, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig(*a, **kw)
        e.record()
        rec.append((name, keyfn(*a, **kw), s, e))
        return y

    setattr(ops, name, timed)
    return orig


def conv_key(x, pw, x2=None, bias=None, emb=None, residual=None, scale=1.0, act=0, stride=1, out=None, out_fp32=False,
             epi_mode=0, qkv=None, twin=False, emb_step=None):
    n, hh, ww, _ = x.shape
    k = pw.ksize
    ho, wo = (hh + 2 * (k // 2) - k) // stride + 1, (ww + 2 * (k // 2) - k) // stride + 1
    m = n * ho * wo
    flops = 2.0 * m * pw.cout * (pw.c1 + pw.c2) * k * k
    tag = {0: "", 1: " geglu", 2: " qkv"}[epi_mode] + (" s2" if stride == 2 else "") + (" +res" if residual is not None else "") \
        + (" f32" if out_fp32 else "") + (" twin" if twin else "") + (" emb" if emb is not None else "")
    return (f"k{k} {pw.c1}+{pw.c2}->{pw.cout} M={m}{tag}", flops, pw.data.numel() * 2)


def gn_key(x, gamma, beta, eps, silu, x2=None, groups=32, out=None):
    c = x.shape[3] + (x2.shape[3] if x2 is not None else 0)
    nbytes = x.numel() * x.element_size() + (x2.numel() * x2.element_size() if x2 is not None else 0)
    rows = x.shape[0] * x.shape[1] * x.shape[2]
    return (f"C={c} rows={rows} {'f32' if x.dtype == torch.float32 else 'bf16'}", 0.0, nbytes + rows * c * 2)


def ln_key(x, gamma, beta, eps=1e-5):
    return (f"C={x.shape[-1]} rows={x.numel() // x.shape[-1]}", 0.0, x.numel() * (x.element_size() + 2))


def att_key(q, k, vt, batch, heads, nq, nkv, d, ldv, scale, out=None):
    return (f"B={batch} h={heads} nq={nq} nkv={nkv} d={d}", 4.0 * batch * heads * nq * nkv * d, 0)


generic = lambda *a, **kw: ("", 0.0, 0)
wrap("conv2d", conv_key)
wrap("groupnorm", gn_key)
wrap("layernorm", ln_key)
wrap("attention", att_key)
for nm in ("upsample_nearest2x", "silu", "timestep_embedding", "cfg_ddim_step", "counter_add", "add_scaled", "to_bf16",
           "to_f32", "nchw_to_nhwc", "nhwc_to_nchw"):
    wrap(nm, generic)

eng.reset_latent()
torch.cuda.synchronize()
rec.clear()
torch.cuda._sleep(int(2.5e9))
eng._step()
torch.cuda.synchronize()

agg = collections.OrderedDict()
for name, (key, flops, nbytes), s, e in rec:
    k = (name, key)
    a = agg.setdefault(k, [0, 0.0, flops, nbytes, []])
    a[0] += 1
    a[4].append(s.elapsed_time(e) * 1000.0)
for a in agg.values():
    med = sorted(a[4])[len(a[4]) // 2]
    a[1] = med * a[0]  # median x count: robust against the odd host-side stall inside an event pair
tot = sum(a[1] for a in agg.values())
print(f"{len(rec)} calls, sum of kernel times {tot:.1f} us")
by_op = collections.Counter()
cnt_op = collections.Counter()
for (name, key), a in agg.items():
    by_op[name] += a[1]
    cnt_op[name] += a[0]
for name, t in by_op.most_common():
    print(f"  {name:20s} {cnt_op[name]:4d} calls {t:9.1f} us {100 * t / tot:5.1f}%  avg {t / cnt_op[name]:6.1f} us")
print(f"{'op':10s} {'shape':58s} {'n':>3s} {'total us':>9s} {'avg us':>7s} {'%':>5s} {'TFLOP/s':>8s} {'GB/s':>7s}")
for (name, key), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
    avg = a[1] / a[0]
    tf = a[2] / avg / 1e6 if a[2] else 0.0
    gb = a[3] / avg / 1e3 if a[3] else 0.0
    print(f"{name:10s} {key:58s} {a[0]:3d} {a[1]:9.1f} {avg:7.1f} {100 * a[1] / tot:5.1f} {tf:8.1f} {gb:7.1f}")
