"""Per-kernel GPU time of ONE replay of the captured step graph (CUPTI via torch.profiler), grouped by kernel and grid:
python tools/step_breakdown_cupti.py [--latent H W] [--batch B]. Unlike the library's block-0 trace this sees whole
multi-wave kernels (768x768 batch 4)."""
import argparse, collections, os, re, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import synth
from stablediffusioneo_b200.cldm.cldm import ControlLDM
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler
ap = argparse.ArgumentParser()
ap.add_argument("--latent", type=int, nargs=2, default=[96, 96])
ap.add_argument("--batch", type=int, default=4)
args = ap.parse_args()
h, w = args.latent
B = args.batch
dev = torch.device("cuda:0")
with torch.device(dev):
    model = ControlLDM().eval()
synth.randomize_(model)
x_T = torch.randn((B, 4, h, w), device=dev)
ctx = lambda s: torch.randn((B, 77, 768), generator=torch.Generator().manual_seed(s)).to(dev)
hint = (torch.rand((B, 1, 8 * h, 8 * w)) > 0.9).float().expand(-1, 3, -1, -1).contiguous().to(dev)
cond = {"c_concat": [hint], "c_crossattn": [ctx(1)]}
uncond = {"c_concat": [hint], "c_crossattn": [ctx(2)]}
sampler = DDIMSampler(model)
sampler.sample(4, B, (4, h, w), cond, verbose=False, eta=0.0, x_T=x_T, unconditional_guidance_scale=9.0,
               unconditional_conditioning=uncond)
torch.cuda.synchronize()
eng = sampler._engine
eng.reset_latent()
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = re.sub(r"^void (sdeo::)?", "", ev.name)
    m = re.match(r"(\w+)(<[^>]*>)?", name)
    key = (m.group(1), (m.group(2) or "")[:40]) if m else (name[:40], "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time
    t0 = ev.time_range.start
    t1 = ev.time_range.end
    t_min = t0 if t_min is None else min(t_min, t0)
    t_max = t1 if t_max is None else max(t_max, t1)
tot = sum(a[1] for a in agg.values())
print(f"latent {h}x{w} batch {B}: step span {(t_max - t_min) / 1e3:.2f} ms, sum of kernel times {tot / 1e3:.2f} ms")
kinds = collections.Counter()
for (k, t), a in agg.items():
    kinds[k] += a[1]
for k, v in kinds.most_common():
    print(f"  {k:28s} {v / 1e3:8.2f} ms {100 * v / tot:5.1f}%")
print("by instantiation:")
for (k, t), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {a[1] / 1e3:8.2f} ms {100 * a[1] / tot:5.1f}%  x{a[0]:3d}  {k}{t}")
