"""Timeline of ONE replay of the captured denoising-step graph, from the library's own kernel trace (sdeo_set_trace):
python tools/step_timeline.py [--csv out.csv]. Per kernel: start of block 0, the moment its grid dependency resolved
(PDL), end of block 0 -- all on the GPU's globaltimer. Shows where the step's wall time goes: busy vs idle, per-kind
sums, the longest kernels, and how much prologue time PDL hides."""
import argparse
import collections
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import _lib, ops, synth  # noqa: E402
from stablediffusioneo_b200.cldm.cldm import ControlLDM  # noqa: E402
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--latent", type=int, nargs=2, default=[32, 48])
ap.add_argument("--csv", default="")
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
h, w = args.latent
dev = torch.device("cuda:0")
with torch.device(dev):
    model = ControlLDM().eval()
synth.randomize_(model)
B = args.batch
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
assert eng.graph is not None
from stablediffusioneo_b200 import trace  # noqa: E402

eng.reset_latent()
for _ in range(3):
    eng.step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def one_step():
    s.record()
    eng.step()
    e.record()


rec = trace.capture(one_step, dev)
summ = trace.summarize(rec)
n = len(rec)
print(f"{n} kernels traced; event time {s.elapsed_time(e) * 1e3:.1f} us; trace span {summ['span_us']:.1f} us")
print(f"union of [dependency resolved, block-0 end] intervals: {summ['busy_union_us']:.1f} us busy, "
      f"{summ['span_us'] - summ['busy_union_us']:.1f} us with no kernel past its dependency")
for k, a in sorted(summ["kinds"].items(), key=lambda kv: -kv[1]["busy_us"]):
    print(f"  {k:12s} {a['launches']:4d} kernels  sum(dep->end) {a['busy_us']:8.1f} us  avg {a['busy_us'] / a['launches']:6.2f}  "
          f"idle tail after them {a['tail_us']:7.1f} us  sum(start->dep, hidden by PDL) {a['hidden_prologue_us']:8.1f} us")
print("longest kernels (dep->end us):")
for r in sorted(rec, key=lambda r: -(r.end - r.dep))[:25]:
    print(f"  {r.kind:10s} grid {r.grid:5d} mode {r.mode} splits {r.splits} BN {r.bn:3d}  {r.end - r.dep:7.2f} us  (start {r.start:8.1f})")
print("by launch signature (kind grid mode splits BN halo): count, sum(dep->end) us, avg us")
agg = collections.OrderedDict()
for r in rec:
    a = agg.setdefault((r.kind, r.grid, r.mode, r.splits, r.bn, r.halo), [0, 0.0])
    a[0] += 1
    a[1] += r.end - r.dep
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[0]:10s} grid {k[1]:5d} mode {k[2]} splits {k[3]} BN {k[4]:3d} halo {k[5]}  x{a[0]:3d}  {a[1]:8.1f} us  avg {a[1] / a[0]:6.2f}")
hist = collections.Counter()
for r in rec:
    if r.kind == "conv":
        hist[min(int((r.end - r.dep) // 2) * 2, 40)] += 1
print("conv dep->end histogram (us bucket: count):", dict(sorted(hist.items())))
if args.csv:
    with open(args.csv, "w") as f:
        f.write("kind,grid,mode,splits,bn,start_us,dep_us,end_us\n")
        for r in sorted(rec, key=lambda r: r.start):
            f.write(f"{r.kind},{r.grid},{r.mode},{r.splits},{r.bn},{r.start:.2f},{r.dep:.2f},{r.end:.2f}\n")
