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
sampler.sample(4, 1, (4, h, w), cond, verbose=False, eta=0.0, x_T=x_T, unconditional_guidance_scale=9.0,
               unconditional_conditioning=uncond)
torch.cuda.synchronize()
eng = sampler._engine
assert eng.graph is not None
cap = 4096
buf = torch.zeros(4 + 4 * cap, dtype=torch.int64, device=dev)
buf[1] = cap
lib = _lib.load()
eng.reset_latent()
for _ in range(3):
    eng.step()
torch.cuda.synchronize()
assert lib.sdeo_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
eng.step()
e.record()
torch.cuda.synchronize()
lib.sdeo_set_trace(None)
b = buf.cpu()
n = int(b[0])
rec = b[4:4 + 4 * n].reshape(n, 4)
tag = rec[:, 0] & 0xFF
grid = (rec[:, 0] >> 8) & 0xFFFFFFFF
mode = (rec[:, 0] >> 40) & 0xF
splits = (rec[:, 0] >> 44) & 0xF
bn = (rec[:, 0] >> 48) & 0xFFF
t0 = int(rec[:, 1].min())
ts, td, te = (rec[:, 1] - t0).double() / 1e3, (rec[:, 2] - t0).double() / 1e3, (rec[:, 3] - t0).double() / 1e3
names = {1: "conv", 2: "attention", 3: "groupnorm", 4: "layernorm", 5: "elementwise"}
print(f"{n} kernels traced; event time {s.elapsed_time(e) * 1e3:.1f} us; trace span {float(te.max()):.1f} us")
# busy union of [dep, end]
iv = sorted((float(a), float(c)) for a, c in zip(td, te))
busy, cur_s, cur_e = 0.0, None, None
for a, c in iv:
    if cur_e is None or a > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = a, c
    else:
        cur_e = max(cur_e, c)
busy += cur_e - cur_s
print(f"union of [dependency resolved, block-0 end] intervals: {busy:.1f} us busy, {float(te.max()) - busy:.1f} us with no kernel past its dependency")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i in range(n):
    a = agg[names.get(int(tag[i]), "?")]
    a[0] += 1
    a[1] += float(te[i] - td[i])
    a[2] += float(td[i] - ts[i])
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:12s} {a[0]:4d} kernels  sum(dep->end) {a[1]:8.1f} us  avg {a[1] / a[0]:6.2f}  sum(start->dep, hidden by PDL) {a[2]:8.1f} us")
order = sorted(range(n), key=lambda i: -float(te[i] - td[i]))
print("longest kernels (dep->end us):")
for i in order[:25]:
    print(f"  {names.get(int(tag[i]), '?'):10s} grid {int(grid[i]):5d} mode {int(mode[i])} splits {int(splits[i])} BN {int(bn[i]):3d}  "
          f"{float(te[i] - td[i]):7.2f} us  (start {float(ts[i]):8.1f})")
hist = collections.Counter()
for i in range(n):
    if int(tag[i]) == 1:
        hist[min(int(float(te[i] - td[i]) // 2) * 2, 40)] += 1
print("conv dep->end histogram (us bucket: count):", dict(sorted(hist.items())))
if args.csv:
    with open(args.csv, "w") as f:
        f.write("idx,kind,grid,mode,splits,bn,start_us,dep_us,end_us\n")
        for i in sorted(range(n), key=lambda i: float(ts[i])):
            f.write(f"{i},{names.get(int(tag[i]), '?')},{int(grid[i])},{int(mode[i])},{int(splits[i])},{int(bn[i])},"
                    f"{float(ts[i]):.2f},{float(td[i]):.2f},{float(te[i]):.2f}\n")
