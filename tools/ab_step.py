"""Device-time of the captured denoising step (256x384, cond+uncond), mean of N graph replays:
python tools/ab_step.py [--reps N]. With SDEO_LIB=<other build of libsdeo.so> it times that build (A/B on one box)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import synth  # noqa: E402
from stablediffusioneo_b200.cldm.cldm import ControlLDM  # noqa: E402
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--latent", type=int, nargs=2, default=[32, 48])
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
eng = sampler._engine
best = []
for _ in range(3):
    eng.reset_latent()
    for _ in range(5):
        eng.step()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for i in range(args.reps):
        if i % 4 == 0:
            eng.reset_latent()
        eng.step()
    e.record()
    torch.cuda.synchronize()
    best.append(s.elapsed_time(e) / args.reps)
print(f"{os.environ.get('SDEO_LIB', 'in-tree build')}: ms/step {min(best):.4f} (runs: {' '.join(f'{b:.4f}' for b in best)})")
