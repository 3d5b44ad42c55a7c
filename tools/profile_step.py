"""Runs the 256x384 workload (cond+uncond batch 2) so that ncu can list / capture the kernels of one denoising step:
python tools/profile_step.py [--latent H W] [--steps N] [--graph]. After the warm-up sample (which also autotunes the conv
shapes) ONE more step runs inside the NVTX range "sdeo_step":
  ncu --nvtx --nvtx-include "sdeo_step/" --graph-profiling node --metrics gpu__time_duration.sum ... python tools/profile_step.py --graph
profiles exactly the kernel nodes of one replay of the captured step graph."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import ops, synth  # noqa: E402
from stablediffusioneo_b200.cldm.cldm import ControlLDM  # noqa: E402
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--latent", type=int, nargs=2, default=[32, 48])
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--graph", action="store_true")
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
sampler.use_cuda_graph = args.graph
n0 = ops.LAUNCHES
samples, _ = sampler.sample(args.steps, 1, (4, h, w), cond, verbose=False, eta=0.0, x_T=x_T,
                            unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
torch.cuda.synchronize()
eng = sampler._engine
eng.reset_latent()
eng.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("sdeo_step")
eng.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("launches", ops.LAUNCHES - n0, "finite", bool(torch.isfinite(samples).all()))
