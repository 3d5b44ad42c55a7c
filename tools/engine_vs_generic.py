"""Engine path (batched cond+uncond, tables, fused statistics, CUDA graph) against the generic sampler path (two
apply_model calls per step through the public modules) at a given latent size: python tools/engine_vs_generic.py H W [S].
Both are this library; the generic path is the one the per-step parity tests pin to the oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import synth  # noqa: E402
from stablediffusioneo_b200.cldm.cldm import ControlLDM  # noqa: E402
from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler  # noqa: E402

h, w = int(sys.argv[1]), int(sys.argv[2])
S = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
with torch.device(dev):
    model = ControlLDM().eval()
synth.randomize_(model)
x_T = torch.randn((1, 4, h, w), generator=torch.Generator().manual_seed(5)).to(dev)
ctx = lambda s: torch.randn((1, 77, 768), generator=torch.Generator().manual_seed(s)).to(dev)
hint = (torch.rand((1, 1, 8 * h, 8 * w), generator=torch.Generator().manual_seed(9)) > 0.9).float().expand(-1, 3, -1, -1).contiguous().to(dev)
cond = {"c_concat": [hint], "c_crossattn": [ctx(1)]}
uncond = {"c_concat": [hint], "c_crossattn": [ctx(2)]}
outs = {}
for name, use_engine in (("engine", True), ("generic", False)):
    sampler = DDIMSampler(model)
    sampler.use_engine = use_engine
    outs[name], _ = sampler.sample(S, 1, (4, h, w), cond, verbose=False, eta=0.0, x_T=x_T, unconditional_guidance_scale=9.0,
                                   unconditional_conditioning=uncond)
a, b = outs["engine"], outs["generic"]
print(f"latent {h}x{w}, {S} steps: engine vs generic rel L2 {float((a - b).norm() / b.norm()):.3e}, finite {bool(torch.isfinite(a).all())}")
