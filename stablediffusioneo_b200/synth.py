"""Device-side synthetic weights for benchmarks: random-init of the SD1.5 architecture without touching the host.
Same distribution family as the test oracle's scheme: default-init variance 1/(3 fan_in) for regular
weights, N(0, 1/fan_in) for the tensors the reference zero-initialises (so no layer is a no-op), small biases."""
import math

import torch

_ZERO_INIT_SUFFIXES = ("out_layers.3.weight", "proj_out.weight", "out.2.weight", "input_hint_block.14.weight")


def _is_zero_init(name):
    return name.endswith(_ZERO_INIT_SUFFIXES) or ".zero_convs." in name or ".middle_block_out." in name


@torch.no_grad()
def randomize_(model, seed=1234):
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() > 1:
            fan_in = p[0].numel()
            vae = name.startswith("first_stage_model.")
            var = 1.0 / fan_in if (vae or _is_zero_init(name)) else 1.0 / (3.0 * fan_in)
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * math.sqrt(var))
        elif name.endswith("bias"):
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.1)
        else:  # normalisation scale
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device))
    return model
