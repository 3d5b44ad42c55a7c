"""Shared test helpers: build our modules with the oracle's deterministic weights (tests only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import sd15_oracle as O  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def load_golden(tag):
    return torch.load(os.path.join(GOLDEN, f"{tag}.pt"), map_location="cpu", weights_only=False)


def oracle_weights(cfg, vcfg):
    sd_unet = O.make_weights(O.unet_param_spec(cfg), seed=1234, prefix="unet.")
    sd_cn = O.make_weights(O.controlnet_param_spec(cfg), seed=1234, prefix="control.")
    sd_vae = O.make_weights(O.vae_param_spec(vcfg), seed=1234, prefix="vae.")
    return sd_unet, sd_cn, sd_vae


def unet_kwargs(cfg):
    return dict(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                model_channels=cfg.model_channels, num_res_blocks=cfg.num_res_blocks,
                attention_resolutions=list(cfg.attention_resolutions), channel_mult=list(cfg.channel_mult),
                num_heads=cfg.num_heads, use_spatial_transformer=True, transformer_depth=cfg.transformer_depth,
                context_dim=cfg.context_dim, use_checkpoint=False, legacy=False)


def vae_kwargs(vcfg):
    return dict(double_z=True, z_channels=vcfg.z_channels, resolution=256, in_channels=3, out_ch=vcfg.out_ch,
                ch=vcfg.ch, ch_mult=list(vcfg.ch_mult), num_res_blocks=vcfg.num_res_blocks, attn_resolutions=[],
                dropout=0.0)


def build_control_ldm(cfg, vcfg, device, weights=None):
    """Our ControlLDM on `device`, loaded (strict) with the oracle's weights under the SD checkpoint prefixes.
    The modules are constructed on the `meta` device and take the oracle's CPU tensors by assignment, then move to the
    GPU as plain host->device copies: no torch init kernels (uniform_ / fill_) are launched, so a profiler's launch list
    of a test or of smoke() starts with the library's own kernels."""
    from stablediffusioneo_b200.cldm.cldm import ControlLDM
    sd_unet, sd_cn, sd_vae = weights if weights is not None else oracle_weights(cfg, vcfg)
    with torch.device("meta"):
        model = ControlLDM(unet_config=unet_kwargs(cfg), first_stage_config=vae_kwargs(vcfg)).eval()
    sd = {}
    sd.update({"model.diffusion_model." + k: v for k, v in sd_unet.items()})
    sd.update({"control_model." + k: v for k, v in sd_cn.items()})
    sd.update({"first_stage_model." + k: v for k, v in sd_vae.items()})
    missing, unexpected = model.load_state_dict(sd, strict=True, assign=True)
    assert not missing and not unexpected
    model.register_schedule(device="cpu")   # the (non-persistent) schedule buffers were created on meta
    model = model.to(device)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def inputs_on(cfg, h, w, device, hint=None):
    x_T, cond, uncond = O.make_inputs(cfg, 1, h, w, hint=hint)
    mv = lambda c: {"c_concat": [t.to(device) for t in c["c_concat"]], "c_crossattn": [t.to(device) for t in c["c_crossattn"]]}
    return x_T.to(device), mv(cond), mv(uncond)


def canny_hint():
    """The golden's hint: cv2.Canny of pictures_croped/bird_0.jpg, committed bit-packed as canny_bird0.npy."""
    import numpy as np
    bits = np.unpackbits(np.load(os.path.join(GOLDEN, "canny_bird0.npy")))[: 256 * 384].reshape(256, 384)
    e = torch.from_numpy(bits.astype("float32"))
    return e[None, None].expand(1, 3, 256, 384).contiguous()
