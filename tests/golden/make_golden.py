"""Generates tests/golden/*.pt by running the REAL reference modules from /root/reference (this container only —
the reference does not travel to the GPU box) on the oracle's deterministic weights and inputs.

    python tests/golden/make_golden.py [--skip-full]

The reference lacks ldm/models/* and omegaconf; they are satisfied by in-process sys.modules
shims, and ControlLDM.apply_model's 14 lines (cldm/cldm.py:328-341) are restated around the real ControlNet /
ControlledUnetModel objects. DDIMSampler (cldm/ddim_hacked.py) runs unmodified against a duck-typed model object,
with register_buffer patched only because this container has no CUDA device (ddim_hacked.py:17-21 hard-codes cuda).
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SDEO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import sd15_oracle as O  # noqa: E402


def install_shims():
    lc = types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass

    lc.ListConfig = ListConfig
    om = types.ModuleType("omegaconf")
    om.listconfig = lc
    sys.modules.setdefault("omegaconf", om)
    sys.modules.setdefault("omegaconf.listconfig", lc)
    for name in ("ldm.models", "ldm.models.diffusion"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    ddpm = types.ModuleType("ldm.models.diffusion.ddpm")

    class LatentDiffusion(torch.nn.Module):
        pass

    ddpm.LatentDiffusion = LatentDiffusion
    ddim = types.ModuleType("ldm.models.diffusion.ddim")

    class DDIMSampler:
        pass

    ddim.DDIMSampler = DDIMSampler
    sys.modules["ldm.models.diffusion.ddpm"] = ddpm
    sys.modules["ldm.models.diffusion.ddim"] = ddim


def quiet():
    """The reference prints 5 lines per attention call (attention.py:208-224)."""
    return contextlib.redirect_stdout(io.StringIO())


def build_reference(cfg):
    from cldm.cldm import ControlledUnetModel, ControlNet
    common = dict(image_size=32, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
                  num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions),
                  channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads, use_spatial_transformer=True,
                  transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim, use_checkpoint=False,
                  legacy=False)
    with quiet():
        unet = ControlledUnetModel(out_channels=cfg.out_channels, **common).eval()
        cn = ControlNet(hint_channels=cfg.hint_channels, **common).eval()
    return unet, cn


def build_reference_vae(vcfg):
    from ldm.modules.diffusionmodules.model import Decoder
    with quiet():
        dec = Decoder(ch=vcfg.ch, out_ch=vcfg.out_ch, ch_mult=vcfg.ch_mult, num_res_blocks=vcfg.num_res_blocks,
                      attn_resolutions=[], dropout=0.0, in_channels=3, resolution=256, z_channels=vcfg.z_channels).eval()
    pq = torch.nn.Conv2d(vcfg.z_channels, vcfg.z_channels, 1)
    return dec, pq


class RefModel:
    """Duck-typed stand-in for ControlLDM: what DDIMSampler needs (ddim_hacked.py:14,26-32,130,188-197)."""

    def __init__(self, unet, cn):
        self.unet, self.cn = unet, cn
        self.num_timesteps = 1000
        self.parameterization = "eps"
        self.device = torch.device("cpu")
        betas = O.make_beta_schedule()
        ac = np.cumprod(1.0 - betas, axis=0)
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(ac, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32)
        self.control_scales = [1.0] * 13
        self.only_mid_control = False
        self.calls = []

    def apply_model(self, x_noisy, t, cond):  # cldm/cldm.py:328-341
        cond_txt = torch.cat(cond["c_crossattn"], 1)
        with quiet():
            if cond["c_concat"] is None:
                eps = self.unet(x=x_noisy, timesteps=t, context=cond_txt, control=None,
                                only_mid_control=self.only_mid_control)
            else:
                control = self.cn(x=x_noisy, hint=torch.cat(cond["c_concat"], 1), timesteps=t, context=cond_txt)
                control = [c * scale for c, scale in zip(control, self.control_scales)]
                eps = self.unet(x=x_noisy, timesteps=t, context=cond_txt, control=control,
                                only_mid_control=self.only_mid_control)
        self.calls.append((int(t[0]), eps.clone()))
        return eps


def run_reference_sampler(model, x_T, cond, uncond, S, scale):
    from cldm.ddim_hacked import DDIMSampler
    sampler = DDIMSampler(model)
    sampler.register_buffer = lambda name, attr: setattr(sampler, name, attr)  # no CUDA device here
    with quiet():
        samples, inter = sampler.sample(S, x_T.shape[0], tuple(x_T.shape[1:]), cond, verbose=False, eta=0.0, x_T=x_T,
                                        unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
    return samples, sampler


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def golden_for(cfg, vcfg, tag, latent_hw, S, with_sampler=True, hint=None):
    h, w = latent_hw
    t0 = time.time()
    sd_unet = O.make_weights(O.unet_param_spec(cfg), seed=1234, prefix="unet.")
    sd_cn = O.make_weights(O.controlnet_param_spec(cfg), seed=1234, prefix="control.")
    sd_vae = O.make_weights(O.vae_param_spec(vcfg), seed=1234, prefix="vae.")
    print(f"[{tag}] weights: {time.time() - t0:.1f}s; unet {sum(v.numel() for v in sd_unet.values()) / 1e6:.2f} M, "
          f"controlnet {sum(v.numel() for v in sd_cn.values()) / 1e6:.2f} M, vae {sum(v.numel() for v in sd_vae.values()) / 1e6:.2f} M")
    unet, cn = build_reference(cfg)
    # strict=True: the oracle's parameter spec must equal the reference modules' state-dict keys and shapes
    unet.load_state_dict(sd_unet, strict=True)
    cn.load_state_dict(sd_cn, strict=True)
    dec, pq = build_reference_vae(vcfg)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd_vae.items() if k.startswith("decoder.")}, strict=True)
    pq.load_state_dict({"weight": sd_vae["post_quant_conv.weight"], "bias": sd_vae["post_quant_conv.bias"]})

    x_T, cond, uncond = O.make_inputs(cfg, 1, h, w, hint=hint)
    out = {"cfg": tag, "latent_hw": (h, w), "S": S}
    model = RefModel(unet, cn)
    with torch.no_grad():
        ts = torch.full((1,), 951, dtype=torch.long)
        t0 = time.time()
        with quiet():
            control = cn(x=x_T, hint=cond["c_concat"][0], timesteps=ts, context=cond["c_crossattn"][0])
        out["control_stats"] = torch.tensor([[c.float().mean().item(), c.float().norm().item()] for c in control])
        out["control_last"] = control[-1].clone()
        out["control_first"] = control[0].clone()
        eps_c = model.apply_model(x_T, ts, cond)
        eps_u = model.apply_model(x_T, ts, uncond)
        with quiet():
            eps_nocontrol = unet(x=x_T, timesteps=ts, context=cond["c_crossattn"][0], control=None)
        print(f"[{tag}] reference single step: {time.time() - t0:.1f}s |eps_c| mean {eps_c.abs().mean():.3f}")
        out.update(eps_c_t951=eps_c, eps_u_t951=eps_u, eps_nocontrol_t951=eps_nocontrol)
        # oracle vs reference, same weights & inputs
        o_c = O.apply_model(sd_unet, sd_cn, cfg, x_T, ts, cond)
        o_nc = O.unet_forward(sd_unet, cfg, x_T, ts, cond["c_crossattn"][0])
        print(f"[{tag}] oracle vs reference eps: rel L2 {rel(o_c, eps_c):.2e}; no-control {rel(o_nc, eps_nocontrol):.2e}")
        assert rel(o_c, eps_c) < 1e-4 and rel(o_nc, eps_nocontrol) < 1e-4
        if with_sampler:
            model.calls.clear()
            t0 = time.time()
            samples, sampler = run_reference_sampler(model, x_T, cond, uncond, S, 9.0)
            print(f"[{tag}] reference DDIM {S} steps: {time.time() - t0:.1f}s")
            out["samples"] = samples
            out["call_timesteps"] = torch.tensor([c[0] for c in model.calls])
            out["eps_calls"] = torch.stack([c[1] for c in model.calls])  # [2S, 1, 4, h, w] cond, uncond, cond, ...
            out["ddim_alphas"] = torch.as_tensor(np.asarray(sampler.ddim_alphas), dtype=torch.float64)
            out["ddim_alphas_prev"] = torch.as_tensor(np.asarray(sampler.ddim_alphas_prev), dtype=torch.float64)
            out["ddim_sigmas"] = torch.as_tensor(np.asarray(sampler.ddim_sigmas), dtype=torch.float64)
            out["ddim_sqrt_one_minus_alphas"] = torch.as_tensor(np.asarray(sampler.ddim_sqrt_one_minus_alphas),
                                                                dtype=torch.float64)
            out["ddim_timesteps"] = torch.as_tensor(np.asarray(sampler.ddim_timesteps))
            z = samples
        else:
            z = x_T * 0.18215
        t0 = time.time()
        img = dec(pq(z / vcfg.scale_factor))
        print(f"[{tag}] reference VAE decode: {time.time() - t0:.1f}s")
        out["decode_in"] = z
        out["decoded"] = img
        o_img = O.vae_decode(sd_vae, vcfg, z)
        print(f"[{tag}] oracle vs reference decode: rel L2 {rel(o_img, img):.2e}")
        assert rel(o_img, img) < 1e-4
    torch.save(out, os.path.join(HERE, f"{tag}.pt"))
    print(f"[{tag}] wrote {tag}.pt ({os.path.getsize(os.path.join(HERE, tag + '.pt')) / 1e6:.2f} MB)")


def golden_modules():
    """Small per-module fixtures + the reference's own fused-QKV KAT
    (ldm_torch/modules/test_attention_onnx_torch_error.py:172-200: x=randn(2,10,512), ctx=randn(2,10,77), seed 0)."""
    from ldm.modules.attention import CrossAttention, BasicTransformerBlock
    from ldm.modules.diffusionmodules.openaimodel import ResBlock
    from ldm.modules.diffusionmodules.util import timestep_embedding, GroupNorm32
    out = {}
    torch.manual_seed(0)
    with quiet(), torch.no_grad():
        att_self = CrossAttention(query_dim=512, heads=8, dim_head=64).eval()
        att_cross = CrossAttention(query_dim=512, context_dim=77, heads=8, dim_head=64).eval()
        x = torch.randn(2, 10, 512)
        ctx = torch.randn(2, 10, 77)
        out["kat_x"], out["kat_ctx"] = x, ctx
        out["kat_self_sd"] = {k: v.clone() for k, v in att_self.state_dict().items()}
        out["kat_cross_sd"] = {k: v.clone() for k, v in att_cross.state_dict().items()}
        out["kat_self_y"] = att_self(x)
        out["kat_cross_y"] = att_cross(x, ctx)
        # the fused tensors the reference builds at construction time (attention.py:170,173)
        out["kat_qkv_w_sub"] = att_self.qkv_w[:, ::64].clone()   # subsampled to keep the fixture small
        out["kat_kv_w_sub"] = att_cross.kv_w[:, ::64].clone()
        # the reference script's own check: fused projection == separate projections
        qkv = torch.matmul(x, att_self.qkv_w)
        q, k, v = qkv.chunk(3, dim=-1)
        assert torch.allclose(q, att_self.to_q(x), atol=1e-6) and torch.allclose(v, att_self.to_v(x), atol=1e-6)

        g = torch.Generator().manual_seed(11)
        t = torch.tensor([951, 501, 1])
        out["temb_t"], out["temb"] = t, timestep_embedding(t, 320)
        gn = GroupNorm32(32, 320).eval()
        gn.weight.copy_(1 + 0.1 * torch.randn(320, generator=g))
        gn.bias.copy_(0.1 * torch.randn(320, generator=g))
        xg = torch.randn(2, 320, 8, 12, generator=g) * 2 + 0.5
        out["gn_w"], out["gn_b"], out["gn_x"], out["gn_y"] = gn.weight.clone(), gn.bias.clone(), xg, gn(xg)

        rb = ResBlock(64, 256, 0.0, out_channels=128).eval()
        for p in rb.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        xr, emb = torch.randn(2, 64, 8, 12, generator=g), torch.randn(2, 256, generator=g)
        out["rb_sd"] = {k: v.clone() for k, v in rb.state_dict().items()}
        out["rb_x"], out["rb_emb"], out["rb_y"] = xr, emb, rb(xr, emb)

        tb = BasicTransformerBlock(64, 8, 8, context_dim=96).eval()
        xt, ct = torch.randn(2, 24, 64, generator=g), torch.randn(2, 77, 96, generator=g)
        out["tb_sd"] = {k: v.clone() for k, v in tb.state_dict().items()}
        out["tb_x"], out["tb_ctx"], out["tb_y"] = xt, ct, tb(xt, ct)
    torch.save(out, os.path.join(HERE, "modules.pt"))
    print(f"[modules] wrote modules.pt ({os.path.getsize(os.path.join(HERE, 'modules.pt')) / 1e6:.2f} MB)")


def canny_hint():
    """cv2.Canny(pictures_croped/bird_0.jpg, 100, 200) -> HWC3 -> /255 -> [1,3,256,384]
    (canny2image_torch.py:30-38; annotator/canny/__init__.py:4-6; annotator/util.py:9-38)."""
    import cv2
    img = cv2.imread(os.path.join(REF, "pictures_croped", "bird_0.jpg"))
    assert img is not None and img.shape[:2] == (256, 384), img.shape if img is not None else None
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB) if False else img  # compute_score_torch.py feeds cv2.imread output as is
    edges = cv2.Canny(img, 100, 200)
    hwc3 = np.stack([edges] * 3, axis=2)
    hint = torch.from_numpy(hwc3.copy()).float() / 255.0
    hint = hint.permute(2, 0, 1)[None].contiguous()
    np.save(os.path.join(HERE, "canny_bird0.npy"), np.packbits(edges > 0))
    np.save(os.path.join(HERE, "bird0_bgr.npy"), img)  # decoded pixels: input fixture of the device-side Canny tests
    return hint


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    install_shims()
    golden_modules()
    golden_for(O.TINY, O.TINY_VAE, "tiny", (8, 16), S=4)
    if "--skip-full" not in sys.argv:
        hint = canny_hint()
        golden_for(O.SD15, O.SD15_VAE, "sd15_256x384", (32, 48), S=20, hint=hint)
