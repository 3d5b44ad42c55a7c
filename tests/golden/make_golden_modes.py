"""Generates tests/golden/tiny_modes.pt: the sampler modes beyond plain sampling and the VAE encoder, from
the REAL reference classes in /root/reference (this container only) on the tiny configuration:

  * DDIMSampler.sample with eta > 0 (noise_like draws recorded), and with mask / x0 inpainting blending (q_sample draws
    recorded; LatentDiffusion.q_sample itself is absent from the reference checkout, the duck-typed model supplies the
    standard definition),
  * DDIMSampler.encode / decode / stochastic_encode (cldm/ddim_hacked.py:233-317),
  * ldm.modules.diffusionmodules.model.Encoder (+ a 1x1 quant_conv) on an image.

    python tests/golden/make_golden_modes.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs sys.path entries for the repo and the reference)
from make_golden import O  # noqa: E402

S = 4


def main():
    MG.install_shims()
    cfg, vcfg = O.TINY, O.TINY_VAE
    sd_unet = O.make_weights(O.unet_param_spec(cfg), seed=1234, prefix="unet.")
    sd_cn = O.make_weights(O.controlnet_param_spec(cfg), seed=1234, prefix="control.")
    unet, cn = MG.build_reference(cfg)
    unet.load_state_dict(sd_unet, strict=True)
    cn.load_state_dict(sd_cn, strict=True)
    x_T, cond, uncond = O.make_inputs(cfg, 1, 8, 16)
    model = MG.RefModel(unet, cn)
    ac = model.alphas_cumprod
    q_noises = []

    def q_sample(x_start, t, noise=None):  # ddpm.py LatentDiffusion.q_sample (absent): standard forward diffusion
        noise = torch.randn_like(x_start) if noise is None else noise
        q_noises.append(noise.clone())
        a = ac[t].reshape(-1, 1, 1, 1)
        return a.sqrt() * x_start + (1 - a).sqrt() * noise

    model.q_sample = q_sample
    import cldm.ddim_hacked as RH
    step_noises = []
    real_noise_like = RH.noise_like

    def recording_noise_like(shape, device, repeat=False):
        n = real_noise_like(shape, device, repeat)
        step_noises.append(n.clone())
        return n

    RH.noise_like = recording_noise_like
    eps_fn = lambda x, t, c: O.apply_model(sd_unet, sd_cn, cfg, x, t, c)
    out = {"S": S}

    def new_sampler():
        s = RH.DDIMSampler(model)
        s.register_buffer = lambda name, attr: setattr(s, name, attr)  # no CUDA device here (ddim_hacked.py:17-21)
        return s

    with torch.no_grad(), MG.quiet():
        # ---- eta > 0 -------------------------------------------------------------------------------------------
        torch.manual_seed(11)
        step_noises.clear()
        smp, _ = new_sampler().sample(S, 1, (4, 8, 16), cond, verbose=False, eta=0.5, x_T=x_T,
                                      unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
        out["eta"] = 0.5
        out["eta_noises"] = torch.stack(step_noises)
        out["eta_samples"] = smp
        o_smp, _ = O.ddim_sample(eps_fn, x_T, cond, uncond, S=S, scale=9.0, eta=0.5, noises=list(out["eta_noises"]))
        out["_oracle_eta"] = MG.rel(o_smp, smp)
        # ---- mask / x0 -----------------------------------------------------------------------------------------
        torch.manual_seed(12)
        g = torch.Generator().manual_seed(21)
        x0 = torch.randn((1, 4, 8, 16), generator=g)
        mask = (torch.rand((1, 1, 8, 16), generator=g) > 0.5).float()
        q_noises.clear()
        smp, _ = new_sampler().sample(S, 1, (4, 8, 16), cond, verbose=False, eta=0.0, x_T=x_T, mask=mask, x0=x0,
                                      unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
        out.update(mask=mask, mask_x0=x0, mask_q_noises=torch.stack(q_noises), mask_samples=smp)
        o_smp, _ = O.ddim_sample(eps_fn, x_T, cond, uncond, S=S, scale=9.0, mask=mask, x0=x0,
                                 q_noises=list(out["mask_q_noises"]))
        out["_oracle_mask"] = MG.rel(o_smp, smp)
        # ---- encode / decode / stochastic_encode -----------------------------------------------------------------
        s = new_sampler()
        s.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
        z0 = x0 * 0.5
        enc, _ = s.encode(z0, cond, t_enc=3)
        out.update(encode_x0=z0, encode_t_enc=3, encoded=enc)
        out["_oracle_encode"] = MG.rel(O.ddim_encode(eps_fn, z0, cond, 3, S=S), enc)
        dec = s.decode(enc, cond, t_start=3, unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
        out.update(decode_t_start=3, decoded_latent=dec)
        out["_oracle_decode"] = MG.rel(O.ddim_decode(eps_fn, enc, cond, 3, S=S, scale=9.0, uncond=uncond), dec)
        t_idx = torch.tensor([2], dtype=torch.long)
        noise = torch.randn((1, 4, 8, 16), generator=g)
        st = s.stochastic_encode(z0, t_idx, noise=noise)
        out.update(stoch_t=t_idx, stoch_noise=noise, stoch_encoded=st)
        out["_oracle_stoch"] = MG.rel(O.stochastic_encode(z0, t_idx, S=S, noise=noise), st)
        # ---- VAE encoder -----------------------------------------------------------------------------------------
        from ldm.modules.diffusionmodules.model import Encoder
        encoder = Encoder(ch=vcfg.ch, out_ch=vcfg.out_ch, ch_mult=vcfg.ch_mult, num_res_blocks=vcfg.num_res_blocks,
                          attn_resolutions=[], dropout=0.0, in_channels=3, resolution=256, z_channels=vcfg.z_channels,
                          double_z=True).eval()
        sd_enc = O.make_weights(O.vae_encoder_param_spec(vcfg), seed=1234, prefix="vae.")
        encoder.load_state_dict({k[len("encoder."):]: v for k, v in sd_enc.items() if k.startswith("encoder.")}, strict=True)
        qc = torch.nn.Conv2d(2 * vcfg.z_channels, 2 * vcfg.z_channels, 1)
        qc.load_state_dict({"weight": sd_enc["quant_conv.weight"], "bias": sd_enc["quant_conv.bias"]})
        img = torch.rand((1, 3, 64, 128), generator=g) * 2 - 1
        moments = qc(encoder(img))
        out.update(enc_image=img, enc_moments=moments)
        out["_oracle_vae_encode"] = MG.rel(O.vae_encode(sd_enc, vcfg, img), moments)
    RH.noise_like = real_noise_like
    for k, v in out.items():
        if k.startswith("_oracle"):
            print(f"oracle vs reference {k[8:]}: rel L2 {v:.2e}")
            assert v < 1e-4, k
    torch.save(out, os.path.join(HERE, "tiny_modes.pt"))
    print("wrote tiny_modes.pt", os.path.getsize(os.path.join(HERE, "tiny_modes.pt")) / 1e3, "KB")


if __name__ == "__main__":
    main()
