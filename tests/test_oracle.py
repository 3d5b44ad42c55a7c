"""CPU tests (-m "not gpu"): the oracle (oracle/sd15_oracle.py) against the golden vectors produced by the REAL
reference modules (tests/golden/make_golden.py), plus the reference's own fused-QKV known-answer test."""
import numpy as np
import pytest
import torch

from helpers import O, canny_hint, load_golden, oracle_weights, rel_l2


@pytest.fixture(scope="module")
def mods():
    return load_golden("modules")


def test_fused_qkv_kat(mods):
    """ldm_torch/modules/test_attention_onnx_torch_error.py:172-200 — x=randn(2,10,512), ctx=randn(2,10,77), seed 0:
    the fused-weight projection (x @ cat([Wq,Wk,Wv]).T, ctx @ cat([Wk,Wv]).T) equals the separate projections, atol 1e-6."""
    x, ctx = mods["kat_x"], mods["kat_ctx"]
    sd_s = {"a." + k: v for k, v in mods["kat_self_sd"].items()}
    sd_c = {"a." + k: v for k, v in mods["kat_cross_sd"].items()}
    y_plain = O.cross_attention(sd_s, "a.", x, None, 8)
    y_fused = O.cross_attention_fused(sd_s, "a.", x, None, 8)
    assert torch.allclose(y_fused, y_plain, atol=1e-6)
    assert torch.allclose(y_plain, mods["kat_self_y"], atol=1e-5)
    yc_plain = O.cross_attention(sd_c, "a.", x, ctx, 8)
    yc_fused = O.cross_attention_fused(sd_c, "a.", x, ctx, 8)
    assert torch.allclose(yc_fused, yc_plain, atol=1e-6)
    assert torch.allclose(yc_plain, mods["kat_cross_y"], atol=1e-5)
    # the fused tensors' layout: [in, 3*inner] = cat([Wq, Wk, Wv]).T (attention.py:170), [ctx, 2*inner] (attention.py:173)
    qkv_w = torch.cat([sd_s["a.to_q.weight"], sd_s["a.to_k.weight"], sd_s["a.to_v.weight"]]).t()
    assert torch.equal(qkv_w[:, ::64], mods["kat_qkv_w_sub"])
    kv_w = torch.cat([sd_c["a.to_k.weight"], sd_c["a.to_v.weight"]]).t()
    assert torch.equal(kv_w[:, ::64], mods["kat_kv_w_sub"])


def test_leaf_ops(mods):
    assert torch.allclose(O.timestep_embedding(mods["temb_t"], 320), mods["temb"], atol=1e-6)
    sd = {"g.weight": mods["gn_w"], "g.bias": mods["gn_b"]}
    assert torch.allclose(O._gn(sd, "g.", mods["gn_x"], 1e-5), mods["gn_y"], atol=1e-5)
    sd = {"r." + k: v for k, v in mods["rb_sd"].items()}
    assert rel_l2(O.resblock(sd, "r.", mods["rb_x"], mods["rb_emb"]), mods["rb_y"]) < 1e-5
    sd = {"t." + k: v for k, v in mods["tb_sd"].items()}
    assert rel_l2(O.transformer_block(sd, "t.", mods["tb_x"], mods["tb_ctx"], 8), mods["tb_y"]) < 1e-5


def test_tiny_network_sampler_decoder():
    """Same topology as SD1.5 at 1/5 width: ControlNet (13 outputs), controlled UNet, 4-step DDIM with CFG 9, VAE decode."""
    g = load_golden("tiny")
    cfg, vcfg = O.TINY, O.TINY_VAE
    sd_unet, sd_cn, sd_vae = oracle_weights(cfg, vcfg)
    x_T, cond, uncond = O.make_inputs(cfg, 1, 8, 16)
    ts = torch.full((1,), 951, dtype=torch.long)
    with torch.no_grad():
        control = O.controlnet_forward(sd_cn, cfg, x_T, cond["c_concat"][0], ts, cond["c_crossattn"][0])
        assert len(control) == 13
        assert rel_l2(control[-1], g["control_last"]) < 1e-5 and rel_l2(control[0], g["control_first"]) < 1e-5
        assert rel_l2(O.apply_model(sd_unet, sd_cn, cfg, x_T, ts, cond), g["eps_c_t951"]) < 1e-5
        assert rel_l2(O.apply_model(sd_unet, sd_cn, cfg, x_T, ts, uncond), g["eps_u_t951"]) < 1e-5
        assert rel_l2(O.unet_forward(sd_unet, cfg, x_T, ts, cond["c_crossattn"][0]), g["eps_nocontrol_t951"]) < 1e-5
        samples, trace = O.ddim_sample(lambda x, t, c: O.apply_model(sd_unet, sd_cn, cfg, x, t, c), x_T, cond, uncond,
                                       S=g["S"], scale=9.0, collect=True)
        assert rel_l2(samples, g["samples"]) < 1e-4
        # call order and timesteps of the reference sampler: cond then uncond at t = flip(ddim_timesteps)
        assert [tr["t"] for tr in trace] == g["call_timesteps"][::2].tolist()
        assert rel_l2(O.vae_decode(sd_vae, vcfg, g["decode_in"]), g["decoded"]) < 1e-5


def test_ddim_schedule_tables():
    g = load_golden("sd15_256x384")
    sch = O.ddim_schedule(20)
    assert np.array_equal(sch["timesteps"], g["ddim_timesteps"].numpy())
    assert sch["timesteps"][0] == 1 and sch["timesteps"][-1] == 951
    for k, gk in (("alphas", "ddim_alphas"), ("alphas_prev", "ddim_alphas_prev"), ("sigmas", "ddim_sigmas"),
                  ("sqrt_one_minus_alphas", "ddim_sqrt_one_minus_alphas")):
        assert np.allclose(np.asarray(sch[k], dtype=np.float64), g[gk].numpy(), rtol=1e-6, atol=1e-9), k
    # SURVEY §8 a-1 table ends
    assert abs(float(sch["alphas"][19]) - 0.008155) < 1e-5 and abs(float(sch["alphas_prev"][19]) - 0.014005) < 1e-5
    assert abs(float(sch["alphas"][0]) - 0.998296) < 1e-5 and abs(float(sch["alphas_prev"][0]) - 0.999150) < 1e-5


def test_sd15_trajectory_from_reference_eps():
    """Replaying the reference's per-call eps through the oracle's CFG + DDIM update reproduces its 20-step latents."""
    g = load_golden("sd15_256x384")
    sch = O.ddim_schedule(20)
    x = O.make_inputs(O.SD15, 1, 32, 48, hint=canny_hint())[0]
    assert g["call_timesteps"].tolist() == [t for s in reversed(sch["timesteps"].tolist()) for t in (s, s)]
    for i in range(20):
        index = 19 - i
        e_c, e_u = g["eps_calls"][2 * i], g["eps_calls"][2 * i + 1]
        x, _ = O.ddim_update(x, e_u + 9.0 * (e_c - e_u), float(sch["alphas"][index]), float(sch["alphas_prev"][index]),
                             0.0, float(sch["sqrt_one_minus_alphas"][index]))
    assert rel_l2(x, g["samples"]) < 1e-5


def test_sd15_single_step_full_size():
    """Full SD1.5 ControlNet + UNet on the 256x384 workload at t=951 vs the real reference (about a minute on CPU)."""
    g = load_golden("sd15_256x384")
    cfg = O.SD15
    sd_unet = O.make_weights(O.unet_param_spec(cfg), seed=1234, prefix="unet.")
    sd_cn = O.make_weights(O.controlnet_param_spec(cfg), seed=1234, prefix="control.")
    x_T, cond, _ = O.make_inputs(cfg, 1, 32, 48, hint=canny_hint())
    with torch.no_grad():
        eps = O.apply_model(sd_unet, sd_cn, cfg, x_T, torch.full((1,), 951, dtype=torch.long), cond)
    assert rel_l2(eps, g["eps_c_t951"]) < 1e-4


def test_sampler_modes_and_vae_encoder_vs_reference():
    """eta > 0, mask / x0 blending, encode / decode / stochastic_encode and the VAE Encoder, oracle vs the
    fixtures made by the real reference classes (tests/golden/make_golden_modes.py)."""
    g = load_golden("tiny_modes")
    cfg, vcfg = O.TINY, O.TINY_VAE
    sd_unet, sd_cn, _ = oracle_weights(cfg, vcfg)
    eps_fn = lambda x, t, c: O.apply_model(sd_unet, sd_cn, cfg, x, t, c)
    x_T, cond, uncond = O.make_inputs(cfg, 1, 8, 16)
    S = g["S"]
    with torch.no_grad():
        smp, _ = O.ddim_sample(eps_fn, x_T, cond, uncond, S=S, scale=9.0, eta=g["eta"], noises=list(g["eta_noises"]))
        assert rel_l2(smp, g["eta_samples"]) < 1e-4
        smp, _ = O.ddim_sample(eps_fn, x_T, cond, uncond, S=S, scale=9.0, mask=g["mask"], x0=g["mask_x0"],
                               q_noises=list(g["mask_q_noises"]))
        assert rel_l2(smp, g["mask_samples"]) < 1e-4
        enc = O.ddim_encode(eps_fn, g["encode_x0"], cond, g["encode_t_enc"], S=S)
        assert rel_l2(enc, g["encoded"]) < 1e-4
        dec = O.ddim_decode(eps_fn, g["encoded"], cond, g["decode_t_start"], S=S, scale=9.0, uncond=uncond)
        assert rel_l2(dec, g["decoded_latent"]) < 1e-4
        assert rel_l2(O.stochastic_encode(g["encode_x0"], g["stoch_t"], S=S, noise=g["stoch_noise"]), g["stoch_encoded"]) < 1e-6
        sd_enc = O.make_weights(O.vae_encoder_param_spec(vcfg), seed=1234, prefix="vae.")
        assert rel_l2(O.vae_encode(sd_enc, vcfg, g["enc_image"]), g["enc_moments"]) < 1e-5


def test_canny_oracle_vs_cv2_and_fixture():
    """oracle/canny_oracle.py against cv2.Canny itself (the third-party arithmetic behind annotator/canny) and against the
    committed hint fixture (cv2.Canny of pictures_croped/bird_0.jpg as cv2.imread returns it, thresholds 100 / 200)."""
    import os
    from helpers import GOLDEN
    from oracle.canny_oracle import canny
    img = np.load(os.path.join(GOLDEN, "bird0_bgr.npy"))
    bits = np.unpackbits(np.load(os.path.join(GOLDEN, "canny_bird0.npy")))[: 256 * 384].reshape(256, 384)
    assert np.array_equal(canny(img, 100, 200), bits * 255)
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    cases = [(img, 100, 200), (img[:, :, 1].copy(), 50, 120)]
    for shape in ((64, 96, 3), (37, 53, 3), (40, 40)):
        noise = rng.integers(0, 256, shape, dtype=np.uint8)
        cases += [(noise, 300, 700), (cv2.GaussianBlur(noise, (0, 0), 2.0), 20, 60)]
    for a, lo, hi in cases:
        assert np.array_equal(canny(a, lo, hi), cv2.Canny(a, lo, hi)), (a.shape, lo, hi)
