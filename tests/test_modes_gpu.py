"""SURVEY 8f-4 on the GPU: sampler modes beyond plain sampling (eta > 0, mask / x0 inpainting blend, encode / decode /
stochastic_encode of cldm/ddim_hacked.py:154-157,233-317) and the VAE Encoder (model.py:452-543, asymmetric-pad
Downsample :66-87), against fixtures produced by the REAL reference classes (tests/golden/make_golden_modes.py) and the
CPU oracle. Random draws are replayed from the fixtures (the reference draws them from torch's CPU generator)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import O, build_control_ldm, inputs_on, load_golden, oracle_weights, rel_l2, unet_kwargs, vae_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny(cuda_device):
    return build_control_ldm(O.TINY, O.TINY_VAE, cuda_device), load_golden("tiny_modes")


def _sampler(model):
    from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler
    return DDIMSampler(model)


# bf16 gates are the sampler-level ones of tests/test_model_gpu.py (errors compound over the steps and are multiplied by the
# guidance scale 9); the fp32-mode run of the same code pins the LOGIC of each mode at 1e-3.
PRECISIONS = [("bf16", 3e-2), ("fp32", 1e-3)]


@pytest.fixture(params=PRECISIONS, ids=[p[0] for p in PRECISIONS])
def prec(request, tiny):
    tiny[0].precision = request.param[0]
    yield request.param
    tiny[0].precision = "bf16"


def test_eta_sampling(tiny, cuda_device, monkeypatch, prec):
    import stablediffusioneo_b200.cldm.ddim_hacked as H
    model, g = tiny
    x_T, cond, uncond = inputs_on(O.TINY, 8, 16, cuda_device)
    draws = iter(g["eta_noises"])
    monkeypatch.setattr(H, "noise_like", lambda shape, device, repeat=False: next(draws).to(device))
    smp, _ = _sampler(model).sample(g["S"], 1, (4, 8, 16), cond, verbose=False, eta=g["eta"], x_T=x_T,
                                    unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
    err = rel_l2(smp, g["eta_samples"])
    print(f"eta=0.5 samples rel L2 ({prec[0]}):", err)
    assert err < prec[1]


def test_mask_blend_sampling(tiny, cuda_device, monkeypatch, prec):
    model, g = tiny
    x_T, cond, uncond = inputs_on(O.TINY, 8, 16, cuda_device)
    draws = iter(g["mask_q_noises"])
    real_q = model.q_sample
    monkeypatch.setattr(model, "q_sample", lambda x0, t, noise=None: real_q(x0, t, noise=next(draws).to(x0.device)),
                        raising=False)
    sampler = _sampler(model)
    smp, _ = sampler.sample(g["S"], 1, (4, 8, 16), cond, verbose=False, eta=0.0, x_T=x_T, mask=g["mask"].to(cuda_device),
                            x0=g["mask_x0"].to(cuda_device), unconditional_guidance_scale=9.0,
                            unconditional_conditioning=uncond)
    # bf16: the captured-graph engine (blend = first node of the step graph, q_sample draws tabled up front); fp32 mode: step by step
    assert (sampler._engine is not None) == (prec[0] == "bf16")
    err = rel_l2(smp, g["mask_samples"])
    print(f"masked samples rel L2 ({prec[0]}):", err)
    assert err < prec[1]
    # where mask == 1 the last blend happened BEFORE the final step, so nothing is bit-copied; the kernel itself:
    from stablediffusioneo_b200 import ops
    gen = torch.Generator().manual_seed(3)
    x0, nz, img = (torch.randn((2, 4, 8, 16), generator=gen) for _ in range(3))
    for mc in (1, 4):
        m = (torch.rand((2, mc, 8, 16), generator=gen) > 0.5).float()
        a, b = torch.tensor([0.3, 0.9]), torch.tensor([0.7, 0.1])
        ref = (a.view(2, 1, 1, 1) * x0 + b.view(2, 1, 1, 1) * nz) * m + (1 - m) * img
        dev = cuda_device
        got = ops.mask_blend(x0.to(dev), nz.to(dev), img.to(dev), m.to(dev), a.to(dev), b.to(dev))
        assert torch.allclose(got.cpu(), ref, atol=1e-6)


def test_encode_decode_stochastic(tiny, cuda_device, prec):
    model, g = tiny
    dev = cuda_device
    _, cond, uncond = inputs_on(O.TINY, 8, 16, dev)
    s = _sampler(model)
    s.make_schedule(ddim_num_steps=g["S"], ddim_eta=0.0, verbose=False)
    enc, info = s.encode(g["encode_x0"].to(dev), cond, t_enc=g["encode_t_enc"])
    e1 = rel_l2(enc, g["encoded"])
    assert info["x_encoded"] is enc
    dec = s.decode(g["encoded"].to(dev), cond, t_start=g["decode_t_start"], unconditional_guidance_scale=9.0,
                   unconditional_conditioning=uncond)
    e2 = rel_l2(dec, g["decoded_latent"])
    st = s.stochastic_encode(g["encode_x0"].to(dev), g["stoch_t"].to(dev), noise=g["stoch_noise"].to(dev))
    e3 = rel_l2(st, g["stoch_encoded"])
    print(f"encode / decode / stochastic_encode rel L2 ({prec[0]}):", e1, e2, e3)
    # (decode: 3 guided steps onto a small-norm latent; bf16 measured 6e-2, fp32 mode shows the logic is exact)
    assert e1 < prec[1] and e2 < 3 * prec[1] and e3 < 1e-6
    # encode with classifier-free guidance (two calls for dict conditionings) vs the oracle
    sd_unet, sd_cn, _ = oracle_weights(O.TINY, O.TINY_VAE)
    _, cond_c, uncond_c = O.make_inputs(O.TINY, 1, 8, 16)
    eps_fn = lambda x, t, c: O.apply_model(sd_unet, sd_cn, O.TINY, x, t, c)
    with torch.no_grad():
        ref = O.ddim_encode(eps_fn, g["encode_x0"], cond_c, 2, S=g["S"], scale=3.0, uncond=uncond_c)
    enc2, _ = s.encode(g["encode_x0"].to(dev), cond, t_enc=2, unconditional_guidance_scale=3.0,
                       unconditional_conditioning=uncond)
    assert rel_l2(enc2, ref) < prec[1]


def test_use_original_steps(tiny, cuda_device, prec):
    """use_original_steps (ddim_hacked.py:203-206, 236-244, 301): decode / encode over the 1000-step DDPM tables instead of
    the DDIM subsequence, against the oracle's eps + the closed-form update with those tables. (The reference's own
    p_sample_ddim reads the sigmas from the wrong object on this path and raises; the tables are what it specifies.)"""
    import numpy as np
    model, g = tiny
    dev = cuda_device
    _, cond, uncond = inputs_on(O.TINY, 8, 16, dev)
    sd_unet, sd_cn, _ = oracle_weights(O.TINY, O.TINY_VAE)
    _, cond_c, uncond_c = O.make_inputs(O.TINY, 1, 8, 16)
    eps_fn = lambda x, t, c: O.apply_model(sd_unet, sd_cn, O.TINY, x, t, c)
    ac = O.alphas_cumprod().astype(np.float32)
    ac_prev = np.append(np.float32(1.0), ac[:-1])
    s = _sampler(model)
    s.make_schedule(ddim_num_steps=g["S"], ddim_eta=0.0, verbose=False)
    x = torch.randn((1, 4, 8, 16), generator=torch.Generator().manual_seed(7))
    t_start, scale = 3, 3.0
    dec = s.decode(x.to(dev), cond, t_start=t_start, unconditional_guidance_scale=scale, unconditional_conditioning=uncond,
                   use_original_steps=True)
    ref = x.clone()
    with torch.no_grad():
        for step in reversed(range(t_start)):      # DDPM timesteps t_start-1 .. 0, index == timestep
            t = torch.full((1,), step, dtype=torch.long)
            e_c, e_u = eps_fn(ref, t, cond_c), eps_fn(ref, t, uncond_c)
            e_t = e_u + scale * (e_c - e_u)
            ref, _ = O.ddim_update(ref, e_t, float(ac[step]), float(ac_prev[step]), 0.0, float(np.sqrt(1.0 - ac[step])))
    e1 = rel_l2(dec, ref)
    enc, _ = s.encode(x.to(dev), cond, t_enc=3, use_original_steps=True)
    ref = x.clone()
    with torch.no_grad():
        for i in range(3):                          # ddim_hacked.py:242-265 with alphas_cumprod[_prev][:num_steps]
            t = torch.full((1,), i, dtype=torch.long)
            e = eps_fn(ref, t, cond_c)
            an, a = float(ac[i]), float(ac_prev[i])
            ref = np.sqrt(an / a) * ref + np.sqrt(an) * (np.sqrt(1 / an - 1) - np.sqrt(1 / a - 1)) * e
    e2 = rel_l2(enc, ref)
    print(f"use_original_steps decode / encode rel L2 ({prec[0]}):", e1, e2)
    assert e1 < 3 * prec[1] and e2 < prec[1]


def test_asymmetric_pad_downsample_conv(cuda_device):
    """F.pad(x, (0,1,0,1)) + conv3x3 stride 2 padding 0 through sdeo_conv_args::pad_hi, even and odd sizes."""
    from stablediffusioneo_b200 import ops
    gen = torch.Generator().manual_seed(5)
    for n, c, h, w, co in ((1, 128, 64, 96, 128), (2, 32, 16, 24, 32), (1, 64, 15, 21, 96)):
        x = torch.randn((n, c, h, w), generator=gen)
        wt = torch.randn((co, c, 3, 3), generator=gen) * (c * 9) ** -0.5
        b = torch.randn((co,), generator=gen) * 0.1
        xb, wb = x.bfloat16().float(), wt.bfloat16().float()
        ref = F.conv2d(F.pad(xb, (0, 1, 0, 1)), wb, b, stride=2)
        dev = cuda_device
        y = ops.conv2d(xb.permute(0, 2, 3, 1).contiguous().bfloat16().to(dev), ops.pack_conv_weight(wt.to(dev)),
                       bias=b.to(dev), stride=2, out_fp32=True, pad_hi=1)
        assert y.shape == (n, ref.shape[2], ref.shape[3], co)
        assert rel_l2(y.permute(0, 3, 1, 2), ref) < 1e-3, (n, c, h, w)


@pytest.mark.parametrize("cout", [8, 24, 40])
@pytest.mark.parametrize("f32", [False, True])
def test_conv_cout_not_multiple_of_16(cuda_device, cout, f32):
    """cout % 16 == 8 (the VAE encoder's 8-channel conv_out / quant_conv): the last N tile has 8 columns beyond cout that
    must not be stored (regression: the vector epilogue wrote them over the next pixel)."""
    from stablediffusioneo_b200 import ops
    gen = torch.Generator().manual_seed(cout)
    x = torch.randn((2, 128, 8, 16), generator=gen).bfloat16()
    wt = (torch.randn((cout, 128, 3, 3), generator=gen) * (128 * 9) ** -0.5).bfloat16().float()
    b = torch.randn((cout,), generator=gen) * 0.1
    r = torch.randn((2, cout, 8, 16), generator=gen).bfloat16()
    ref = F.conv2d(x.float(), wt, b, padding=1) + r.float()
    dev = cuda_device
    y = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().to(dev), ops.pack_conv_weight(wt.to(dev)), bias=b.to(dev),
                   residual=r.permute(0, 2, 3, 1).contiguous().to(dev), out_fp32=f32)
    assert y.shape == (2, 8, 16, cout)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < (1e-3 if f32 else 6e-3)


def test_vae_encoder(cuda_device):
    """Tiny VAE Encoder + quant_conv vs the real reference Encoder's moments; encode_first_stage = scale_factor * mean;
    encode -> decode round trip shape."""
    from stablediffusioneo_b200.cldm.cldm import ControlLDM
    g = load_golden("tiny_modes")
    dev = cuda_device
    with torch.device(dev):
        model = ControlLDM(unet_config=unet_kwargs(O.TINY), first_stage_config=vae_kwargs(O.TINY_VAE),
                           first_stage_encoder=True).eval()
    sd_enc = O.make_weights(O.vae_encoder_param_spec(O.TINY_VAE), seed=1234, prefix="vae.")
    missing, unexpected = model.first_stage_model.load_state_dict(sd_enc, strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "post_quant_conv.")) for k in missing)
    img = g["enc_image"].to(dev)
    m = model.first_stage_model.encode_moments(img)
    err = rel_l2(m, g["enc_moments"])
    print("VAE encoder moments rel L2:", err)
    assert m.shape == g["enc_moments"].shape and err < 2e-2
    z = model.encode_first_stage(img)
    assert rel_l2(z, g["enc_moments"][:, :4] * 0.18215) < 2e-2
    nz = torch.randn((1, 4, 8, 16), generator=torch.Generator().manual_seed(1))
    zs = model.encode_first_stage(img, sample=True, noise=nz.to(dev))
    mom = g["enc_moments"]
    ref = 0.18215 * (mom[:, :4] + torch.exp(0.5 * mom[:, 4:].clamp(-30, 20)) * nz)
    assert rel_l2(zs, ref) < 2e-2
    assert model.decode_first_stage(z).shape == (1, 3, 64, 128)


@pytest.mark.parametrize("graph", [False, True])
def test_engine_infer_surface(tiny, cuda_device, graph):
    """the reference's TensorRT-engine call pattern (cldm_trt/ddim_hacked.py:140-152) against
    Engine.infer of this package -- controlnet engine -> 13 controls (dict values [4:17]) -> unet engine -> 'latent' -- equals
    apply_model; second call with other inputs replays the captured graph."""
    from stablediffusioneo_b200.Engine import Engine
    model, _ = tiny
    model.precision = "bf16"
    dev = cuda_device
    x_T, cond, uncond = inputs_on(O.TINY, 8, 16, dev)
    cn = Engine(model, "controlnet", latent_h=8, latent_w=16)
    un = Engine(model, "unet", latent_h=8, latent_w=16)
    stream = torch.cuda.Stream(device=dev)
    stream.wait_stream(torch.cuda.current_stream())
    for c in (cond, uncond):
        ts = torch.full((1,), 951, dtype=torch.int32, device=dev)  # the ONNX export feeds int32 timesteps
        out = cn.infer({"x_noisy": x_T, "hint": c["c_concat"][0], "timestep": ts, "context": c["c_crossattn"][0]},
                       stream=stream, use_cuda_graph=graph)
        control = list(out.values())
        assert len(control) == 17
        feed = {"x_noisy": x_T, "timestep": ts, "context": c["c_crossattn"][0]}
        feed.update({f"control{i}": control[4 + i] for i in range(13)})
        eps = un.infer(feed, stream, use_cuda_graph=graph)["latent"].clone()
        stream.synchronize()
        ref = model.apply_model(x_T, ts.long(), c)
        err = rel_l2(eps, ref)
        print("engine surface vs apply_model rel L2:", err)
        assert err < 1.5e-2  # two bf16 paths of ours: unfused control injection (bf16 adds) vs the fused epilogue
    dec = Engine(model, "decoder", latent_h=8, latent_w=16)
    img = dec.infer({"latent": x_T * 0.18215})["images"]
    assert img.shape == (1, 3, 64, 128)


def test_canny_bit_exact(cuda_device):
    """the hint preprocessing on the device. Bit-exact against the numpy oracle (itself pinned to cv2.Canny)
    and the committed bird_0 fixture; CannyDetector numpy / tensor interfaces; the [N,3,H,W] hint tensor."""
    import os
    import numpy as np
    from helpers import GOLDEN, canny_hint
    from oracle.canny_oracle import canny as canny_ref
    from stablediffusioneo_b200.annotator.canny import CannyDetector
    from stablediffusioneo_b200.annotator.util import HWC3, resize_image
    det = CannyDetector(cuda_device)
    img = np.load(os.path.join(GOLDEN, "bird0_bgr.npy"))
    img = resize_image(HWC3(img), 256)                         # identity at this size (canny2image_torch.py:30)
    bits = np.unpackbits(np.load(os.path.join(GOLDEN, "canny_bird0.npy")))[: 256 * 384].reshape(256, 384)
    edges = det(img, 100, 200)
    assert edges.dtype == np.uint8 and np.array_equal(edges, bits * 255)
    hint, dmap = det.hint(img, 100, 200, num_samples=2)
    assert hint.shape == (2, 3, 256, 384) and torch.equal(hint[1:].cpu(), canny_hint())
    assert torch.equal(dmap.cpu(), torch.from_numpy(bits * 255).to(torch.uint8))
    rng = np.random.default_rng(1)
    for shape, lo, hi in (((64, 96, 3), 300, 700), ((37, 53, 3), 20, 60), ((40, 40), 100, 200), ((512, 512, 3), 250, 500),
                          ((33, 1000, 1), 200, 100)):
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        got = det(torch.from_numpy(a).to(cuda_device), lo, hi).cpu().numpy()
        ref = canny_ref(a, lo, hi)
        assert np.array_equal(got, ref), (shape, int((got != ref).sum()))
    try:
        import cv2
        a = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
        a = cv2.GaussianBlur(a, (0, 0), 1.5)
        assert np.array_equal(det(a, 30, 90), cv2.Canny(a, 30, 90))
    except ImportError:
        pass


@pytest.mark.parametrize("guess_mode", [False, True])
def test_process_pipeline(tiny, cuda_device, guess_mode):
    """hackathon.process (canny2image_torch.py:28-70) end to end on the tiny model: image -> Canny hint -> seeded x_T ->
    DDIM (engine path; generic path in guess mode) -> VAE decode -> uint8, against the oracle pipeline fed the SAME x_T
    (drawn by the seeded CUDA generator, as the reference draws it)."""
    import os
    import numpy as np
    from helpers import GOLDEN
    from oracle.canny_oracle import canny as canny_ref
    from stablediffusioneo_b200.canny2image import hackathon, seed_everything
    model, _ = tiny
    model.precision = "bf16"
    dev = cuda_device
    img = np.load(os.path.join(GOLDEN, "bird0_bgr.npy"))[64:128, 128:256].copy()          # 64 x 128 crop
    _, cond_c, uncond_c = O.make_inputs(O.TINY, 1, 8, 16)
    ctx_c, ctx_u = cond_c["c_crossattn"][0], uncond_c["c_crossattn"][0]
    hk = hackathon()
    hk.initialize(model=model, device=dev)
    try:
        out = hk.process(img, ctx_c, "", ctx_u, 1, 64, 4, guess_mode, 0.9, 9.0, 2946901, 0.0, 100, 200)
        scales = list(model.control_scales)
    finally:
        model.control_scales = [1.0] * 13
    assert len(out) == 1 and out[0].shape == (64, 128, 3) and out[0].dtype == np.uint8
    # oracle pipeline
    edges = canny_ref(img, 100, 200)
    assert np.array_equal(hk.detected_map.cpu().numpy(), edges)
    hint = torch.from_numpy(np.stack([edges] * 3, 0)[None].astype(np.float32) / 255.0)
    seed_everything(2946901)
    x_T = torch.randn((1, 4, 8, 16), device=dev).cpu()
    sd_unet, sd_cn, sd_vae = oracle_weights(O.TINY, O.TINY_VAE)
    eps_fn = lambda x, t, c: O.apply_model(sd_unet, sd_cn, O.TINY, x, t, c, control_scales=scales)
    cond = {"c_concat": [hint], "c_crossattn": [ctx_c]}
    un = {"c_concat": None if guess_mode else [hint], "c_crossattn": [ctx_u]}
    with torch.no_grad():
        lat, _ = O.ddim_sample(eps_fn, x_T, cond, un, S=4, scale=9.0)
        ref = O.to_uint8_image(O.vae_decode(sd_vae, O.TINY_VAE, lat))[0]
    diff = np.abs(out[0].astype(np.int32) - ref.astype(np.int32)).astype(np.float64)
    psnr = 10 * np.log10(255.0 ** 2 / max((diff ** 2).mean(), 1e-12))
    print(f"process() image vs oracle pipeline (guess_mode={guess_mode}): mean abs diff {diff.mean():.3f}, PSNR {psnr:.1f} dB")
    assert psnr > 30.0
