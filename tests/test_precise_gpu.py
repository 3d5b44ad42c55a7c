"""fp32 ("precise") mode: split-operand GEMMs on the bf16 tensor cores + fp32 norms / attention / activations
(stablediffusioneo_b200/precise.py, csrc/precise.cu). Gate from north_star / SURVEY 8c: per-step eps relative L2 <= 1e-4
against the reference's fp32 PyTorch path (golden vectors made by the real reference modules, tests/golden/make_golden.py).
Kernel-level checks compare against torch on the CPU in float64."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import O, build_control_ldm, canny_hint, inputs_on, load_golden, rel_l2

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("terms,tol", [(3, 3e-5), (6, 1e-5)])
@pytest.mark.parametrize("rows,kdim,ndim", [(3072, 320, 320), (48, 1280, 1280), (77, 768, 640), (2, 1280, 320)])
def test_split_linear(cuda_device, terms, tol, rows, kdim, ndim):
    """x @ W.T + b through the split-term bf16 GEMM vs float64."""
    from stablediffusioneo_b200 import ops
    x, w, b = _rand((rows, kdim), 1), _rand((ndim, kdim), 2, kdim ** -0.5), _rand((ndim,), 3, 0.1)
    ref = x.double() @ w.double().t() + b.double()
    pw = ops.pack_conv_weight_split(w.to(cuda_device), terms)
    y = ops.linear(ops.split_terms(x.to(cuda_device), terms), pw, bias=b.to(cuda_device), out_fp32=True)
    err = rel_l2(y, ref)
    print(f"split linear terms={terms} rows={rows} K={kdim}: rel L2 {err:.2e}")
    assert err < tol


@pytest.mark.parametrize("case", ["plain3x3", "stride2", "concat", "nchw4", "upsample", "epilogue"])
def test_split_conv(cuda_device, case):
    """3x3 / 1x1 convolutions of fp32 mode (stride 2, fused concat, ragged NCHW input, nearest-upsampled input, the full
    epilogue act(acc + bias + emb) * scale + residual) vs F.conv2d in float64."""
    from stablediffusioneo_b200 import ops
    from stablediffusioneo_b200._lib import SDEO_ACT_SILU
    dev = cuda_device
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    if case == "plain3x3":
        x, w = _rand((2, 320, 16, 24), 1), _rand((320, 320, 3, 3), 2, (320 * 9) ** -0.5)
        ref = F.conv2d(x.double(), w.double(), padding=1)
        y = ops.conv2d(ops.split_terms(nhwc(x).to(dev)), ops.pack_conv_weight_split(w.to(dev)), out_fp32=True)
    elif case == "stride2":
        x, w = _rand((1, 96, 32, 48), 1), _rand((256, 96, 3, 3), 2, (96 * 9) ** -0.5)
        ref = F.conv2d(x.double(), w.double(), padding=1, stride=2)
        y = ops.conv2d(ops.split_terms(nhwc(x).to(dev)), ops.pack_conv_weight_split(w.to(dev)), stride=2, out_fp32=True)
    elif case == "concat":
        a, b, w = _rand((2, 640, 8, 12), 1), _rand((2, 320, 8, 12), 3), _rand((640, 960, 1, 1), 2, 960 ** -0.5)
        ref = F.conv2d(torch.cat([a, b], 1).double(), w.double())
        y = ops.conv2d(ops.split_terms(nhwc(a).to(dev)), ops.pack_conv_weight_split(w.to(dev), c1=640, c2=320),
                       x2=ops.split_terms(nhwc(b).to(dev)), out_fp32=True)
    elif case == "nchw4":
        x, w = _rand((2, 4, 32, 48), 1), _rand((320, 4, 3, 3), 2, 36 ** -0.5)
        ref = F.conv2d(x.double(), w.double(), padding=1)
        y = ops.conv2d(ops.split_terms(x.to(dev), nchw=True), ops.pack_conv_weight_split(w.to(dev)), out_fp32=True)
    elif case == "upsample":
        x, w = _rand((1, 640, 8, 12), 1), _rand((640, 640, 3, 3), 2, (640 * 9) ** -0.5)
        ref = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), w.double(), padding=1)
        up = ops.upsample_nearest2x(ops.split_terms(nhwc(x).to(dev)))
        y = ops.conv2d(up, ops.pack_conv_weight_split(w.to(dev)), out_fp32=True)
    else:
        x, w = _rand((2, 320, 16, 24), 1), _rand((320, 320, 3, 3), 2, (320 * 9) ** -0.5)
        b, e, r = _rand((320,), 3, 0.1), _rand((2, 320), 4, 0.3), _rand((2, 320, 16, 24), 5)
        ref = F.silu(F.conv2d(x.double(), w.double(), b.double(), padding=1) + e.double()[:, :, None, None]) * 0.7 + r.double()
        y = ops.conv2d(ops.split_terms(nhwc(x).to(dev)), ops.pack_conv_weight_split(w.to(dev)), bias=b.to(dev),
                       emb=e.to(dev), residual=nhwc(r).to(dev), scale=0.7, act=SDEO_ACT_SILU, out_fp32=True)
    err = rel_l2(y.permute(0, 3, 1, 2), ref)
    print(f"split conv {case}: rel L2 {err:.2e}")
    assert err < 3e-5


def test_precise_norms_and_activations(cuda_device):
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    # GroupNorm (+SiLU), plain and with the concat seam inside a group (960 + 320 channels, 40 per group)
    for c1, c2, silu, eps in ((320, 0, True, 1e-5), (960, 320, False, 1e-6), (32, 0, True, 1e-5)):
        a = _rand((2, 7, 9, c1), 1) * 3 + 0.5
        b = _rand((2, 7, 9, c2), 2) - 1.0 if c2 else None
        gam, bet = _rand((c1 + c2,), 3) + 1, _rand((c1 + c2,), 4)
        full = a if b is None else torch.cat([a, b], -1)
        ref = F.group_norm(full.permute(0, 3, 1, 2).double(), 32, gam.double(), bet.double(), eps)
        ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
        y = ops.groupnorm_f32(a.to(dev), gam.to(dev), bet.to(dev), eps, silu, x2=None if b is None else b.to(dev))
        assert rel_l2(y, ref) < 2e-6, (c1, c2)
    # LayerNorm
    for rows, c in ((3072, 320), (5, 1280)):
        x, gam, bet = _rand((rows, c), 5) * 2 + 1, _rand((c,), 6) + 1, _rand((c,), 7)
        ref = F.layer_norm(x.double(), (c,), gam.double(), bet.double(), 1e-5)
        assert rel_l2(ops.layernorm_f32(x.to(dev), gam.to(dev), bet.to(dev)), ref) < 2e-6
    # GEGLU, SiLU
    x = _rand((100, 2 * 640), 8) * 2
    ref = x[:, :640].double() * F.gelu(x[:, 640:].double())
    assert rel_l2(ops.geglu_f32(x.to(dev)), ref) < 2e-6
    assert rel_l2(ops.silu_f32(x.to(dev)), F.silu(x.double())) < 2e-6
    # sinusoidal embedding vs the reference formula (util.py:154-174) evaluated in float32 like the reference does
    t = torch.tensor([951, 1, 501, 0], dtype=torch.long)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    got = ops.timestep_embedding_f32(t.to(dev), 320).cpu()
    assert (got - ref).abs().max().item() < 2e-4  # fp32 argument rounding at t*f ~ 951 rad: ~6e-5 absolute
    assert rel_l2(got, ref) < 5e-5


@pytest.mark.parametrize("b,heads,nq,nkv,d", [(2, 8, 1536, 1536, 40), (2, 8, 96, 77, 160), (1, 8, 24, 24, 160),
                                              (2, 4, 50, 33, 80), (1, 2, 17, 1, 8)])
def test_precise_attention(cuda_device, b, heads, nq, nkv, d):
    from stablediffusioneo_b200 import ops
    c = heads * d
    q, k, v = _rand((b, nq, c), 1), _rand((b, nkv, c), 2), _rand((b, nkv, c), 3)
    split = lambda t: t.double().reshape(b, -1, heads, d).transpose(1, 2)
    p = torch.softmax(split(q) @ split(k).transpose(-1, -2) * d ** -0.5, -1)
    ref = (p @ split(v)).transpose(1, 2).reshape(b, nq, c)
    dev = cuda_device
    out = ops.attention_f32(q.to(dev), k.to(dev), v.to(dev), heads, d ** -0.5)
    err = rel_l2(out, ref)
    print(f"fp32 attention b={b} h={heads} nq={nq} nkv={nkv} d={d}: rel L2 {err:.2e}")
    assert err < 5e-6


def test_tiny_fp32_mode(cuda_device):
    """Tiny ControlNet+UNet in fp32 mode against the real reference modules' outputs: eps with and without control, a
    4-step sampler run (generic path: fp32 mode never takes the bf16 step engine), graded scales / only_mid_control."""
    from helpers import oracle_weights
    from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler
    dev = cuda_device
    model, g = build_control_ldm(O.TINY, O.TINY_VAE, dev), load_golden("tiny")
    model.precision = "fp32"
    ts = torch.full((1,), 951, dtype=torch.long, device=dev)
    x_T, cond, uncond = inputs_on(O.TINY, 8, 16, dev)
    errs = [rel_l2(model.apply_model(x_T, ts, cond), g["eps_c_t951"]),
            rel_l2(model.apply_model(x_T, ts, uncond), g["eps_u_t951"]),
            rel_l2(model.apply_model(x_T, ts, dict(cond, c_concat=None)), g["eps_nocontrol_t951"])]
    print("tiny fp32-mode eps rel L2 (cond, uncond, no control):", errs)
    assert max(errs) < FP32_TOL
    sampler = DDIMSampler(model)
    samples, _ = sampler.sample(g["S"], 1, (4, 8, 16), cond, verbose=False, eta=0.0, x_T=x_T,
                                unconditional_guidance_scale=9.0, unconditional_conditioning=uncond)
    err = rel_l2(samples, g["samples"])
    print("tiny fp32-mode free-running latents rel L2:", err)
    assert err < 10 * FP32_TOL
    sd_unet, sd_cn, _ = oracle_weights(O.TINY, O.TINY_VAE)
    x_c, cond_c, _ = O.make_inputs(O.TINY, 1, 8, 16)
    scales = [0.9 * (0.825 ** float(12 - i)) for i in range(13)]
    for only_mid in (False, True):
        model.control_scales, model.only_mid_control = scales, only_mid
        with torch.no_grad():
            ref = O.apply_model(sd_unet, sd_cn, O.TINY, x_c, ts.cpu(), cond_c, control_scales=scales,
                                only_mid_control=only_mid)
        assert rel_l2(model.apply_model(x_T, ts, cond), ref) < FP32_TOL, only_mid


def test_sd15_fp32_mode_eps(cuda_device):
    """BASELINE configs[1] (SD1.5, 256x384, batch 1): eps at t=951 (cond, uncond) and at t=451 along the reference's own
    trajectory, fp32 mode vs the reference's fp32 PyTorch path -- gate 1e-4."""
    import stablediffusioneo_b200.precise as precise
    dev = cuda_device
    model, g = build_control_ldm(O.SD15, O.SD15_VAE, dev), load_golden("sd15_256x384")
    model.precision = "fp32"
    x_T, cond, uncond = inputs_on(O.SD15, 32, 48, dev, hint=canny_hint())
    ts = torch.full((1,), 951, dtype=torch.long, device=dev)
    errs = (rel_l2(model.apply_model(x_T, ts, cond), g["eps_c_t951"]),
            rel_l2(model.apply_model(x_T, ts, uncond), g["eps_u_t951"]))
    print("SD1.5 fp32-mode eps rel L2 (cond, uncond), 3 terms:", errs)
    assert max(errs) < FP32_TOL
    # teacher-forced mid-trajectory step: rebuild the reference's x_t from its recorded eps calls
    sch = O.ddim_schedule(20)
    x = O.make_inputs(O.SD15, 1, 32, 48)[0]
    for i in range(10):
        e_c, e_u = g["eps_calls"][2 * i], g["eps_calls"][2 * i + 1]
        x, _ = O.ddim_update(x, e_u + 9.0 * (e_c - e_u), float(sch["alphas"][19 - i]), float(sch["alphas_prev"][19 - i]),
                             0.0, float(sch["sqrt_one_minus_alphas"][19 - i]))
    t10 = torch.full((1,), int(g["call_timesteps"][20]), dtype=torch.long, device=dev)
    err = rel_l2(model.apply_model(x.to(dev), t10, cond), g["eps_calls"][20])
    print(f"SD1.5 fp32-mode eps rel L2 at t={int(t10[0])} (teacher-forced):", err)
    assert err < FP32_TOL
    # the 6-term pattern (full fp32 operand precision) as a cross-check of where the remaining error comes from
    precise.TERMS = 6
    try:
        e6 = rel_l2(model.apply_model(x_T, ts, cond), g["eps_c_t951"])
    finally:
        precise.TERMS = 3
    print("SD1.5 fp32-mode eps rel L2 (cond), 6 terms:", e6)
    assert e6 < FP32_TOL
