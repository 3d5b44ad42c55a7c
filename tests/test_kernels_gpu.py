"""GPU parity of each C-ABI kernel against plain torch fp32 math on the same bf16-rounded inputs.
Tolerances: outputs are bf16 (8 mantissa bits) from fp32 accumulation -> relative L2 <= 4e-3 per op
(north_star's end-to-end bar for bf16 is 1e-2)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 4e-3


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def bf16r(t):
    return t.to(torch.bfloat16).float()


def gen(shape, seed, dev, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def ref_conv(x_nchw, w, bias, stride, emb=None, act=False, scale=1.0, residual=None):
    k = w.shape[-1]
    y = F.conv2d(bf16r(x_nchw), bf16r(w), bias, stride=stride, padding=1 if k == 3 else 0)
    if emb is not None:
        y = y + emb[:, :, None, None]
    if act:
        y = F.silu(y)
    y = y * scale
    if residual is not None:
        y = y + bf16r(residual)
    return y


CONV_CASES = [
    # (n, cin, cout, h, w, k, stride)   -- shapes from SURVEY Appendix A (256x384, cond+uncond batch 2)
    (2, 320, 320, 32, 48, 3, 1),
    (2, 320, 320, 32, 48, 1, 1),
    (2, 320, 320, 32, 48, 3, 2),     # Downsample
    (2, 640, 1280, 8, 12, 3, 1),
    (2, 1280, 1280, 8, 12, 3, 1),    # split-K, 96-row tiles
    (2, 1280, 1280, 4, 6, 3, 1),     # tile box spans the batch
    (1, 1280, 1280, 4, 6, 1, 1),
    (2, 8, 320, 32, 48, 3, 1),       # conv_in (4 channels padded to 8)
    (2, 320, 4, 32, 48, 3, 1),       # UNet out conv
    (1, 16, 32, 64, 96, 3, 2),       # hint block stride 2, tiny channels
    (1, 96, 96, 20, 28, 3, 1),       # ragged tiles
    (3, 64, 48, 5, 7, 3, 1),         # odd everything
    (1, 128, 3, 40, 56, 3, 1),       # VAE conv_out
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_plain(cuda_device, case):
    from stablediffusioneo_b200 import ops
    n, cin, cout, h, w, k, stride = case
    dev = cuda_device
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, k, k), 2, dev, scale=1.0 / math.sqrt(cin * k * k))
    b = gen((cout,), 3, dev)
    pw = ops.pack_conv_weight(wt)
    y = ops.conv2d(nhwc(x), pw, bias=b, stride=stride)
    ref = ref_conv(x, wt, b, stride)
    assert y.shape == (n, ref.shape[2], ref.shape[3], cout)
    err = rel_l2(y.permute(0, 3, 1, 2), ref)
    assert err < TOL, f"rel L2 {err}"


HALO_CASES = [
    # (n, c1, c2, cout, h, w)  3x3 stride 1: the HALO tiling (one A tile per channel chunk serves the nine taps)
    (2, 320, 0, 320, 32, 48),      # 14 tiles of 5 x 24 pixels, pitch 26
    (2, 640, 320, 320, 32, 48),    # fused concat: chunks from two tensor maps
    (2, 640, 0, 640, 16, 24),
    (2, 1280, 0, 1280, 8, 12),     # one tile per sample; split-K over (chunk, tap) steps
    (1, 96, 0, 96, 20, 28),        # ragged tiles: halo boxes hang over the right / bottom edge
    (1, 64, 0, 128, 64, 64),
    (1, 128, 0, 64, 40, 100),      # pitch 102: one image row per tile
    (1, 24, 0, 48, 13, 9),         # partial channel chunk (24 of 64), tiny map
]


@pytest.mark.parametrize("splits", [0, 1, 2, 3, 8])
@pytest.mark.parametrize("case", HALO_CASES)
def test_conv2d_halo_tiles(cuda_device, case, splits):
    """SDEO_HALO=1 forces the halo tiling (0 forces the tap-by-tap tiling), SDEO_PAIR=1 forces CTA pairs (cta_group::2:
    two CTAs run one M=256 MMA, each staging half of the weight tile; an odd M-tile count adds a null tile): all four
    combinations against F.conv2d on the bf16-rounded operands and against each other, plain bf16 output and the full fp32 epilogue (bias, residual, twin, GroupNorm
    statistics). splits = forced K slices (cluster split-K; slices start in the middle of a channel chunk)."""
    import os
    from stablediffusioneo_b200 import ops
    n, c1, c2, cout, h, w = case
    dev = cuda_device
    cin = c1 + c2
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, 3, 3), 2, dev, scale=1.0 / math.sqrt(cin * 9))
    b = gen((cout,), 3, dev)
    res = gen((n, h, w, cout), 5, dev)
    pw = ops.pack_conv_weight(wt, c1=c1, c2=c2) if c2 else ops.pack_conv_weight(wt)
    xa = nhwc(x[:, :c1])
    xb = nhwc(x[:, c1:]) if c2 else None
    ref = ref_conv(x, wt, b, 1)
    ref32 = ref + res.permute(0, 3, 1, 2)
    got = {}
    try:
        if splits:
            os.environ["SDEO_FORCE_SPLITS"] = str(splits)
        for halo, pair in (("1", "1"), ("1", "0"), ("0", "1"), ("0", "0")):
            os.environ["SDEO_HALO"], os.environ["SDEO_PAIR"] = halo, pair
            if pair == "1" and (splits > 4 or n * ((h * w + 127) // 128) < 2):
                continue  # a CTA pair leaves room for 4 K slices in a cluster of 8; needs two M tiles
            try:
                y = ops.conv2d(xa, pw, x2=xb, bias=b)
                ys, yt = ops.conv2d(xa, pw, x2=xb, bias=b, residual=res, out_fp32=True, twin=True, gn_stats=True)
            except Exception as ex:  # split count does not fit shared memory for this shape
                assert splits > 1, ex
                pytest.skip(f"splits={splits} does not fit: {ex}")
            torch.cuda.synchronize()
            assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL, (halo, pair)
            assert rel_l2(ys.permute(0, 3, 1, 2), ref32) < 2e-3, (halo, pair)
            assert torch.equal(yt, ys.to(torch.bfloat16))
            st = getattr(ys, "_gn_stats", None)
            if cout % 32 == 0:
                assert st is not None
                gamma, beta = gen((cout,), 6, dev) * 0.2 + 1.0, gen((cout,), 7, dev) * 0.2
                out = ops.groupnorm(ys, gamma, beta, 1e-5, True, stats=st)
                gref = F.silu(F.group_norm(ys.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
                assert rel_l2(out.permute(0, 3, 1, 2), gref) < TOL, (halo, pair)
            got[halo + pair] = ys
    finally:
        os.environ.pop("SDEO_HALO", None)
        os.environ.pop("SDEO_PAIR", None)
        os.environ.pop("SDEO_FORCE_SPLITS", None)
    base = got["00"]
    for k_, v_ in got.items():
        assert rel_l2(v_, base) < 1e-5, k_   # same products, different fp32 summation order


def test_conv2d_full_epilogue(cuda_device):
    """bias + per-sample time-embedding term + SiLU + scale + residual, fp32 and bf16 outputs."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    n, cin, cout, h, w = 2, 320, 640, 16, 24
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, 3, 3), 2, dev, scale=1.0 / math.sqrt(cin * 9))
    b = gen((cout,), 3, dev)
    emb = gen((n, cout), 4, dev)
    res = gen((n, cout, h, w), 5, dev)
    pw = ops.pack_conv_weight(wt)
    for act in (False, True):
        ref = ref_conv(x, wt, b, 1, emb=emb, act=act, scale=0.7, residual=res)
        y = ops.conv2d(nhwc(x), pw, bias=b, emb=emb.contiguous(), residual=nhwc(res), scale=0.7, act=1 if act else 0)
        assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL
        y32 = ops.conv2d(nhwc(x), pw, bias=b, emb=emb.contiguous(), residual=nhwc(res), scale=0.7, act=1 if act else 0,
                         out_fp32=True)
        assert y32.dtype == torch.float32
        assert rel_l2(y32.permute(0, 3, 1, 2), ref) < 2e-3
        # fp32 residual stream: fp32 residual in, fp32 out + bf16 twin
        res32 = res.permute(0, 2, 3, 1).contiguous()
        ref32 = ref_conv(x, wt, b, 1, emb=emb, act=act, scale=0.7) + res
        ys, yt = ops.conv2d(nhwc(x), pw, bias=b, emb=emb.contiguous(), residual=res32, scale=0.7, act=1 if act else 0,
                            out_fp32=True, twin=True)
        assert ys.dtype == torch.float32 and yt.dtype == torch.bfloat16
        assert rel_l2(ys.permute(0, 3, 1, 2), ref32) < 2e-3
        assert torch.equal(yt, ys.to(torch.bfloat16))


@pytest.mark.parametrize("c1,c2,cout,h,w,k", [(640, 320, 320, 32, 48, 3), (1280, 640, 1280, 8, 12, 1),
                                               (1280, 1280, 1280, 4, 6, 3)])
def test_conv2d_fused_concat(cuda_device, c1, c2, cout, h, w, k):
    """torch.cat([h, skip], 1) -> conv (cldm/cldm.py:41 + ResBlock) as a dual-source K loop."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    xa, xb = gen((2, c1, h, w), 1, dev), gen((2, c2, h, w), 2, dev)
    wt = gen((cout, c1 + c2, k, k), 3, dev, scale=1.0 / math.sqrt((c1 + c2) * k * k))
    b = gen((cout,), 4, dev)
    pw = ops.pack_conv_weight(wt, c1=c1, c2=c2)
    y = ops.conv2d(nhwc(xa), pw, x2=nhwc(xb), bias=b)
    ref = ref_conv(torch.cat([xa, xb], 1), wt, b, 1)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL


@pytest.mark.parametrize("rows,kdim,ndim", [(3072, 320, 320), (768, 640, 640), (192, 1280, 1280), (48, 1280, 1280),
                                            (2, 320, 1280), (2, 1280, 1280), (154, 768, 640), (77, 768, 2560)])
def test_linear(cuda_device, rows, kdim, ndim):
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    x = gen((rows, kdim), 1, dev)
    wt = gen((ndim, kdim), 2, dev, scale=1.0 / math.sqrt(kdim))
    b = gen((ndim,), 3, dev)
    res = gen((rows, ndim), 4, dev)
    pw = ops.pack_conv_weight(wt)
    y = ops.linear(x.to(torch.bfloat16), pw, bias=b, residual=res.to(torch.bfloat16))
    ref = F.linear(bf16r(x), bf16r(wt), b) + bf16r(res)
    assert rel_l2(y, ref) < TOL


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("rows,kdim,ndim", [(384, 320, 320), (640, 640, 1280), (1000, 96, 48), (3072, 1280, 320)])
def test_linear_cta_pairs(cuda_device, rows, kdim, ndim, pair):
    """Linears under forced CTA pairs (cta_group::2), including ODD M-tile counts (384 rows = 3 tiles, 640 = 5, 1000 = 8
    with a ragged last tile): the pair's null tile loads nothing but zeros and writes nothing. Plain, residual-stream
    (fp32 + twin + row statistics) and GEGLU epilogues."""
    import os
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    x = gen((rows, kdim), 1, dev)
    wt = gen((ndim, kdim), 2, dev, scale=1.0 / math.sqrt(kdim))
    b = gen((ndim,), 3, dev)
    res = gen((rows, ndim), 4, dev)
    xb = x.to(torch.bfloat16)
    ref = bf16r(x) @ bf16r(wt).t() + b
    os.environ["SDEO_PAIR"] = pair
    try:
        y = ops.linear(xb, ops.pack_conv_weight(wt), bias=b)
        assert rel_l2(y, ref) < TOL
        ys, yt = ops.linear(xb, ops.pack_conv_weight(wt), bias=b, residual=res, out_fp32=True, twin=True,
                            row_stats=ndim % 64 == 0)   # row statistics need an N tile of 64 / 128 / 256 columns
        assert rel_l2(ys, ref + res) < 2e-3 and torch.equal(yt, ys.to(torch.bfloat16))
        rs = getattr(ys, "_row_stats", None)
        if rs is not None:
            buf, parts, nrows = rs
            tot = buf[:parts].sum(0)
            assert torch.allclose(tot[:, 0], ys.sum(1), rtol=1e-3, atol=1e-2)
        if ndim % 64 == 0:
            w2 = gen((2 * ndim, kdim), 5, dev, scale=1.0 / math.sqrt(kdim))
            b2 = gen((2 * ndim,), 6, dev)
            pw = ops.pack_conv_weight(w2, geglu=True)
            yg = ops.linear(xb, pw, bias=ops.pack_geglu_bias(b2, pw.geglu_bn), geglu=True)
            full = bf16r(x) @ bf16r(w2).t() + b2
            gref = full[:, :ndim] * F.gelu(full[:, ndim:])
            assert rel_l2(yg, gref) < TOL
    finally:
        os.environ.pop("SDEO_PAIR", None)


@pytest.mark.parametrize("case", [(2, 128, 128, 64, 64, 3), (2, 320, 320, 48, 96, 3), (4, 256, 128, 40, 40, 1), (1, 640, 640, 96, 96, 3)])
def test_conv2d_plan_switches(cuda_device, case):
    """The plan / pipeline switches of conv_gemm_kernel on multi-wave grids: two CTAs per SM (SDEO_OCC2: <= 112 KB of shared
    memory, the 80-register entry point), the MMA issuer's barrier probe (SDEO_NO_PROBE), the L2 weight prefetch
    (SDEO_L2_PREFETCH) and the residual TMA prefetch (SDEO_NO_RES_PREFETCH): every combination against F.conv2d on the
    bf16-rounded operands, plain bf16 output and the fp32 + twin + residual + GroupNorm-statistics epilogue."""
    import os
    from stablediffusioneo_b200 import ops
    n, cin, cout, h, w, k = case
    dev = cuda_device
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, k, k), 2, dev, scale=1.0 / math.sqrt(cin * k * k))
    b = gen((cout,), 3, dev)
    res = gen((n, h, w, cout), 5, dev)
    pw = ops.pack_conv_weight(wt)
    ref = ref_conv(x, wt, b, 1)
    ref32 = ref + res.permute(0, 3, 1, 2)
    keys = ("SDEO_OCC2", "SDEO_NO_PROBE", "SDEO_L2_PREFETCH", "SDEO_NO_RES_PREFETCH")
    base = None
    try:
        for occ2 in ("0", "1"):
            for extra in ({}, {"SDEO_NO_PROBE": "1"}, {"SDEO_L2_PREFETCH": "1"}, {"SDEO_NO_RES_PREFETCH": "1"}):
                for k_ in keys:
                    os.environ.pop(k_, None)
                os.environ["SDEO_OCC2"] = occ2
                os.environ.update(extra)
                y = ops.conv2d(nhwc(x), pw, bias=b)
                ys, yt = ops.conv2d(nhwc(x), pw, bias=b, residual=res, out_fp32=True, twin=True, gn_stats=True)
                torch.cuda.synchronize()
                assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL, (occ2, extra)
                assert rel_l2(ys.permute(0, 3, 1, 2), ref32) < 2e-3, (occ2, extra)
                assert torch.equal(yt, ys.to(torch.bfloat16))
                if base is None:
                    base = ys
                assert rel_l2(ys, base) < 1e-5, (occ2, extra)
    finally:
        for k_ in keys:
            os.environ.pop(k_, None)


@pytest.mark.parametrize("rows,c", [(3072, 320), (768, 640), (192, 1280), (48, 1280)])
def test_linear_geglu(cuda_device, rows, c):
    """GEGLU (attention.py:49-56): proj -> chunk(2) -> x * gelu(gate), fused into the projection epilogue."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    inner = 4 * c
    x = gen((rows, c), 1, dev)
    wt = gen((2 * inner, c), 2, dev, scale=1.0 / math.sqrt(c))
    b = gen((2 * inner,), 3, dev, scale=0.5)
    pw = ops.pack_conv_weight(wt, geglu=True)
    bp = ops.pack_geglu_bias(b, pw.geglu_bn)
    y = ops.linear(x.to(torch.bfloat16), pw, bias=bp, geglu=True)
    h = F.linear(bf16r(x), bf16r(wt), b)
    a, g = h.chunk(2, dim=-1)
    ref = a * F.gelu(g)
    assert y.shape == (rows, inner)
    assert rel_l2(y, ref) < TOL


@pytest.mark.parametrize("b,t,c,heads", [(2, 1536, 320, 8), (2, 384, 640, 8), (2, 96, 1280, 8), (2, 24, 1280, 8)])
def test_qkv_projection_layout(cuda_device, b, t, c, heads):
    """Fused qkv_w = cat([Wq, Wk, Wv]).T (attention.py:170,193-194) with the head split (attention.py:227)."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    d = c // heads
    x = gen((b, t, c), 1, dev)
    wq, wk, wv = (gen((c, c), s, dev, scale=1.0 / math.sqrt(c)) for s in (2, 3, 4))
    pw = ops.pack_conv_weight(torch.cat([wq, wk, wv], 0))
    q = torch.empty((b * heads, t, d), dtype=torch.bfloat16, device=dev)
    k = torch.empty_like(q)
    vt = torch.empty((b * heads, d, t), dtype=torch.bfloat16, device=dev)
    ops.qkv_project(x.to(torch.bfloat16), pw, heads, d, 0, q=q, k=k, vt=vt, ldv=t)
    xr = bf16r(x)

    def split(wm):
        y = F.linear(xr, bf16r(wm))  # [b, t, c]
        return y.view(b, t, heads, d).permute(0, 2, 1, 3).reshape(b * heads, t, d)

    assert rel_l2(q, split(wq)) < TOL
    assert rel_l2(k, split(wk)) < TOL
    assert rel_l2(vt, split(wv).transpose(1, 2)) < TOL


GN_CASES = [
    # (n, c1, c2, h, w, eps, silu)
    (2, 320, 0, 32, 48, 1e-5, True),
    (2, 320, 0, 32, 48, 1e-6, False),
    (2, 640, 320, 32, 48, 1e-5, True),     # C=960: concat seam inside group 21
    (2, 1280, 640, 16, 24, 1e-5, True),    # C=1920: concat seam inside group 21
    (2, 1280, 1280, 8, 12, 1e-5, True),    # C=2560 (> 256 channel vectors)
    (2, 1280, 0, 4, 6, 1e-5, True),
    (1, 128, 0, 64, 96, 1e-6, True),       # VAE: 4 channels per group
    (3, 512, 0, 9, 7, 1e-6, False),
    (2, 128, 0, 384, 320, 1e-6, True),     # big-tensor grid (two waves of resident CTAs), pipelined row batches
    (3, 64, 0, 301, 277, 1e-6, False),     # same, ragged: chunk tails shorter than a batch of rows
    (5, 64, 32, 250, 200, 1e-5, True),     # same, over a concat with 3 channels per group
]


@pytest.mark.parametrize("case", GN_CASES)
def test_groupnorm(cuda_device, case):
    from stablediffusioneo_b200 import ops
    n, c1, c2, h, w, eps, silu = case
    dev = cuda_device
    xa = gen((n, c1, h, w), 1, dev) * 1.7 + 0.3
    xb = gen((n, c2, h, w), 2, dev) * 0.6 - 0.5 if c2 else None
    gamma = gen((c1 + c2,), 3, dev) * 0.2 + 1.0
    beta = gen((c1 + c2,), 4, dev) * 0.2
    y = ops.groupnorm(nhwc(xa), gamma, beta, eps, silu, x2=nhwc(xb) if c2 else None)
    xcat = bf16r(torch.cat([xa, xb], 1) if c2 else xa)
    ref = F.group_norm(xcat, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL
    # fp32 (residual-stream) inputs
    f32 = lambda t: t.permute(0, 2, 3, 1).contiguous()
    y32 = ops.groupnorm(f32(xa), gamma, beta, eps, silu, x2=f32(xb) if c2 else None)
    ref32 = F.group_norm(torch.cat([xa, xb], 1) if c2 else xa, 32, gamma, beta, eps)
    assert rel_l2(y32.permute(0, 3, 1, 2), F.silu(ref32) if silu else ref32) < TOL


def test_groupnorm_repeated_and_deterministic(cuda_device):
    """The single-launch variant synchronises its CTAs through arrival counters that live in the workspace across
    calls: back-to-back calls (different shapes sharing the workspace) must stay correct and bit-identical."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    outs = []
    for it in range(6):
        for (n, c, h, w) in ((2, 320, 32, 48), (2, 1280, 4, 6), (1, 640, 16, 24)):
            x = (gen((n, c, h, w), 10 + c, dev) * 1.3 + 0.2).permute(0, 2, 3, 1).contiguous()
            gamma = gen((c,), 3, dev) * 0.2 + 1.0
            beta = gen((c,), 4, dev) * 0.2
            y = ops.groupnorm(x, gamma, beta, 1e-5, True)
            if it == 0:
                ref = F.silu(F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
                assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL
                outs.append(y.clone())
            else:
                assert torch.equal(y, outs[len(outs) - 3 + [320, 1280, 640].index(c)] if False else outs[[320, 1280, 640].index(c)])


@pytest.mark.parametrize("autotune", [False, True])
@pytest.mark.parametrize("case", [
    # (n, cin, cout, h, w, k, cat)  cat: GroupNorm over the concat of two conv outputs (decoder blocks)
    (2, 320, 320, 32, 48, 3, False),     # 12 M tiles per sample
    (2, 640, 640, 16, 24, 1, False),
    (2, 1280, 1280, 8, 12, 3, True),     # split-K cluster, 96-row tiles; concat 2560 channels (80 per group)
    (2, 640, 320, 32, 48, 3, True),      # concat 640 channels
    (1, 96, 96, 20, 28, 3, False),       # ragged tiles, 3 channels per group
    (2, 1280, 1280, 4, 6, 3, False),     # tile box spans the batch: statistics per K-slice rank (or standalone pass)
    (2, 2560, 1280, 4, 6, 3, True),
    (1, 1280, 1280, 1, 2, 3, False),     # 2 output rows, up to 8 K slices: ranks without rows own zero slots
])
def test_groupnorm_from_conv_statistics(cuda_device, case, autotune):
    """conv epilogue -> per-channel partial statistics -> GroupNorm that reads the tensor once, against
    F.group_norm of the very tensor the conv wrote (and against the standalone GroupNorm kernel)."""
    from stablediffusioneo_b200 import ops
    n, cin, cout, h, w, k, cat = case
    dev = cuda_device
    ops.set_autotune(autotune)
    try:
        outs = []
        for seed in ((1, 2) if cat else (1,)):
            x = gen((n, cin, h, w), seed, dev)
            wt = gen((cout, cin, k, k), 10 + seed, dev) / math.sqrt(cin * k * k)
            bias = gen((cout,), 20 + seed, dev) * 0.1
            res = gen((n, h, w, cout), 30 + seed, dev)
            y, _ = ops.conv2d(nhwc(x), ops.pack_conv_weight(wt), bias=bias, residual=res, out_fp32=True, twin=True,
                              gn_stats=True)
            outs.append(y)
    finally:
        ops.set_autotune(False)
    spans_batch = h * w * n <= 128 and n > 1   # statistics then need K slices whose row ranges stay inside a sample
    for y in outs:
        assert spans_batch or getattr(y, "_gn_stats", None) is not None
    ctot = cout * len(outs)
    gamma = gen((ctot,), 3, dev) * 0.2 + 1.0
    beta = gen((ctot,), 4, dev) * 0.2
    st = [getattr(y, "_gn_stats", None) for y in outs]
    x2 = outs[1] if cat else None
    y_fused = ops.groupnorm(outs[0], gamma, beta, 1e-5, True, x2=x2, stats=st[0], stats2=st[1] if cat else None)
    y_plain = ops.groupnorm(outs[0], gamma, beta, 1e-5, True, x2=x2)
    full = torch.cat(outs, 3).permute(0, 3, 1, 2)
    ref = F.silu(F.group_norm(full, 32, gamma, beta, 1e-5))
    assert rel_l2(y_fused.permute(0, 3, 1, 2), ref) < TOL
    assert rel_l2(y_fused, y_plain) < 2e-3   # same math, different fp32 summation order (+ bf16 rounding flips)


def test_conv2d_pair_null_tile_with_emb(cuda_device):
    """CTA pairs over an ODD number of M tiles (n = 1, 16x24 -> three 128-pixel tiles + one null tile) with a per-sample
    time-embedding row: the null tile lies beyond the last sample and must not read emb[n] (regression: it read 4 * cout
    bytes past the end of emb -- an illegal address when emb ends a mapped block, as arranged here)."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    n, c, h, w = 1, 640, 16, 24
    x = gen((n, c, h, w), 1, dev)
    wt = gen((c, c, 3, 3), 2, dev) / math.sqrt(c * 9)
    bias = gen((c,), 3, dev) * 0.1
    block = torch.empty((20 << 20) // 4, dtype=torch.float32, device=dev)   # one allocator block; emb = its last row
    emb = block[-c:].view(1, c)
    emb.copy_(gen((1, c), 4, dev))
    os.environ["SDEO_PAIR"] = "1"
    try:
        y = ops.conv2d(nhwc(x), ops.pack_conv_weight(wt), bias=bias, emb=emb, out_fp32=True, gn_stats=True)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("SDEO_PAIR", None)
    ref = ref_conv(x, wt, bias, 1, emb=emb)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL


@pytest.mark.parametrize("case", [
    # (n, cin, cout, h, w)
    (1, 512, 512, 32, 48),      # VAE decoder, first Upsample at 256x384
    (2, 256, 256, 64, 64),
    (1, 64, 48, 13, 9),         # ragged tiles, odd sizes
    (3, 128, 128, 20, 28),
])
def test_upsample2x_conv_subpixel(cuda_device, case):
    """conv3x3(nearest_x2(x)) as four 2x2 phase convolutions over the low-resolution input (sdeo_conv_args::up2_phase)
    against F.interpolate + F.conv2d in fp32 on the bf16-rounded input (weights: fp32 sums rounded once, so the gate is the
    per-op one), and against the library's own upsample kernel + 3x3 conv."""
    from stablediffusioneo_b200 import ops
    n, cin, cout, h, w = case
    dev = cuda_device
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, 3, 3), 2, dev) / math.sqrt(cin * 9)
    bias = gen((cout,), 3, dev) * 0.1
    phases = ops.upsample2x_conv_weights(wt)
    y = ops.upsample2x_conv(nhwc(x), phases, bias=bias)
    assert y.shape == (n, 2 * h, 2 * w, cout)
    ref = F.conv2d(F.interpolate(bf16r(x), scale_factor=2, mode="nearest"), wt, bias, padding=1)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL
    y2 = ops.conv2d(ops.upsample_nearest2x(nhwc(x)), ops.pack_conv_weight(wt), bias=bias)
    assert rel_l2(y, y2) < TOL


GNF_CASES = [
    # (n, c_a, c_b, cout, h, w, k, silu, halo)   producer convs -> GroupNorm(+SiLU) folded into the consumer conv
    (2, 320, 0, 320, 32, 48, 3, True, 1),       # ResBlock in_layers at the top level, HALO tiling
    (2, 320, 0, 320, 32, 48, 3, True, 0),       # ... tap-by-tap tiling (padding rows depend on the tap)
    (2, 320, 0, 320, 32, 48, 1, False, 0),      # SpatialTransformer norm -> proj_in (no SiLU)
    (2, 640, 320, 320, 32, 48, 3, True, 1),     # decoder block: GroupNorm over a concat (30 channels per group), HALO
    (2, 640, 320, 320, 32, 48, 3, True, 0),
    (2, 1280, 1280, 1280, 8, 12, 3, True, 0),   # split-K cluster, 96-row tiles, concat 2560 channels
    (2, 1280, 0, 1280, 4, 6, 3, True, 0),       # tile box spans the batch: one table row per sample of the tile
    (2, 1280, 1280, 1280, 4, 6, 3, True, 0),
    (1, 96, 0, 96, 20, 28, 3, True, 1),         # ragged tiles, 3 channels per group, K padding (96 -> 128)
    (1, 96, 0, 96, 20, 28, 3, True, 0),
    (3, 64, 0, 48, 5, 7, 3, True, 0),           # odd everything
    (1, 128, 0, 128, 64, 96, 3, True, 1),       # VAE ResnetBlock shape
    (2, 320, 0, 4, 32, 48, 3, True, 0),         # UNet out: GroupNorm + SiLU + conv to 4 channels (masked epilogue)
]


@pytest.mark.parametrize("stream_in", [True, False])
@pytest.mark.parametrize("case", GNF_CASES)
def test_conv2d_folded_groupnorm(cuda_device, case, stream_in):
    """producer conv(s) (epilogue statistics) -> consumer conv with the GroupNorm (+ SiLU) folded into its operand path,
    against F.group_norm + F.silu + F.conv2d in fp32 on the very bf16 tensor(s) the consumer reads (the fp32 stream's
    bf16 twin, or a bf16 output). The normalised operand is bf16 in both (the reference rounds it like the kernel does)."""
    from stablediffusioneo_b200 import ops
    n, ca, cb, cout, h, w, k, silu, halo = case
    dev = cuda_device
    srcs = []
    for i, c in enumerate([ca, cb] if cb else [ca]):
        x = gen((n, c, h, w), 1 + i, dev)
        pk = 3 if n * h * w <= 128 else 1   # (tiny maps: a tile spans samples; only a split-K producer leaves statistics)
        wt = gen((c, c, pk, pk), 10 + i, dev) / math.sqrt(c * pk * pk)
        bias = gen((c,), 20 + i, dev) * 0.3
        if stream_in:
            y, twin = ops.conv2d(nhwc(x), ops.pack_conv_weight(wt), bias=bias, out_fp32=True, twin=True, gn_stats=True)
            st = getattr(y, "_gn_stats", None)
        else:
            twin = ops.conv2d(nhwc(x), ops.pack_conv_weight(wt), bias=bias, gn_stats=True)
            st = getattr(twin, "_gn_stats", None)
        if st is None:
            pytest.skip("this producer geometry leaves no statistics (M tile spans samples without K slices)")
        srcs.append((twin, st))
    ctot = ca + cb
    gamma = gen((ctot,), 3, dev) * 0.2 + 1.0
    beta = gen((ctot,), 4, dev) * 0.2
    wt = gen((cout, ctot, k, k), 5, dev) / math.sqrt(ctot * k * k)
    bias = gen((cout,), 6, dev) * 0.1
    pw = ops.pack_conv_weight(wt, c1=ca, c2=cb) if cb else ops.pack_conv_weight(wt)
    gnf = ops.GnFold(srcs[0][1], srcs[1][1] if cb else None, gamma, beta, 32, 1e-5, silu)
    os.environ["SDEO_HALO"] = str(halo)
    try:
        y = ops.conv2d(srcs[0][0], pw, x2=srcs[1][0] if cb else None, bias=bias, out_fp32=True, gnf=gnf)
    finally:
        os.environ.pop("SDEO_HALO", None)
    full = torch.cat([t.float() for t, _ in srcs], 3).permute(0, 3, 1, 2)
    hn = F.group_norm(full, 32, gamma, beta, 1e-5)
    if silu:
        hn = F.silu(hn)
    ref = F.conv2d(bf16r(hn), bf16r(wt), bias, padding=1 if k == 3 else 0)
    err = rel_l2(y.permute(0, 3, 1, 2), ref)
    assert err < TOL, err
    # and the same through the standalone GroupNorm kernel + plain conv (what the fold replaces)
    hn2 = ops.groupnorm(srcs[0][0], gamma, beta, 1e-5, silu, x2=srcs[1][0] if cb else None)
    pw1 = ops.pack_conv_weight(wt)
    y2 = ops.conv2d(hn2, pw1, bias=bias, out_fp32=True)
    assert rel_l2(y, y2) < 3e-3


def test_conv2d_folded_groupnorm_many_parts(cuda_device):
    """A large feature map leaves hundreds of partial slots per sample: ops.gn_stats_fold reduces them first."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    n, c, h, w = 2, 128, 192, 256
    x = gen((n, c, h, w), 1, dev)
    wt = gen((c, c, 3, 3), 10, dev) / math.sqrt(c * 9)
    t = ops.conv2d(nhwc(x), ops.pack_conv_weight(wt), gn_stats=True)
    st = t._gn_stats
    assert st[1] > ops.GN_FOLD_MAX_PARTS
    st2 = ops.gn_stats_fold(st, n, c)
    assert st2[1] <= ops.GN_FOLD_MAX_PARTS
    full = st[0][:n * st[1]].reshape(n, st[1], c, 2).sum(1)
    assert torch.allclose(st2[0][:n * st2[1]].reshape(n, st2[1], c, 2).sum(1), full, rtol=1e-5, atol=1e-3)
    gamma = gen((c,), 3, dev) * 0.2 + 1.0
    beta = gen((c,), 4, dev) * 0.2
    w2 = gen((c, c, 3, 3), 5, dev) / math.sqrt(c * 9)
    y = ops.conv2d(t, ops.pack_conv_weight(w2), gnf=ops.GnFold(st2, None, gamma, beta, 32, 1e-6, True))
    hn = F.silu(F.group_norm(t.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6))
    ref = F.conv2d(bf16r(hn), bf16r(w2), None, padding=1)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < TOL


def test_groupnorm_statistics_every_tile_config(cuda_device):
    """The fused statistics under every (N tile, K slices) configuration the planner / autotuner can pick, forced
    through SDEO_FORCE_BN / SDEO_FORCE_SPLITS (N tile 256 carries 512 statistics values for 384 threads)."""
    import os
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    n, cin, cout, h, w, k = 2, 1280, 1280, 8, 12, 3
    x = gen((n, cin, h, w), 1, dev)
    wt = gen((cout, cin, k, k), 11, dev) / math.sqrt(cin * k * k)
    res = gen((n, h, w, cout), 31, dev)
    pw = ops.pack_conv_weight(wt)
    gamma = gen((cout,), 3, dev) * 0.2 + 1.0
    beta = gen((cout,), 4, dev) * 0.2
    checked = 0
    try:
        for bn in (64, 80, 128, 160, 256):
            for sp in (1, 2, 4, 8):
                os.environ["SDEO_FORCE_BN"], os.environ["SDEO_FORCE_SPLITS"] = str(bn), str(sp)
                try:
                    y, _ = ops.conv2d(nhwc(x), pw, residual=res, out_fp32=True, twin=True, gn_stats=True)
                except Exception:
                    continue  # configuration does not fit shared memory
                st = getattr(y, "_gn_stats", None)
                assert st is not None
                out = ops.groupnorm(y, gamma, beta, 1e-5, True, stats=st)
                ref = F.silu(F.group_norm(y.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
                assert rel_l2(out.permute(0, 3, 1, 2), ref) < TOL, (bn, sp)
                checked += 1
    finally:
        os.environ.pop("SDEO_FORCE_BN", None)
        os.environ.pop("SDEO_FORCE_SPLITS", None)
    assert checked >= 12


@pytest.mark.parametrize("rows,c,geglu", [(3072, 320, False), (768, 640, True), (192, 1280, False), (48, 1280, True),
                                          (24, 1280, False)])
def test_layernorm_folded_into_gemm(cuda_device, rows, c, geglu):
    """Producer GEMM (fp32 stream + bf16 twin + per-row statistics) -> consumer GEMM on the RAW twin with the LayerNorm
    folded (gamma in the weight, beta in the bias, mean / rstd correction in the epilogue), against
    F.layer_norm(fp32 stream) @ W in fp32 and against the standalone LayerNorm kernel + plain GEMM."""
    from stablediffusioneo_b200 import ops
    from stablediffusioneo_b200.ldm.modules.diffusionmodules import util
    dev = cuda_device
    xin = gen((rows, c), 1, dev).to(torch.bfloat16)
    wp = gen((c, c), 2, dev) / math.sqrt(c)
    res = gen((rows, c), 3, dev) * 1.5 + 0.4          # rows with a clearly non-zero mean
    y, y16 = ops.linear(xin, ops.pack_conv_weight(wp), residual=res, out_fp32=True, twin=True, row_stats=True)
    assert getattr(y, "_row_stats", None) is not None
    ln = torch.nn.LayerNorm(c).to(dev)
    with torch.no_grad():
        ln.weight.copy_(gen((c,), 4, dev) * 0.2 + 1.0)
        ln.bias.copy_(gen((c,), 5, dev) * 0.2)
    nout = 2 * c if geglu else 3 * c
    w = gen((nout, c), 6, dev) / math.sqrt(c)
    bias = gen((nout,), 7, dev) * 0.1
    pw, bp, csum = util.fold_ln_weight(w, bias, ln, geglu=geglu)
    buf, parts, nrows = y._row_stats
    fold = ops.LnFold(buf, parts, nrows, c, ln.eps, csum)
    out = ops.linear(y16, pw, bias=bp, geglu=geglu, ln=fold, out_fp32=not geglu)
    z = F.layer_norm(y, (c,), ln.weight, ln.bias, ln.eps) @ bf16r(w).t() + bias
    ref = z[:, :c] * F.gelu(z[:, c:]) if geglu else z
    # the raw rows (mean 0.2 sigma here) are rounded to bf16 BEFORE the mean is removed: slightly above the per-op TOL
    assert rel_l2(out, ref) < 1.5 * TOL
    # the unfused path on the same data
    lnx = ops.layernorm(y, ln.weight.detach(), ln.bias.detach(), ln.eps)
    if geglu:
        pw0 = ops.pack_conv_weight(w, geglu=True)
        plain = ops.linear(lnx, pw0, bias=ops.pack_geglu_bias(bias, pw0.geglu_bn), geglu=True)
    else:
        plain = ops.linear(lnx, ops.pack_conv_weight(w), bias=bias, out_fp32=True)
    assert rel_l2(out, plain) < 6e-3


@pytest.mark.parametrize("rows,c", [(3072, 320), (768, 640), (192, 1280), (48, 1280), (5, 512)])
def test_layernorm(cuda_device, rows, c):
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    x = gen((rows, c), 1, dev) * 2.0 + 0.5
    gamma = gen((c,), 2, dev) * 0.2 + 1.0
    beta = gen((c,), 3, dev) * 0.2
    y = ops.layernorm(x.to(torch.bfloat16), gamma, beta, 1e-5)
    ref = F.layer_norm(bf16r(x), (c,), gamma, beta, 1e-5)
    assert rel_l2(y, ref) < TOL
    y32 = ops.layernorm(x.contiguous(), gamma, beta, 1e-5)  # fp32 residual-stream input
    assert y32.dtype == torch.bfloat16 and rel_l2(y32, F.layer_norm(x, (c,), gamma, beta, 1e-5)) < TOL


ATT_CASES = [
    # (batch, heads, nq, nkv, d)
    (2, 8, 1536, 1536, 40),
    (2, 8, 384, 384, 80),
    (2, 8, 96, 96, 160),
    (2, 8, 24, 24, 160),
    (2, 8, 1536, 77, 40),   # cross-attention, 77-token CLIP context
    (2, 8, 384, 77, 80),
    (2, 8, 96, 77, 160),
    (2, 8, 24, 77, 160),
    (1, 8, 576, 576, 80),   # 768x768: ragged query tile (576 = 4.5 x 128)
    (1, 2, 200, 333, 40),   # ragged everything
]


@pytest.mark.parametrize("case", ATT_CASES)
def test_attention(cuda_device, case):
    """softmax(q k^T d^-1/2) v with fp32 scores (attention.py:227-249)."""
    from stablediffusioneo_b200 import ops
    b, heads, nq, nkv, d = case
    dev = cuda_device
    q = gen((b * heads, nq, d), 1, dev)
    k = gen((b * heads, nkv, d), 2, dev)
    v = gen((b * heads, nkv, d), 3, dev)
    ldv = (nkv + 7) // 8 * 8
    vt = torch.zeros((b * heads, d, ldv), dtype=torch.bfloat16, device=dev)
    vt[:, :, :nkv] = v.transpose(1, 2).to(torch.bfloat16)
    scale = d ** -0.5
    o = ops.attention(q.to(torch.bfloat16), k.to(torch.bfloat16), vt, b, heads, nq, nkv, d, ldv, scale)
    sim = torch.einsum("bid,bjd->bij", bf16r(q), bf16r(k)) * scale
    ref = torch.einsum("bij,bjd->bid", sim.softmax(-1), bf16r(v))
    ref = ref.view(b, heads, nq, d).permute(0, 2, 1, 3).reshape(b, nq, heads * d)
    err = rel_l2(o, ref)
    assert err < 8e-3, f"rel L2 {err}"  # P is rounded to bf16 before the PV product


@pytest.mark.parametrize("case", [(2, 12, 77, 64), (1, 4, 300, 64), (2, 8, 128, 40)])
def test_attention_causal(cuda_device, case):
    """Causal self-attention of the CLIP text encoder (query i sees keys 0..i), also across several K/V tiles."""
    from stablediffusioneo_b200 import ops
    b, heads, n, d = case
    dev = cuda_device
    q = gen((b * heads, n, d), 1, dev)
    k = gen((b * heads, n, d), 2, dev)
    v = gen((b * heads, n, d), 3, dev)
    ldv = (n + 7) // 8 * 8
    vt = torch.zeros((b * heads, d, ldv), dtype=torch.bfloat16, device=dev)
    vt[:, :, :n] = v.transpose(1, 2).to(torch.bfloat16)
    scale = d ** -0.5
    o = ops.attention(q.to(torch.bfloat16), k.to(torch.bfloat16), vt, b, heads, n, n, d, ldv, scale, causal=True)
    sim = torch.einsum("bid,bjd->bij", bf16r(q), bf16r(k)) * scale
    sim = sim.masked_fill(torch.ones(n, n, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    ref = torch.einsum("bij,bjd->bid", sim.softmax(-1), bf16r(v))
    ref = ref.view(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, heads * d)
    assert rel_l2(o, ref) < 8e-3


def test_embedding_and_quick_gelu(cuda_device):
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    tok = gen((1000, 768), 1, dev)
    pos = gen((77, 768), 2, dev)
    ids = torch.randint(0, 1000, (3, 77), generator=torch.Generator().manual_seed(3)).to(dev)
    y, y2 = ops.embedding_add(ids, tok, pos)
    ref = tok[ids] + pos[None]
    assert torch.equal(y, ref) and rel_l2(y2, ref) < TOL
    x = gen((77, 768), 4, dev)
    w = gen((3072, 768), 5, dev) / math.sqrt(768)
    bias = gen((3072,), 6, dev) * 0.1
    out = ops.linear(x.to(torch.bfloat16), ops.pack_conv_weight(w), bias=bias, act=ops.SDEO_ACT_QUICK_GELU)
    z = bf16r(x) @ bf16r(w).t() + bias
    assert rel_l2(out, z * torch.sigmoid(1.702 * z)) < TOL


def test_elementwise(cuda_device):
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    # layout round trip with channel padding
    x = gen((2, 4, 32, 48), 1, dev)
    y = ops.nchw_to_nhwc(x, 8)
    assert y.shape == (2, 32, 48, 8)
    assert torch.equal(y[..., :4].float(), bf16r(x).permute(0, 2, 3, 1))
    assert torch.count_nonzero(y[..., 4:]) == 0
    back = ops.nhwc_to_nchw(y, 4)
    assert torch.equal(back, bf16r(x))
    x3 = gen((2, 320, 7, 5), 2, dev)
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(x3)), bf16r(x3))
    # upsample
    u = ops.upsample_nearest2x(nhwc(x3))
    refu = F.interpolate(bf16r(x3), scale_factor=2, mode="nearest")
    assert torch.equal(u.permute(0, 3, 1, 2).float(), refu)
    # add_scaled
    a, b = gen((3, 5, 7, 24), 3, dev), gen((3, 5, 7, 24), 4, dev)
    s = ops.add_scaled(a.to(torch.bfloat16), b.to(torch.bfloat16), 0.5)
    assert rel_l2(s, bf16r(a) + 0.5 * bf16r(b)) < TOL
    # timestep embedding: [cos | sin] (util.py:165-169)
    t = torch.tensor([951, 1], dtype=torch.int64, device=dev)
    e = ops.timestep_embedding(t, 2, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=dev) / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert (e.float() - ref).abs().max().item() < 1e-2
    # silu / casts
    z = gen((1000,), 5, dev)
    assert rel_l2(ops.silu(z.to(torch.bfloat16)), F.silu(bf16r(z))) < TOL
    assert torch.equal(ops.to_f32(ops.to_bf16(z)), bf16r(z))
    # row softmax
    # (register-resident single-read kernel for 2 / 4 / 8 / 12 vectors per thread; 1538 and 13000 columns take the three-pass one)
    for rows, cols in ((37, 1536), (5, 4096), (3, 6000), (2, 9216), (4, 12288), (3, 1538), (2, 13000), (7, 4)):
        sm = gen((rows, cols), 6 + cols, dev) * 3
        sm[0, cols // 2] = 40.0   # a dominant score: the row maximum matters
        p = ops.softmax_rows(sm, 0.125)
        ref = (sm * 0.125).softmax(-1)
        assert rel_l2(p, ref) < TOL, (rows, cols)
        assert (p.float().sum(-1) - 1).abs().max().item() < 2e-2


def test_cfg_ddim_step(cuda_device):
    """e = eu + s(ec - eu); x0 = (x - sqrt(1-a_t) e)/sqrt(a_t); x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev-sig^2) e + sig*noise
    (cldm/ddim_hacked.py:192,215,226-230)."""
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    n, c, h, w = 2, 4, 32, 48
    ec, eu, x, noise = (gen((n, c, h, w), s, dev) for s in (1, 2, 3, 4))
    a_t, a_prev, sigma, s = 0.35, 0.52, 0.1, 9.0
    row = [s, math.sqrt(1 - a_t), 1 / math.sqrt(a_t), math.sqrt(a_prev), math.sqrt(1 - a_prev - sigma ** 2), sigma, 0, 0]
    table = torch.tensor([[0.0] * 8, row], dtype=torch.float32, device=dev)
    idx = torch.tensor([1], dtype=torch.int32, device=dev)
    e = eu + s * (ec - eu)
    x0 = (x - math.sqrt(1 - a_t) * e) / math.sqrt(a_t)
    ref = math.sqrt(a_prev) * x0 + math.sqrt(1 - a_prev - sigma ** 2) * e + sigma * noise
    pred = torch.empty_like(x)
    xn = torch.empty((2 * n, h, w, 8), dtype=torch.bfloat16, device=dev)
    xp, _ = ops.cfg_ddim_step(ec, eu, x, table, step_idx=idx, noise=noise, pred_x0=pred, x_next=xn, dup=2)
    assert rel_l2(xp, ref) < 1e-5
    assert rel_l2(pred, x0) < 1e-5
    assert torch.equal(xn[:n, ..., :4].float(), bf16r(xp).permute(0, 2, 3, 1))
    assert torch.equal(xn[n:], xn[:n])
    assert torch.count_nonzero(xn[..., 4:]) == 0
    # NHWC eps (the UNet out-conv layout), no noise
    ec_l = ec.permute(0, 2, 3, 1).contiguous()
    eu_l = eu.permute(0, 2, 3, 1).contiguous()
    xp2, _ = ops.cfg_ddim_step(ec_l, eu_l, x, table, step_idx=idx, eps_nhwc=True)
    assert rel_l2(xp2, ref - sigma * noise) < 1e-5
    # noise as a table indexed by the device step counter (eta > 0 inside the captured step graph)
    nt = torch.stack([torch.full_like(noise, float("nan")), noise])
    xp3, _ = ops.cfg_ddim_step(ec, eu, x, table, step_idx=idx, noise_table=nt)
    assert rel_l2(xp3, ref) < 1e-5
    ops.counter_add(idx, 1)
    assert idx.item() == 2


# ---------------------------------------------------------------------------------------------------------------
# The TensorRT plugin's exact contract (fp16 NHWC in/out, fp32 gamma/beta, bSwish) against the REFERENCE's own CUDA
# kernels, compiled stand-alone from /root/reference into oracle/_ref/libgroupnorm_ref.so (oracle/Makefile)
# ---------------------------------------------------------------------------------------------------------------
def _ref_groupnorm_lib():
    import ctypes
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libgroupnorm_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libgroupnorm_ref.so not built (needs the reference checkout at build time)")
    lib = ctypes.CDLL(path)
    lib.ref_groupnorm_workspace_bytes.restype = ctypes.c_size_t
    lib.ref_groupnorm_enqueue.restype = ctypes.c_int
    lib.ref_groupnorm_enqueue.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 2
    return lib


@pytest.mark.parametrize("n,c,h,w,swish", [(2, 320, 32, 48, True), (2, 640, 16, 24, False), (2, 1280, 8, 12, True),
                                           (1, 2560, 8, 12, True), (1, 1920, 16, 24, True), (2, 960, 32, 48, False),
                                           (1, 512, 64, 96, True), (1, 256, 128, 192, True), (1, 128, 256, 384, True)])
def test_groupnorm_f16_plugin_contract_vs_reference_kernels(cuda_device, n, c, h, w, swish):
    """sdeo_groupnorm_nhwc_f16 vs GroupNormPlugin::enqueue's kernels (groupNormKernel.cu:49-266) on the UNet / VAE shapes the
    plugin's cPerBlock table serves, and both vs torch in float64. The reference kernels ignore epsilon
    (groupNormKernel.cu:190-194); with unit-scale inputs that is far below fp16 resolution."""
    import torch.nn.functional as F
    from stablediffusioneo_b200 import ops
    ref_lib = _ref_groupnorm_lib()
    dev = cuda_device
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3).half()
    gamma, beta = torch.randn((c,), generator=g) * 0.5 + 1.0, torch.randn((c,), generator=g) * 0.2
    gold = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
    gold = (F.silu(gold) if swish else gold).permute(0, 2, 3, 1)
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    ours = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
    y_ref = torch.empty_like(xd)
    ws = torch.empty(ref_lib.ref_groupnorm_workspace_bytes(), dtype=torch.uint8, device=dev)
    rc = ref_lib.ref_groupnorm_enqueue(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), y_ref.data_ptr(), n, c, h, w, int(swish),
                                       ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert ours.dtype == torch.float16
    o, r, gold = ours.float().cpu(), y_ref.float().cpu(), gold.float()
    e_ours, e_ref, e_pair = rel_l2(o, gold), rel_l2(r, gold), rel_l2(o, r)
    print(f"GN fp16 {n}x{c}x{h}x{w}: ours vs f64 {e_ours:.2e}, reference kernels vs f64 {e_ref:.2e}, ours vs reference {e_pair:.2e}")
    assert e_ours < 6e-4 and e_pair < 1e-3


@pytest.mark.parametrize("n,c,h,w,swish", [(3, 64, 7, 5, True), (1, 64, 1, 1, False), (5, 1280, 33, 17, True),
                                           (4, 256, 128, 128, True), (2, 128, 250, 301, False), (40, 320, 16, 24, True)])
@pytest.mark.parametrize("env", [{}, {"SDEO_GN_F16_HINTS": "1"}, {"SDEO_GN_F16_LAG": "0"}, {"SDEO_GN_F16_SWISH": "1"},
                                 {"SDEO_GN_F16_ONE_LEVEL": "1"}, {"SDEO_GN_F16_BUFS": "2", "SDEO_GN_F16_TILE_KB": "8"},
                                 {"SDEO_GN_F16_GROUPS": "2"}, {"SDEO_GN_F16_GROUPS": "2", "SDEO_GN_F16_BUFS": "6", "SDEO_GN_F16_TILE_KB": "8"}])
def test_groupnorm_f16_streamed_schedule_cases(cuda_device, n, c, h, w, swish, env):
    """The streamed kernel (csrc/groupnorm_stream.cu) on ragged tiles (hw not a multiple of the tile), one-pixel samples, more
    tiles than SMs, more samples than SMs per wave, called twice on one workspace (the call clears its own flags), with the
    L2 hints forced on, with the smallest legal apply lag (every CTA folds the partial slots itself), with the fp32 Swish, with
    the two-level fold switched off, with two small tile buffers (many visits per CTA) and with two thread groups on alternate
    units (each with its own buffer set); against torch in float64 and against the two-launch variant."""
    import torch.nn.functional as F
    from stablediffusioneo_b200 import ops
    dev = cuda_device
    g = torch.Generator().manual_seed(n * 1000 + c + h)
    x = (torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3).half()
    x[:, :, :, : c // 32] *= 4.0  # one group per sample on a different scale: a mixed-up (mean, rstd) shows
    gamma, beta = torch.randn((c,), generator=g) * 0.5 + 1.0, torch.randn((c,), generator=g) * 0.2
    gold = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
    gold = (F.silu(gold) if swish else gold).permute(0, 2, 3, 1).float()
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    keys = ("SDEO_GN_F16_HINTS", "SDEO_GN_F16_LAG", "SDEO_GN_F16_SWISH", "SDEO_GN_F16_TWO_PASS", "SDEO_GN_F16_VARIANT",
            "SDEO_GN_F16_ONE_LEVEL", "SDEO_GN_F16_BUFS", "SDEO_GN_F16_TILE_KB", "SDEO_GN_F16_GROUPS")
    saved = {k: os.environ.pop(k, None) for k in keys}
    try:
        os.environ.update(env)
        os.environ["SDEO_GN_F16_VARIANT"] = "stream"  # small samples would otherwise take the resident kernel
        first = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        second = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        torch.cuda.synchronize()
        os.environ["SDEO_GN_F16_TWO_PASS"] = "1"
        two = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        torch.cuda.synchronize()
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    assert torch.equal(first, second), "not deterministic across calls on one workspace"
    o, t = first.float().cpu(), two.float().cpu()
    assert torch.isfinite(o).all()
    e_gold, e_two = rel_l2(o, gold), rel_l2(o, t)
    worst = ((o - gold).abs() / (gold.abs() + 1.0)).max().item()
    print(f"GN fp16 streamed {n}x{c}x{h}x{w} {env}: vs f64 {e_gold:.2e} (worst {worst:.2e}), vs two-launch {e_two:.2e}")
    assert e_gold < 6e-4 and e_two < 6e-4 and worst < 4e-3


@pytest.mark.parametrize("n,c,h,w,swish", [(2, 320, 32, 48, True), (2, 2560, 8, 12, True), (3, 64, 7, 5, True), (1, 64, 1, 1, False),
                                           (5, 1280, 33, 17, True), (2, 640, 32, 48, False), (40, 320, 16, 24, True)])
@pytest.mark.parametrize("cluster", [0, 1, 2, 4, 8, 16])
def test_groupnorm_f16_resident_variant(cuda_device, n, c, h, w, swish, cluster):
    """The resident kernel (sample kept in the shared memory of one thread-block cluster, statistics exchanged through
    distributed shared memory) with the automatic and with forced cluster sizes, ragged row splits and ranks without rows;
    against torch in float64 and bit-for-bit against itself."""
    import ctypes
    import torch.nn.functional as F
    from stablediffusioneo_b200 import _lib, ops
    dev = cuda_device
    info = (ctypes.c_int32 * 3)()
    saved_v = os.environ.get("SDEO_GN_F16_VARIANT")
    os.environ["SDEO_GN_F16_VARIANT"] = "resident"  # (the default dispatch gives these shapes to the slab kernel)
    try:
        fits = _lib.load().sdeo_groupnorm_f16_variant(n, h * w, c, 32, 148, max(cluster, 8), info) == 2
    finally:
        os.environ.pop("SDEO_GN_F16_VARIANT")
        if saved_v is not None:
            os.environ["SDEO_GN_F16_VARIANT"] = saved_v
    if not fits:
        pytest.skip("the sample does not fit a cluster of this size")
    if cluster and (-(-h * w // cluster)) * c * 2 > 180 * 1024:
        pytest.skip("forced cluster size too small for the sample")
    g = torch.Generator().manual_seed(n * 1000 + c + h)
    x = (torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3).half()
    x[:, :, :, : c // 32] *= 4.0
    gamma, beta = torch.randn((c,), generator=g) * 0.5 + 1.0, torch.randn((c,), generator=g) * 0.2
    gold = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
    gold = (F.silu(gold) if swish else gold).permute(0, 2, 3, 1).float()
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    keys = ("SDEO_GN_F16_VARIANT", "SDEO_GN_F16_CLUSTER", "SDEO_GN_F16_TWO_PASS")
    saved = {k: os.environ.pop(k, None) for k in keys}
    try:
        os.environ["SDEO_GN_F16_VARIANT"] = "resident"
        if cluster:
            os.environ["SDEO_GN_F16_CLUSTER"] = str(cluster)
        first = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        second = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        torch.cuda.synchronize()
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    assert torch.equal(first, second)
    o = first.float().cpu()
    assert torch.isfinite(o).all()
    e_gold = rel_l2(o, gold)
    worst = ((o - gold).abs() / (gold.abs() + 1.0)).max().item()
    print(f"GN fp16 resident {n}x{c}x{h}x{w} cluster {cluster or 'auto'}: vs f64 {e_gold:.2e} (worst {worst:.2e})")
    assert e_gold < 6e-4 and worst < 4e-3


@pytest.mark.parametrize("n,c,h,w,swish", [(2, 320, 32, 48, True), (2, 640, 16, 24, False), (2, 1280, 8, 12, True), (2, 2560, 8, 12, True),
                                           (3, 64, 7, 5, True), (1, 64, 1, 1, False), (5, 1280, 33, 17, True), (1, 32, 9, 11, True),
                                           (40, 320, 16, 24, True), (1, 512, 48, 96, True), (2, 128, 100, 77, False),
                                           (1, 4096, 5, 3, True)])
@pytest.mark.parametrize("split", [0, 2, 8])
def test_groupnorm_f16_slab_variant(cuda_device, n, c, h, w, swish, split):
    """The slab kernel (one CTA per sample and slab of whole groups, the slab parked in shared memory: no traffic between
    CTAs) forced on shapes up to the shared-memory limit -- 1 / 2 / 4 / 8 groups per slab, one-pixel samples, ragged row
    counts, 1 to 16 vectors per slab row, the automatic and forced splits of a slab over several CTAs (redundant
    statistics, disjoint rows normalised; pieces without rows); against torch in float64, the two-launch variant and
    bit-for-bit against itself."""
    import ctypes
    import torch.nn.functional as F
    from stablediffusioneo_b200 import _lib, ops
    dev = cuda_device
    g = torch.Generator().manual_seed(n * 1000 + c + h)
    x = (torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3).half()
    x[:, :, :, : c // 32] *= 4.0  # one group per sample on a different scale: a mixed-up (mean, rstd) shows
    gamma, beta = torch.randn((c,), generator=g) * 0.5 + 1.0, torch.randn((c,), generator=g) * 0.2
    gold = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
    gold = (F.silu(gold) if swish else gold).permute(0, 2, 3, 1).float()
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    keys = ("SDEO_GN_F16_VARIANT", "SDEO_GN_F16_SLAB_KB", "SDEO_GN_F16_TWO_PASS", "SDEO_GN_F16_SLAB_SPLIT")
    saved = {k: os.environ.pop(k, None) for k in keys}
    try:
        os.environ["SDEO_GN_F16_VARIANT"] = "slab"
        os.environ["SDEO_GN_F16_SLAB_KB"] = "200"
        if split:
            os.environ["SDEO_GN_F16_SLAB_SPLIT"] = str(split)
        info = (ctypes.c_int32 * 3)()
        assert _lib.load().sdeo_groupnorm_f16_variant(n, h * w, c, 32, 148, 8, info) == 3
        first = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        second = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        torch.cuda.synchronize()
        os.environ.pop("SDEO_GN_F16_VARIANT")
        os.environ["SDEO_GN_F16_TWO_PASS"] = "1"
        two = ops.groupnorm_f16(xd, gd, bd, eps=1e-5, silu=swish)
        torch.cuda.synchronize()
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    assert torch.equal(first, second)
    o, t = first.float().cpu(), two.float().cpu()
    assert torch.isfinite(o).all()
    e_gold, e_two = rel_l2(o, gold), rel_l2(o, t)
    worst = ((o - gold).abs() / (gold.abs() + 1.0)).max().item()
    print(f"GN fp16 slab {n}x{c}x{h}x{w} (groups per slab {info[0]}, vectors per row {info[1]}): vs f64 {e_gold:.2e} "
          f"(worst {worst:.2e}), vs two-launch {e_two:.2e}")
    assert e_gold < 6e-4 and e_two < 6e-4 and worst < 4e-3
