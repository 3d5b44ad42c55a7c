"""Leaf layers and schedule helpers with the reference's names (ldm/modules/diffusionmodules/util.py), executing
on the sm_100a kernels in libsdeo.so. Parameters keep PyTorch's names/shapes/dtypes so state dicts interchange
with the reference modules; the packed bf16 filters the kernels stream are derived caches.

Activation convention ("internal" tensors): bf16, logical shape [N, C, H, W], channels_last memory (physically
NHWC). Any public module also accepts "external" fp32 NCHW tensors and then returns fp32 NCHW (see `boundary`)."""
import functools
import os
import math

import numpy as np
import torch
import torch.nn as nn

from .... import ops
from ...._lib import SDEO_ACT_NONE, SDEO_ACT_SILU

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------------------
# tensor format helpers
# ------------------------------------------------------------------------------------------------------------
class CatPair:
    """A deferred torch.cat([a, b], dim=1) of two internal tensors (cldm/cldm.py:39,41): consumers fuse the concat
    (dual-source GroupNorm, dual-source K loop in the conv)."""

    def __init__(self, a, b):
        assert a.shape[0] == b.shape[0] and a.shape[2:] == b.shape[2:]
        self.a, self.b = a, b

    @property
    def shape(self):
        return (self.a.shape[0], self.a.shape[1] + self.b.shape[1]) + tuple(self.a.shape[2:])


# Residual-stream tensors (block outputs that later blocks add onto) are kept in fp32 — with ~90 sequential
# residual adds per forward, re-rounding the stream to bf16 after each one is the dominant error term (measured:
# eps rel L2 1.3e-2 with a bf16 stream vs the 1e-2 gate). A stream tensor is an fp32 NHWC-physical tensor tagged
# `_sdeo_stream`, optionally carrying `_twin`: the bf16 copy the producing epilogue wrote for TMA consumers.
STREAM_FP32 = True
# Convolutions whose fp32 output feeds a GroupNorm also emit per-channel partial statistics from their epilogue
# (sdeo_conv_args::gn_stats); the GroupNorm then reads the tensor once. SDEO_NO_GN_STATS=1 restores the standalone pass.
FUSE_GN_STATS = not os.environ.get("SDEO_NO_GN_STATS")
# LayerNorms of the transformer blocks are folded into the GEMM that consumes them (no LayerNorm kernel, no normalised
# tensor in memory): the producer of the LayerNorm input leaves per-row (sum, sum of squares) partials, the consumer runs
# on the raw bf16 input with gamma folded into its weight and corrects acc -> rstd * (acc - mean * colsum) + bias' in its
# epilogue. SDEO_NO_LN_FOLD=1 restores the standalone LayerNorm pass.
FOLD_LN = not os.environ.get("SDEO_NO_LN_FOLD")
# GroupNorms whose input carries its producers' partial statistics and whose consumer is a convolution CAN be folded into
# that convolution's operand path (sdeo_conv_args::gnf_*): no GroupNorm launch, no normalised tensor in memory; the conv
# reads the RAW bf16 tensor (an fp32 stream's bf16 twin). Opt-in (SDEO_GN_FOLD=1): measured on B200 it removes 81 of the
# step's 428 launches but the step gets SLOWER (3.79 -> 4.65 ms at 256x384): normalising a 16 KB operand tile costs ~100
# instructions per 8 channels on the seven otherwise idle warps (~2500 cycles per tile against ~320 for the tile's MMAs; the
# HALO tiling amortises that over nine taps, the tap-by-tap tiling of the small feature maps cannot), every CTA re-folds
# the producers' partial statistics (4 us in front of the first MMA), CTA pairs are lost, and reading the bf16 twin
# instead of the fp32 stream adds one rounding per GroupNorm (tiny-model eps 1.13e-2 against the 1e-2 gate). DESIGN 5a.
FOLD_GN = bool(os.environ.get("SDEO_GN_FOLD"))
# bf16 outputs (the VAE's activations) CAN also leave epilogue statistics for their GroupNorm (SDEO_BF16_GN_STATS=1). Opt-in:
# measured on the 512x512 batch-16 decode the statistics epilogue (three block-wide barriers and a shared-memory reduction
# per tile, on grids of thousands of tiles) costs the convs more (~5 ms) than the standalone statistics pass it replaces
# (4.6 ms), and the apply pass then folds the partials first (+1.5 ms).
BF16_GN_STATS = bool(os.environ.get("SDEO_BF16_GN_STATS"))
# ... but SMALL bf16 outputs (the batch-1 decode: tens of tiles per conv, every launch latency-bound) do leave them: the
# statistics launch disappears and the apply reads the tensor once. Limit in output elements (SDEO_BF16_GN_STATS_MAX).
BF16_GN_STATS_MAX = int(os.environ.get("SDEO_BF16_GN_STATS_MAX", "0"))


def _out_elems(x, cout, stride):
    """Output elements of a same-padded conv over x (internal tensor / CatPair), or None when the shape is not at hand."""
    t = x.a if isinstance(x, CatPair) else x
    shp = getattr(t, "shape", None)
    if shp is None or len(shp) != 4:
        return None
    n, _, h, w = shp
    return n * ((h + stride - 1) // stride) * ((w + stride - 1) // stride) * cout


class DeferredGN:
    """GroupNorm(x) (+ SiLU) that has not been computed: the raw input (internal tensor or CatPair), the module, the
    activation flag. Conv2d.run folds it into its operand path; `.value()` materialises it for anything else."""

    def __init__(self, x, gn, silu):
        self.x, self.gn, self.silu = x, gn, silu

    @property
    def shape(self):
        return self.x.shape

    def value(self):
        return self.gn.run(self.x, silu=self.silu)

    def fold(self):
        """(x1, x2 or None, ops.GnFold) with the raw bf16 operands in NHWC."""
        gn = self.gn
        if isinstance(self.x, CatPair):
            a, b = self.x.a, self.x.b
            n = a.shape[0]
            sa = ops.gn_stats_fold(a._gn_stats, n, a.shape[1])
            sb = ops.gn_stats_fold(b._gn_stats, n, b.shape[1])
            return nhwc(operand(a)), nhwc(operand(b)), ops.GnFold(sa, sb, gn.weight.detach(), gn.bias.detach(),
                                                                   gn.num_groups, gn.eps, self.silu)
        x = self.x
        st = ops.gn_stats_fold(x._gn_stats, x.shape[0], x.shape[1])
        return nhwc(operand(x)), None, ops.GnFold(st, None, gn.weight.detach(), gn.bias.detach(), gn.num_groups, gn.eps,
                                                   self.silu)


class DeferredLN:
    """LayerNorm(x) that has not been computed: the raw bf16 x (the fp32 stream's twin), the row statistics its
    producer left, and the LayerNorm module. Consumers (CrossAttention / GEGLU projections) fold it into their GEMM;
    `.value()` materialises it with the standalone kernel for anything else."""

    def __init__(self, x, ln):
        self.x, self.ln = x, ln

    @property
    def shape(self):
        return self.x.shape

    @property
    def device(self):
        return self.x.device

    @property
    def dtype(self):
        return BF16

    def raw(self):
        return operand(self.x)

    def value(self):
        return ops.layernorm(self.x, self.ln.weight.detach(), self.ln.bias.detach(), self.ln.eps)

    def fold(self, csum):
        buf, parts, rows = self.x._row_stats
        return ops.LnFold(buf, parts, rows, self.x.shape[-1], self.ln.eps, csum)


def fold_ln_weight(weight, bias, ln, geglu=False):
    """(packed weight of W * gamma, bias' = bias + W @ beta in the kernel's row order, column sums of the PACKED bf16
    weight) for a Linear [out, in] whose input is LayerNorm `ln`."""
    w = weight.detach().float()
    g, b = ln.weight.detach().float(), ln.bias.detach().float()
    pw = ops.pack_conv_weight(w * g[None, :], geglu=geglu)
    bp = w @ b
    if bias is not None:
        bp = bp + bias.detach().float()
    bp = ops.pack_geglu_bias(bp.contiguous(), pw.geglu_bn) if geglu else bp.contiguous()
    rows = pw.data.shape[0]
    csum = pw.data.float().sum(dim=1).contiguous()
    if bp.numel() < rows:  # rows are padded to a multiple of 16
        bp = torch.cat([bp, torch.zeros(rows - bp.numel(), dtype=bp.dtype, device=bp.device)])
    return pw, bp, csum


def make_stream(y_f32, y_bf16=None):
    """[N,H,W,C] fp32 (+ optional bf16 twin) from a conv epilogue -> tagged internal stream tensor."""
    t = nchw_view(y_f32) if y_f32.dim() == 4 else y_f32
    t._sdeo_stream = True
    t._gn_stats = getattr(y_f32, "_gn_stats", None)
    t._row_stats = getattr(y_f32, "_row_stats", None)
    t._twin = None if y_bf16 is None else (nchw_view(y_bf16) if y_bf16.dim() == 4 else y_bf16)
    return t


def is_stream(x):
    return getattr(x, "_sdeo_stream", False)


def operand(x):
    """bf16 tensor to feed a GEMM through TMA: x itself, its bf16 twin, or (fallback) a cast pass."""
    if x.dtype == BF16:
        return x
    twin = getattr(x, "_twin", None)
    if twin is not None:
        return twin
    if x.dim() == 4:
        return nchw_view(ops.to_bf16(nhwc(x)))
    return ops.to_bf16(x)


def is_internal(x):
    if isinstance(x, CatPair) or is_stream(x):
        return True
    return x.dtype == BF16 and x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()


def nhwc(x):
    """internal [N,C,H,W] channels_last -> zero-copy [N,H,W,C] contiguous view."""
    return x.permute(0, 2, 3, 1)


def nchw_view(y):
    """[N,H,W,C] contiguous -> internal [N,C,H,W] channels_last view."""
    return y.permute(0, 3, 1, 2)


def to_internal(x, pad_to=8, scale=1.0):
    """fp32 NCHW of any layout -> internal tensor = scale * x, channels zero-padded to a multiple of `pad_to`
    (TMA needs a 16-byte pixel stride; convs ignore the padded tail)."""
    if is_internal(x):
        assert scale == 1.0
        return x
    if x.dtype != torch.float32:
        x = x.float()
    c = x.shape[1]
    ld = (c + pad_to - 1) // pad_to * pad_to
    y = ops.nchw_to_nhwc(x.contiguous(), ld, scale)
    return nchw_view(y)


def to_external(x, c=None):
    """internal (bf16) or fp32 NHWC-physical tensor -> fp32 NCHW contiguous (first c channels)."""
    return ops.nhwc_to_nchw(nhwc(x), c)


def boundary(fn):
    """forward(self, x, ...) wrapper: external fp32 NCHW in -> fp32 NCHW out; internal in -> internal out."""

    @functools.wraps(fn)
    def wrapped(self, x, *args, **kwargs):
        if is_internal(x):
            return fn(self, x, *args, **kwargs)
        y = fn(self, to_internal(x), *args, **kwargs)
        if isinstance(y, (list, tuple)):
            return type(y)(to_external(t) for t in y)
        return to_external(y)

    return wrapped


def tokens_boundary(fn):
    """Same for token tensors [B, T, C]: fp32 in -> fp32 out, bf16 in -> bf16 out."""

    @functools.wraps(fn)
    def wrapped(self, x, *args, **kwargs):
        if x.dtype == BF16:
            return fn(self, x.contiguous(), *args, **kwargs)
        y = fn(self, ops.to_bf16(x.float()), *args, **kwargs)
        return ops.to_f32(y)

    return wrapped


def _param_key(*params):
    return tuple((p.data_ptr(), p._version, p.device) for p in params if p is not None)


# ------------------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------------------
class Conv2d(nn.Conv2d):
    """nn.Conv2d (3x3 pad 1 / 1x1 pad 0, stride 1 or 2) on the tcgen05 implicit-GEMM kernel."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        k = self.kernel_size[0]
        # padding 0 with a 3x3 stride-2 filter is the VAE encoder's Downsample, which pads (0,1,0,1) itself (model.py:78-84):
        # only runnable through run(..., pad_hi=1)
        self.tail_pad = k == 3 and self.stride == (2, 2) and self.padding == (0, 0)
        if self.kernel_size != (k, k) or k not in (1, 3) or self.stride[0] not in (1, 2) or self.stride[0] != self.stride[1] \
                or (self.padding != ((k // 2), (k // 2)) and not self.tail_pad) or self.groups != 1 or self.dilation != (1, 1):
            raise NotImplementedError(f"Conv2d configuration not on the ControlNet-SD1.5 path: {self}")
        self._cache = {}

    def packed(self, split=None):
        """PackedWeight for input channels split (c1, c2) (fused concat) or a single source. A ragged cin (4, 3)
        reads an input whose pixel stride is padded to 8 channels; TMA zero-fills beyond cin."""
        key = (_param_key(self.weight), split)
        hit = self._cache.get("w")
        if hit is None or hit[0] != key:
            w = self.weight.detach()
            if split is None:
                pw = ops.pack_conv_weight(w)
            else:
                pw = ops.pack_conv_weight(w, c1=split[0], c2=split[1])
            self._cache["w"] = (key, pw)
            hit = self._cache["w"]
        return hit[1]

    def bias_f32(self):
        return self.bias.detach() if self.bias is not None else None

    def run(self, x, emb=None, residual=None, scale=1.0, act=SDEO_ACT_NONE, out_fp32=False, stream=False, emb_step=None,
            gn_stats=False, row_stats=False, pad_hi=0, out_f32=None, out_twin=None):
        """x: internal tensor (bf16 or fp32 stream) or CatPair. Returns an internal bf16 tensor; fp32 NHWC-physical when
        out_fp32; an fp32 stream tensor with a bf16 twin when `stream`. A bf16 output whose channel count is not a
        multiple of 8 is zero-padded to one (so a following conv can TMA it)."""
        if self.tail_pad != bool(pad_hi):
            raise NotImplementedError("a 3x3 stride-2 padding-0 Conv2d runs only as F.pad(x, (0,1,0,1)) + conv (pad_hi=1)")
        res = nhwc(residual) if residual is not None else None
        out = None
        cout = self.out_channels
        gnf = None
        if isinstance(x, DeferredGN):
            x1, x2, gnf = x.fold()
        if not (out_fp32 or stream) and cout % 8 != 0:
            assert residual is None
            n, _, h, w = x.shape
            k, s = self.kernel_size[0], self.stride[0]
            pd = 0 if pad_hi else k // 2
            ho, wo = (h + 2 * pd + pad_hi - k) // s + 1, (w + 2 * pd + pad_hi - k) // s + 1
            out = torch.empty((n, ho, wo, (cout + 7) // 8 * 8), dtype=BF16, device=self.weight.device)
            ops.memset(out, 0)
        if out_f32 is not None:   # stream output written into caller-owned NHWC buffers (fp32 + bf16 twin)
            assert stream and out_twin is not None
            out = out_f32
        small_out = False
        if gn_stats and BF16_GN_STATS_MAX > 0 and not (out_fp32 or stream):
            ne = _out_elems(x, cout, self.stride[0])
            small_out = ne is not None and ne <= BF16_GN_STATS_MAX
        kw = dict(bias=self.bias_f32(), emb=emb, residual=res, scale=scale, act=act, stride=self.stride[0],
                  out_fp32=out_fp32 or stream, out=out, out2=out_twin, twin=stream, emb_step=emb_step, pad_hi=pad_hi,
                  gn_stats=gn_stats and (out_fp32 or stream or FOLD_GN or BF16_GN_STATS or small_out) and FUSE_GN_STATS,
                  row_stats=row_stats and (out_fp32 or stream) and FOLD_LN)
        if gnf is not None:
            y = ops.conv2d(x1, self.packed((x1.shape[3], x2.shape[3])) if x2 is not None else self.packed(), x2=x2, gnf=gnf, **kw)
        elif isinstance(x, CatPair):
            a, b = operand(x.a), operand(x.b)
            y = ops.conv2d(nhwc(a), self.packed((a.shape[1], b.shape[1])), x2=nhwc(b), **kw)
        else:
            y = ops.conv2d(nhwc(operand(x)), self.packed(), **kw)
        if stream:
            return make_stream(y[0], y[1])
        t = nchw_view(y)
        t._gn_stats = getattr(y, "_gn_stats", None)
        return t

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)), self.out_channels)


class Linear(nn.Linear):
    """nn.Linear on the same GEMM kernel (1x1 case). Token tensors [..., K] bf16."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._cache = {}

    def packed(self):
        key = _param_key(self.weight)
        hit = self._cache.get("w")
        if hit is None or hit[0] != key:
            self._cache["w"] = (key, ops.pack_conv_weight(self.weight.detach()))
            hit = self._cache["w"]
        return hit[1]

    def run(self, x, residual=None, act=SDEO_ACT_NONE, out_fp32=False, stream=False, twin=False, row_stats=False):
        """x: bf16 tokens. stream: fp32 output tagged as residual stream (bf16 twin attached when `twin`); row_stats: the
        output feeds a LayerNorm that its consumer folds (per-row statistics from this epilogue)."""
        b = self.bias.detach() if self.bias is not None else None
        y = ops.linear(operand(x), self.packed(), bias=b, residual=residual, act=act, out_fp32=out_fp32 or stream,
                       twin=stream and twin, row_stats=row_stats and (out_fp32 or stream) and FOLD_LN)
        if stream:
            return make_stream(*y) if twin else make_stream(y)
        return y

    def forward(self, x):
        if x.dtype == BF16:
            return self.run(x.contiguous())
        return ops.to_f32(self.run(ops.to_bf16(x.float())))


class GroupNorm32(nn.GroupNorm):
    """GroupNorm with fp32 statistics (util.py:217-219); `run(..., silu=True)` fuses the following nn.SiLU
    (the groupNormPlugin bSwish contract, groupNormPlugin.cpp:291-304). Accepts a CatPair (concat seam may fall
    inside a group)."""

    def run(self, x, silu=False, defer=False):
        """defer=True (the caller hands the result straight to a Conv2d.run): a DeferredGN when the input carries its
        producers' statistics -- the convolution then applies the normalisation inside its operand path."""
        if defer and FOLD_GN and FUSE_GN_STATS and self.affine:
            parts = (x.a, x.b) if isinstance(x, CatPair) else (x,)
            # (K chunks of 64 channels per source: the first of two sources must fill its chunks)
            ok = all(getattr(p_, "_gn_stats", None) is not None and
                     (p_.dtype == BF16 or getattr(p_, "_twin", None) is not None) for p_ in parts)
            if ok and (len(parts) == 1 or parts[0].shape[1] % 64 == 0):
                return DeferredGN(x, self, silu)
        if isinstance(x, CatPair):
            a, b = x.a, x.b
            sa, sb = getattr(a, "_gn_stats", None), getattr(b, "_gn_stats", None)
            if a.dtype != b.dtype:  # mixed stream / bf16 halves: read both as bf16
                a, b = operand(a), operand(b)
                sa = sb = None
            y = ops.groupnorm(nhwc(a), self.weight.detach(), self.bias.detach(), self.eps, silu, x2=nhwc(b),
                              groups=self.num_groups, stats=sa if sb is not None else None, stats2=sb)
        else:
            y = ops.groupnorm(nhwc(x), self.weight.detach(), self.bias.detach(), self.eps, silu, groups=self.num_groups,
                              stats=getattr(x, "_gn_stats", None))
        return nchw_view(y)

    @boundary
    def forward(self, x):
        return self.run(x, silu=False)


class LayerNorm(nn.LayerNorm):
    def run(self, x):
        """bf16 LayerNorm(x) -- or, when x carries its producer's row statistics and a bf16 twin, a DeferredLN that the
        consuming projection folds into its GEMM."""
        if FOLD_LN and getattr(x, "_row_stats", None) is not None and getattr(x, "_twin", None) is not None:
            return DeferredLN(x, self)
        return ops.layernorm(x, self.weight.detach(), self.bias.detach(), self.eps)

    def forward(self, x):
        if x.dtype == BF16:
            return self.run(x.contiguous())
        return ops.to_f32(self.run(ops.to_bf16(x.float())))


class SiLU(nn.Module):
    """x * sigmoid(x). Inside ResBlock / hint block it is fused into the neighbouring kernel; standalone it is one pass."""

    def forward(self, x):
        if x.dtype == BF16:
            return ops.silu(x.contiguous())
        return ops.to_f32(ops.silu(ops.to_bf16(x.float())))


def conv_nd(dims, *args, **kwargs):
    if dims != 2:
        raise NotImplementedError("only dims=2 is on the ControlNet-SD1.5 path")
    return Conv2d(*args, **kwargs)


def linear(*args, **kwargs):
    return Linear(*args, **kwargs)


def normalization(channels):
    return GroupNorm32(32, channels)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """[cos | sin] sinusoidal embedding (util.py:154-174) -> bf16 [N, dim] on the device."""
    if repeat_only:
        raise NotImplementedError("repeat_only timestep embedding is not on the ControlNet-SD1.5 path")
    t = timesteps.to(torch.int64).contiguous()
    return ops.timestep_embedding(t, t.shape[0], dim, max_period=float(max_period))


# ------------------------------------------------------------------------------------------------------------
# schedules (host side, numpy; util.py:21-74)
# ------------------------------------------------------------------------------------------------------------
def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule != "linear":
        raise NotImplementedError(f"schedule '{schedule}' is not on the ControlNet-SD1.5 path")
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64, device="cpu") ** 2
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule for ddim "
              f"sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def extract_into_tensor(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))
