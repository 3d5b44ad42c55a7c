"""UNetModel and its blocks with the reference's names, constructor arguments and state-dict keys
(ldm/modules/diffusionmodules/openaimodel.py), running on libsdeo.so.

Fusions relative to the reference's op-by-op execution (results identical up to bf16 rounding):
  GroupNorm32 + SiLU                -> one normalisation pass (groupNormPlugin's bSwish contract)
  conv + bias + emb broadcast-add   -> conv epilogue (openaimodel.py:264-273)
  conv + skip_connection(x) + add   -> conv epilogue residual (openaimodel.py:275)
  torch.cat([h, skip], 1) + GN/conv -> dual-source kernels, no concatenated tensor (CatPair)
"""
from abc import abstractmethod

import torch
import torch.nn as nn

from .... import ops
from ..attention import SpatialTransformer
from . import util
from .util import (BF16, CatPair, Conv2d, SiLU, conv_nd, is_internal, linear, nchw_view, nhwc, normalization, operand,
                   timestep_embedding, to_external, to_internal, zero_module)


def exists(x):
    return x is not None


def emb_to_internal(emb):
    """time embedding [N, E]: keep bf16; convert fp32 callers."""
    if emb.dtype == BF16:
        return emb.contiguous()
    return ops.to_bf16(emb.float())


def silu_of(emb):
    """SiLU(emb) is shared by every ResBlock of a forward pass: compute once per emb tensor."""
    cached = getattr(emb, "_sdeo_silu", None)
    if cached is None:
        cached = ops.silu(emb)
        emb._sdeo_silu = cached
    return cached


class StepEmb:
    """Per-image table of every ResBlock's `emb_layers` output for all S DDIM timesteps. The time embedding depends
    only on t (openaimodel.py:769-770, 264-267), so the engine computes [S, Cout] tables once per image and the conv
    epilogue adds row `*step_ctr` (device-side step counter): 2 + 22 (UNet) / 2 + 10 (ControlNet) M=batch GEMV
    launches leave the per-step graph. Built by `StepEmb.build`; passed where `emb` goes."""

    def __init__(self, tables, step_ctr):
        self.tables, self.step_ctr = tables, step_ctr

    @staticmethod
    def build(net, t_emb_all, step_ctr, into=None):
        """net: UNetModel / ControlNet (has time_embed); t_emb_all: bf16 [S, model_channels]."""
        emb_all = net.time_embed[2].run(net.time_embed[0].run(t_emb_all, act=ops.SDEO_ACT_SILU))
        act = ops.silu(emb_all)
        tables = {} if into is None else into.tables
        for mod in net.modules():
            if isinstance(mod, ResBlock):
                t = mod.emb_layers[1].run(act, out_fp32=True)
                if id(mod) in tables:
                    tables[id(mod)].copy_(t)   # the captured step graph holds the table addresses
                else:
                    tables[id(mod)] = t
        return StepEmb(tables, step_ctr) if into is None else into


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        """Apply the module to `x` given `emb` timestep embeddings."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Passes timestep embeddings / context to the children that take them (openaimodel.py:73-87)."""

    def run(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer.run(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer.run(x, context)
            elif isinstance(layer, SiLU):
                raise RuntimeError("SiLU inside TimestepEmbedSequential is fused by the owning module")
            elif isinstance(layer, Conv2d):
                x = layer.run(x, stream=util.STREAM_FP32, gn_stats=True)
            else:
                x = layer.run(x)
        return x

    def forward(self, x, emb=None, context=None):
        if is_internal(x):
            return self.run(x, emb, context)
        ctx = context
        if ctx is not None and not isinstance(ctx, list) and ctx.dtype != BF16:
            ctx = ops.to_bf16(ctx.float())
        e = emb_to_internal(emb) if emb is not None else None
        y = self.run(to_internal(x), e, ctx)
        last = self[-1]
        c = last.out_channels if isinstance(last, Conv2d) else None
        return to_external(y, c)


class Upsample(nn.Module):
    """nearest x2 then conv3x3 (openaimodel.py:90-118)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if dims != 2:
            raise NotImplementedError("only dims=2 is on the ControlNet-SD1.5 path")
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)

    def run(self, x):
        assert x.shape[1] == self.channels
        y = nchw_view(ops.upsample_nearest2x(nhwc(operand(x))))
        return self.conv.run(y, stream=util.STREAM_FP32, gn_stats=True) if self.use_conv else y

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


class Downsample(nn.Module):
    """conv3x3 stride 2 (openaimodel.py:133-159)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if dims != 2 or not use_conv:
            raise NotImplementedError("only the strided-conv 2-D Downsample is on the ControlNet-SD1.5 path")
        self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=padding)

    def run(self, x):
        assert x.shape[1] == self.channels
        return self.op.run(x, stream=util.STREAM_FP32, gn_stats=True)

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


class ResBlock(TimestepBlock):
    """GN+SiLU -> conv3x3 (+emb) -> GN+SiLU -> conv3x3 -> + skip(x)   (openaimodel.py:162-275)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if up or down or use_scale_shift_norm or use_conv:
            raise NotImplementedError("resblock_updown / scale-shift-norm / conv skip are not on the ControlNet-SD1.5 path")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = nn.Sequential(normalization(channels), SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.updown = False
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def run(self, x, emb):
        """x: internal tensor or CatPair (decoder blocks); emb: bf16 [N, emb_channels]."""
        h = self.in_layers[0].run(x, silu=True, defer=True)   # folded into in_layers[2]'s operand path when it can be
        emb_step = None
        if isinstance(emb, StepEmb):
            emb_out, emb_step = emb.tables[id(self)], emb.step_ctr          # fp32 [S, Cout], row = DDIM step
        else:
            emb_out = self.emb_layers[1].run(silu_of(emb), out_fp32=True)   # fp32 [N, Cout]
        # GEMM results that feed a normalisation or a residual add stay fp32 (no extra bf16 rounding in the branch)
        # (with the GroupNorm folded into the consumer, h is only ever read as a bf16 operand: written once, in bf16, with
        #  its statistics taken from the fp32 accumulators)
        fold = util.FOLD_GN and util.FUSE_GN_STATS
        h = self.in_layers[2].run(h, emb=emb_out, emb_step=emb_step, out_fp32=util.STREAM_FP32 and not fold, gn_stats=True)
        h = self.out_layers[0].run(h, silu=True, defer=True)
        if isinstance(self.skip_connection, nn.Identity):
            assert not isinstance(x, CatPair)
            skip = x
        else:
            skip = self.skip_connection.run(x, out_fp32=util.STREAM_FP32)
        return self.out_layers[3].run(h, residual=skip, stream=util.STREAM_FP32, gn_stats=True)

    def forward(self, x, emb):
        if is_internal(x):
            return self.run(x, emb)
        return to_external(self.run(to_internal(x), emb_to_internal(emb)))


class UNetModel(nn.Module):
    """The SD UNet (openaimodel.py:412-788). Same constructor signature; options outside the ControlNet-SD1.5
    configuration space raise NotImplementedError instead of silently diverging."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 disable_self_attentions=None, num_attention_blocks=None, disable_middle_self_attn=False,
                 use_linear_in_transformer=False):
        super().__init__()
        _check_supported(dims, num_classes, use_scale_shift_norm, resblock_updown, use_spatial_transformer, context_dim,
                         n_embed, num_heads, num_head_channels, disable_self_attentions, num_attention_blocks,
                         use_linear_in_transformer, conv_resample)
        if type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = len(channel_mult) * [num_res_blocks] if isinstance(num_res_blocks, int) else list(num_res_blocks)
        if len(self.num_res_blocks) != len(channel_mult):
            raise ValueError("provide num_res_blocks either as an int (globally constant) or as a list/tuple (per-level) "
                             "with the same length as channel_mult")
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = False
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), SiLU(),
                                        linear(time_embed_dim, time_embed_dim))

        def res(cin, cout):
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        def st(ch):
            return SpatialTransformer(ch, num_heads, ch // num_heads, depth=transformer_depth, context_dim=context_dim,
                                      disable_self_attn=False, use_linear=False, use_checkpoint=use_checkpoint)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(self.num_res_blocks[level]):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), st(ch), res(ch, ch))
        self._skip_chans = list(skip_chans)
        self._build_decoder(skip_chans, ch, ds, channel_mult, model_channels, attention_resolutions, res, st,
                            conv_resample, dims, out_channels)

    def _build_decoder(self, skip_chans, ch, ds, channel_mult, model_channels, attention_resolutions, res, st,
                       conv_resample, dims, out_channels):
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(self.num_res_blocks[level] + 1):
                ich = skip_chans.pop()
                layers = [res(ch + ich, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(st(ch))
                if level and i == self.num_res_blocks[level]:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))

    # ---- shared pieces -------------------------------------------------------------------------------------
    def embed_time(self, timesteps):
        """timestep_embedding -> time_embed MLP (openaimodel.py:769-770); returns bf16 [N, 4*model_channels]."""
        t_emb = timestep_embedding(timesteps, self.model_channels, repeat_only=False)
        return self.time_embed[2].run(self.time_embed[0].run(t_emb, act=ops.SDEO_ACT_SILU))

    def run_out(self, h, out_fp32=True):
        """out: GroupNorm32 + SiLU + conv3x3 (openaimodel.py:728-732). fp32 NHWC-physical result by default."""
        return self.out[2].run(self.out[0].run(h, silu=True, defer=True), out_fp32=out_fp32)

    def run(self, x, emb, context):
        hs = []
        h = x
        for module in self.input_blocks:
            h = module.run(h, emb, context)
            hs.append(h)
        h = self.middle_block.run(h, emb, context)
        for module in self.output_blocks:
            h = module.run(CatPair(h, hs.pop()), emb, context)
        return self.run_out(h)

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """x fp32 [N, C, H, W], timesteps int64 [N], context [N, 77, ctx_dim] -> eps fp32 [N, C_out, H, W]."""
        assert y is None, "class-conditional UNet is not on the ControlNet-SD1.5 path"
        ctx = context if context.dtype == BF16 else ops.to_bf16(context.float())
        emb = self.embed_time(timesteps)
        return to_external(self.run(to_internal(x), emb, ctx), self.out_channels)


def _check_supported(dims, num_classes, use_scale_shift_norm, resblock_updown, use_spatial_transformer, context_dim,
                     n_embed, num_heads, num_head_channels, disable_self_attentions, num_attention_blocks,
                     use_linear_in_transformer, conv_resample):
    bad = []
    if dims != 2: bad.append("dims != 2")
    if num_classes is not None: bad.append("num_classes")
    if use_scale_shift_norm: bad.append("use_scale_shift_norm")
    if resblock_updown: bad.append("resblock_updown")
    if not use_spatial_transformer or context_dim is None: bad.append("use_spatial_transformer=False / context_dim=None")
    if n_embed is not None: bad.append("n_embed")
    if num_heads == -1 or num_head_channels != -1: bad.append("num_head_channels (use num_heads)")
    if disable_self_attentions is not None or num_attention_blocks is not None: bad.append("per-level attention switches")
    if use_linear_in_transformer: bad.append("use_linear_in_transformer")
    if not conv_resample: bad.append("conv_resample=False")
    if bad:
        raise NotImplementedError("UNet/ControlNet options outside the ControlNet-SD1.5 path: " + ", ".join(bad))
