"""VAE Decoder and Encoder with the reference's names and state-dict keys (ldm/modules/diffusionmodules/model.py:41-203,
452-543, 546-652), on the same conv / GroupNorm kernels as the UNet."""
import numpy as np
import torch
import torch.nn as nn

from .... import ops
from .util import BF16, Conv2d, GroupNorm32, is_internal, nchw_view, nhwc, to_external, to_internal


# SDEO_NO_SUBPIXEL_UPSAMPLE=1: nearest-x2 + conv3x3 as a materialised upsample followed by the 3x3 conv
SUBPIXEL_UPSAMPLE = not __import__("os").environ.get("SDEO_NO_SUBPIXEL_UPSAMPLE")


def Normalize(in_channels, num_groups=32):
    return GroupNorm32(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    """nearest x2 (+ conv3x3) (model.py:50-65)."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if self.with_conv:
            self.conv = Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def run(self, x):
        if self.with_conv and SUBPIXEL_UPSAMPLE and x.shape[0] * x.shape[2] * x.shape[3] >= 1024:
            # conv3x3(nearest_x2(x)) = four 2x2 phase convolutions over the low-resolution x (ops.upsample2x_conv): 2.25x
            # fewer multiply-adds and no 4x intermediate tensor. (Small maps stay on the plain path: the four phase filters
            # hold 16/9 of the weight bytes, which is what a weight-streaming layer pays for.)
            from .util import _param_key
            key = _param_key(self.conv.weight)
            hit = self.__dict__.get("_phase_cache")
            if hit is None or hit[0] != key:
                hit = self.__dict__["_phase_cache"] = (key, ops.upsample2x_conv_weights(self.conv.weight))
            return nchw_view(ops.upsample2x_conv(nhwc(x), hit[1], bias=self.conv.bias_f32()))
        y = nchw_view(ops.upsample_nearest2x(nhwc(x)))
        return self.conv.run(y, gn_stats=True) if self.with_conv else y

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


class Downsample(nn.Module):
    """F.pad(x, (0,1,0,1)) + conv3x3 stride 2 without padding (model.py:66-87): the trailing zero row / column is the
    TMA out-of-bounds fill of the implicit-GEMM kernel (sdeo_conv_args::pad_hi), no padded copy."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if not with_conv:
            raise NotImplementedError("avg-pool Downsample is not on the SD1.5 VAE path (resamp_with_conv=True)")
        self.conv = Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def run(self, x):
        return self.conv.run(x, pad_hi=1, gn_stats=True)

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


class ResnetBlock(nn.Module):
    """GN+swish -> conv -> GN+swish -> conv -> + shortcut(x), temb unused by the Decoder (model.py:90-149)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if conv_shortcut or temb_channels > 0:
            raise NotImplementedError("conv_shortcut / temb are not on the VAE decoder path")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = torch.nn.Dropout(dropout)
        self.conv2 = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def run(self, x, temb=None):
        h = self.conv1.run(self.norm1.run(x, silu=True, defer=True), gn_stats=True)
        h = self.norm2.run(h, silu=True, defer=True)
        skip = self.nin_shortcut.run(x) if self.in_channels != self.out_channels else x
        return self.conv2.run(h, residual=skip, gn_stats=True)

    def forward(self, x, temb=None):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


class AttnBlock(nn.Module):
    """Single-head spatial self-attention with d = C (model.py:152-203). d = 512 exceeds the flash kernel's TMEM
    budget, and it runs once per image, so it is composed from the GEMM kernel: one fused q|k|v projection whose
    epilogue writes q, k [B, T, C] and v^T [B, C, T]; per sample S = q k^T (fp32), row softmax, O = P v."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self._cache = {}

    def _packed_qkv(self):
        from .util import _param_key
        params = (self.q.weight, self.k.weight, self.v.weight, self.q.bias, self.k.bias, self.v.bias)
        key = _param_key(*params)
        hit = self._cache.get("qkv")
        if hit is None or hit[0] != key:
            w = torch.cat([self.q.weight.detach(), self.k.weight.detach(), self.v.weight.detach()], 0)
            b = torch.cat([self.q.bias.detach(), self.k.bias.detach(), self.v.bias.detach()], 0).contiguous()
            self._cache["qkv"] = (key, ops.pack_conv_weight(w), b)
            hit = self._cache["qkv"]
        return hit[1], hit[2]

    def run(self, x):
        b, c, h, w = x.shape
        t = h * w
        if t % 64 != 0 or c % 64 != 0:
            raise NotImplementedError("VAE AttnBlock needs H*W and C to be multiples of 64")
        hn = self.norm.run(x, silu=False)
        pw, bias = self._packed_qkv()
        q = torch.empty((b, t, c), dtype=BF16, device=x.device)
        k = torch.empty_like(q)
        vt = torch.empty((b, c, t), dtype=BF16, device=x.device)
        ops.qkv_project(nhwc(hn).reshape(b, t, c), pw, 1, c, 0, q=q, k=k, vt=vt, ldv=t, bias=bias)
        o = torch.empty((b, t, c), dtype=BF16, device=x.device)
        scale = float(int(c) ** (-0.5))
        for i in range(b):
            k_as_filter = ops.PackedWeight(k[i], t, 1, c, 0)          # rows = keys, K = channels
            s = ops.linear(q[i], k_as_filter, out_fp32=True)          # [t, t] fp32 scores
            p = ops.softmax_rows(s, scale)                            # bf16
            v_as_filter = ops.PackedWeight(vt[i], c, 1, t, 0)         # rows = channels, K = keys
            ops.conv2d(p.reshape(1, 1, t, t), v_as_filter, out=o[i].reshape(1, 1, t, c))
        return self.proj_out.run(nchw_view(o.reshape(b, h, w, c)), residual=x, gn_stats=True)

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)))


def make_attn(in_channels, attn_type="vanilla", attn_kwargs=None):
    if attn_type != "vanilla":
        raise NotImplementedError(f"attn_type {attn_type} is not on the VAE decoder path")
    return AttnBlock(in_channels)


class Encoder(nn.Module):
    """model.py:452-543: conv_in -> per level ResnetBlocks (+ Downsample) -> mid (Res, Attn, Res) -> GN+swish -> conv_out
    (2 z_channels when double_z: the moments the 1x1 quant_conv turns into mean | logvar)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if use_linear_attn:
            raise NotImplementedError("linear attention is not on the SD1.5 VAE path")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.out_channels = 2 * z_channels if double_z else z_channels
        self.conv_out = Conv2d(block_in, self.out_channels, kernel_size=3, stride=1, padding=1)

    def run(self, x):
        """model.py:514-543 on internal tensors (the reference's `hs` list only ever reads its last element)."""
        h = self.conv_in.run(x, gn_stats=True)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block].run(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block].run(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample.run(h)
        h = self.mid.block_1.run(h)
        h = self.mid.attn_1.run(h)
        h = self.mid.block_2.run(h)
        return self.conv_out.run(self.norm_out.run(h, silu=True, defer=True))

    def forward(self, x):
        if is_internal(x):
            return self.run(x)
        return to_external(self.run(to_internal(x)), self.out_channels)


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or give_pre_end or tanh_out:
            raise NotImplementedError("linear attention / give_pre_end / tanh_out are not on the VAE decoder path")
        self.ch = ch
        self.out_ch = out_ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.give_pre_end = give_pre_end
        self.tanh_out = tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def run(self, z):
        """model.py:619-652 on internal tensors. Returns an internal tensor with out_ch (padded to 8) channels."""
        h = self.conv_in.run(z, gn_stats=True)
        h = self.mid.block_1.run(h)
        h = self.mid.attn_1.run(h)
        h = self.mid.block_2.run(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block].run(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block].run(h)
            if i_level != 0:
                h = self.up[i_level].upsample.run(h)
        return self.conv_out.run(self.norm_out.run(h, silu=True, defer=True))

    def forward(self, z):
        self.last_z_shape = z.shape
        if is_internal(z):
            return self.run(z)
        return to_external(self.run(to_internal(z)), self.out_ch)
