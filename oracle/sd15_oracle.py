"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

A CPU fp32 restatement (plain torch.nn.functional calls on a flat state dict) of the reference's ControlNet-SD1.5
denoising path: ControlNet / ControlledUnetModel / ControlLDM.apply_model, DDIMSampler, and the VAE Decoder.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module,
and only as the checker. Every function cites the reference file:line it restates (paths relative to the
reference checkout, /root/reference).

Pinned: tests/golden/*.pt were produced by tests/golden/make_golden.py, which imports the REAL reference modules
(ldm.modules.diffusionmodules.openaimodel, ldm.modules.attention, cldm.cldm, cldm.ddim_hacked,
ldm.modules.diffusionmodules.model) with identical weights and inputs; tests/test_oracle.py checks this
restatement against those fixtures. The reference itself ships no golden tensors for this path.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ------------------------------------------------------------------------------------------------------------
# configuration (models/cldm_v15.yaml is absent from the reference; values recovered in SURVEY.md §8 a-0)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    hint_channels: int = 3
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_heads: int = 8
    context_dim: int = 768
    transformer_depth: int = 1


@dataclass
class VAEConfig:
    ch: int = 128
    out_ch: int = 3
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    scale_factor: float = 0.18215


SD15 = UNetConfig()
SD15_VAE = VAEConfig()
# a small config with the same topology, for fast CPU checks against the real reference modules
TINY = UNetConfig(model_channels=64, context_dim=96)
TINY_VAE = VAEConfig(ch=32)


@dataclass
class Block:
    """One TimestepEmbedSequential entry: a list of (kind, params) layers."""
    layers: List[Tuple[str, dict]] = field(default_factory=list)


def unet_layout(cfg: UNetConfig):
    """Restates the constructor loops of UNetModel.__init__ (openaimodel.py:544-726): which layers make up each
    input / middle / output block and with which channel counts."""
    mc = cfg.model_channels
    inp: List[Block] = [Block([("conv", dict(cin=cfg.in_channels, cout=mc, k=3, stride=1))])]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", dict(cin=ch, cout=mult * mc))]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(("st", dict(c=ch, heads=cfg.num_heads)))
            inp.append(Block(layers))
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inp.append(Block([("down", dict(c=ch))]))
            chans.append(ch)
            ds *= 2
    mid = Block([("res", dict(cin=ch, cout=ch)), ("st", dict(c=ch, heads=cfg.num_heads)), ("res", dict(cin=ch, cout=ch))])
    out: List[Block] = []
    skip_chans = list(chans)
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = skip_chans.pop()
            layers = [("res", dict(cin=ch + ich, cout=mc * mult, split=(ch, ich)))]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(("st", dict(c=ch, heads=cfg.num_heads)))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", dict(c=ch)))
                ds //= 2
            out.append(Block(layers))
    return inp, mid, out, chans


# ------------------------------------------------------------------------------------------------------------
# parameter specs (state-dict names and shapes of the reference modules) and deterministic weights
# ------------------------------------------------------------------------------------------------------------
def _spec_res(p, cin, cout, emb_ch):
    s = [(p + "in_layers.0.weight", (cin,), "norm_w"), (p + "in_layers.0.bias", (cin,), "norm_b"),
         (p + "in_layers.2.weight", (cout, cin, 3, 3), "w"), (p + "in_layers.2.bias", (cout,), "b"),
         (p + "emb_layers.1.weight", (cout, emb_ch), "w"), (p + "emb_layers.1.bias", (cout,), "b"),
         (p + "out_layers.0.weight", (cout,), "norm_w"), (p + "out_layers.0.bias", (cout,), "norm_b"),
         (p + "out_layers.3.weight", (cout, cout, 3, 3), "w"), (p + "out_layers.3.bias", (cout,), "b")]
    if cin != cout:
        s += [(p + "skip_connection.weight", (cout, cin, 1, 1), "w"), (p + "skip_connection.bias", (cout,), "b")]
    return s


def _spec_attn(p, c, ctx):
    return [(p + "to_q.weight", (c, c), "w"), (p + "to_k.weight", (c, ctx), "w"), (p + "to_v.weight", (c, ctx), "w"),
            (p + "to_out.0.weight", (c, c), "w"), (p + "to_out.0.bias", (c,), "b")]


def _spec_st(p, c, ctx_dim):
    s = [(p + "norm.weight", (c,), "norm_w"), (p + "norm.bias", (c,), "norm_b"),
         (p + "proj_in.weight", (c, c, 1, 1), "w"), (p + "proj_in.bias", (c,), "b")]
    t = p + "transformer_blocks.0."
    s += _spec_attn(t + "attn1.", c, c)
    s += [(t + "ff.net.0.proj.weight", (8 * c, c), "w"), (t + "ff.net.0.proj.bias", (8 * c,), "b"),
          (t + "ff.net.2.weight", (c, 4 * c), "w"), (t + "ff.net.2.bias", (c,), "b")]
    s += _spec_attn(t + "attn2.", c, ctx_dim)
    for i in (1, 2, 3):
        s += [(t + f"norm{i}.weight", (c,), "norm_w"), (t + f"norm{i}.bias", (c,), "norm_b")]
    s += [(p + "proj_out.weight", (c, c, 1, 1), "w"), (p + "proj_out.bias", (c,), "b")]
    return s


def _spec_block(p, block: Block, emb_ch, ctx_dim):
    s = []
    for j, (kind, a) in enumerate(block.layers):
        q = f"{p}{j}."
        if kind == "conv":
            s += [(q + "weight", (a["cout"], a["cin"], 3, 3), "w"), (q + "bias", (a["cout"],), "b")]
        elif kind == "res":
            s += _spec_res(q, a["cin"], a["cout"], emb_ch)
        elif kind == "st":
            s += _spec_st(q, a["c"], ctx_dim)
        elif kind == "down":
            s += [(q + "op.weight", (a["c"], a["c"], 3, 3), "w"), (q + "op.bias", (a["c"],), "b")]
        elif kind == "up":
            s += [(q + "conv.weight", (a["c"], a["c"], 3, 3), "w"), (q + "conv.bias", (a["c"],), "b")]
    return s


def _is_zero_init(name: str) -> bool:
    """Tensors the reference zero-initialises (openaimodel.py:228,731; attention.py:422; cldm/cldm.py:162,282)."""
    return (name.endswith("out_layers.3.weight") or name.endswith("proj_out.weight") or name.startswith("zero_convs.")
            or name.startswith("middle_block_out.") or name == "out.2.weight" or name == "input_hint_block.14.weight")


def _mark_regular(spec):
    """'w' -> 'wk' (PyTorch-default-init variance 1/(3 fan_in)) for every weight the reference does NOT zero-initialise;
    the zero-initialised ones keep 'w' = N(0, 1/fan_in). This is SURVEY.md §8d's synthetic-weight scheme."""
    return [(n, sh, "wk" if (k == "w" and not _is_zero_init(n)) else k) for n, sh, k in spec]


def unet_param_spec(cfg: UNetConfig):
    return _mark_regular(_unet_param_spec(cfg))


def controlnet_param_spec(cfg: UNetConfig):
    return _mark_regular(_controlnet_param_spec(cfg))


def _unet_param_spec(cfg: UNetConfig):
    inp, mid, out, _ = unet_layout(cfg)
    mc, te = cfg.model_channels, cfg.model_channels * 4
    s = [("time_embed.0.weight", (te, mc), "w"), ("time_embed.0.bias", (te,), "b"),
         ("time_embed.2.weight", (te, te), "w"), ("time_embed.2.bias", (te,), "b")]
    for i, b in enumerate(inp):
        s += _spec_block(f"input_blocks.{i}.", b, te, cfg.context_dim)
    s += _spec_block("middle_block.", mid, te, cfg.context_dim)
    for i, b in enumerate(out):
        s += _spec_block(f"output_blocks.{i}.", b, te, cfg.context_dim)
    s += [("out.0.weight", (mc,), "norm_w"), ("out.0.bias", (mc,), "norm_b"),
          ("out.2.weight", (cfg.out_channels, mc, 3, 3), "w"), ("out.2.bias", (cfg.out_channels,), "b")]
    return s


HINT_CHANNELS = [16, 16, 32, 32, 96, 96, 256]   # cldm/cldm.py:147-163
HINT_STRIDES = [1, 1, 2, 1, 2, 1, 2, 1]


def _controlnet_param_spec(cfg: UNetConfig):
    inp, mid, _, chans = unet_layout(cfg)
    mc, te = cfg.model_channels, cfg.model_channels * 4
    s = [("time_embed.0.weight", (te, mc), "w"), ("time_embed.0.bias", (te,), "b"),
         ("time_embed.2.weight", (te, te), "w"), ("time_embed.2.bias", (te,), "b")]
    for i, b in enumerate(inp):
        s += _spec_block(f"input_blocks.{i}.", b, te, cfg.context_dim)
    for i, c in enumerate(chans):
        s += [(f"zero_convs.{i}.0.weight", (c, c, 1, 1), "w"), (f"zero_convs.{i}.0.bias", (c,), "b")]
    cin = cfg.hint_channels
    for i, cout in enumerate(HINT_CHANNELS + [mc]):
        s += [(f"input_hint_block.{2 * i}.weight", (cout, cin, 3, 3), "w"), (f"input_hint_block.{2 * i}.bias", (cout,), "b")]
        cin = cout
    s += _spec_block("middle_block.", mid, te, cfg.context_dim)
    c = chans[-1]
    s += [("middle_block_out.0.weight", (c, c, 1, 1), "w"), ("middle_block_out.0.bias", (c,), "b")]
    return s


def _spec_vae_res(p, cin, cout):
    s = [(p + "norm1.weight", (cin,), "norm_w"), (p + "norm1.bias", (cin,), "norm_b"),
         (p + "conv1.weight", (cout, cin, 3, 3), "w"), (p + "conv1.bias", (cout,), "b"),
         (p + "norm2.weight", (cout,), "norm_w"), (p + "norm2.bias", (cout,), "norm_b"),
         (p + "conv2.weight", (cout, cout, 3, 3), "w"), (p + "conv2.bias", (cout,), "b")]
    if cin != cout:
        s += [(p + "nin_shortcut.weight", (cout, cin, 1, 1), "w"), (p + "nin_shortcut.bias", (cout,), "b")]
    return s


def vae_layout(cfg: VAEConfig):
    """Decoder.__init__ (model.py:562-617): per level (processed from the deepest) the ResnetBlock channel pairs."""
    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[nres - 1]
    levels = {}
    for i_level in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels[i_level] = dict(blocks=blocks, upsample=(i_level != 0), c=block_in)
    return cfg.ch * cfg.ch_mult[nres - 1], levels, block_in


def vae_param_spec(cfg: VAEConfig):
    top, levels, last = vae_layout(cfg)
    z = cfg.z_channels
    s = [("post_quant_conv.weight", (z, z, 1, 1), "w"), ("post_quant_conv.bias", (z,), "b"),
         ("decoder.conv_in.weight", (top, z, 3, 3), "w"), ("decoder.conv_in.bias", (top,), "b")]
    s += _spec_vae_res("decoder.mid.block_1.", top, top)
    a = "decoder.mid.attn_1."
    s += [(a + "norm.weight", (top,), "norm_w"), (a + "norm.bias", (top,), "norm_b")]
    for n in ("q", "k", "v", "proj_out"):
        s += [(a + n + ".weight", (top, top, 1, 1), "w"), (a + n + ".bias", (top,), "b")]
    s += _spec_vae_res("decoder.mid.block_2.", top, top)
    for i_level, lv in levels.items():
        for j, (ci, co) in enumerate(lv["blocks"]):
            s += _spec_vae_res(f"decoder.up.{i_level}.block.{j}.", ci, co)
        if lv["upsample"]:
            c = lv["c"]
            s += [(f"decoder.up.{i_level}.upsample.conv.weight", (c, c, 3, 3), "w"),
                  (f"decoder.up.{i_level}.upsample.conv.bias", (c,), "b")]
    s += [("decoder.norm_out.weight", (last,), "norm_w"), ("decoder.norm_out.bias", (last,), "norm_b"),
          ("decoder.conv_out.weight", (cfg.out_ch, last, 3, 3), "w"), ("decoder.conv_out.bias", (cfg.out_ch,), "b")]
    return s


def make_weights(spec, seed: int, prefix: str = "") -> SD:
    """Deterministic synthetic weights, one Philox stream per tensor name (no dependence on construction order).
    Every tensor is non-zero — the reference's zero-initialised layers (openaimodel.py:228,731; attention.py:422;
    cldm.py:162,282) would make eps identically 0 and the parity check vacuous (SURVEY.md §7 'zero-init trap').
    Kinds: 'w' N(0, 1/fan_in); 'wk' N(0, 1/(3 fan_in)) = the variance of PyTorch's default kaiming-uniform init;
    'b'/'norm_b' N(0, 0.1^2); 'norm_w' 1 + N(0, 0.1^2)."""
    import zlib
    sd = {}
    for name, shape, kind in spec:
        key = (zlib.crc32((prefix + name).encode()) << 32) | (seed & 0xFFFFFFFF)
        rng = np.random.Generator(np.random.Philox(key=key))
        x = rng.standard_normal(shape, dtype=np.float32)
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            x *= 1.0 / math.sqrt(fan_in)
        elif kind == "wk":
            fan_in = int(np.prod(shape[1:]))
            x *= 1.0 / math.sqrt(3.0 * fan_in)
        elif kind == "b" or kind == "norm_b":
            x *= 0.1
        elif kind == "norm_w":
            x = 1.0 + 0.1 * x
        sd[name] = torch.from_numpy(np.ascontiguousarray(x))
    return sd


# ------------------------------------------------------------------------------------------------------------
# leaf ops
# ------------------------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """util.py:154-174 — [cos | sin] order."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=padding)


def _lin(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def _gn(sd, p, x, eps):
    """GroupNorm32 (util.py:217-219, eps 1e-5) / Normalize (attention.py:88-89, model.py:46-47, eps 1e-6)."""
    return F.group_norm(x.float(), 32, sd[p + "weight"], sd[p + "bias"], eps)


def resblock(sd, p, x, emb):
    """ResBlock._forward, non-updown, use_scale_shift_norm=False (openaimodel.py:255-275)."""
    h = _conv(sd, p + "in_layers.2.", F.silu(_gn(sd, p + "in_layers.0.", x, 1e-5)))
    emb_out = _lin(sd, p + "emb_layers.1.", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = _conv(sd, p + "out_layers.3.", F.silu(_gn(sd, p + "out_layers.0.", h, 1e-5)))
    if (p + "skip_connection.weight") in sd:
        x = _conv(sd, p + "skip_connection.", x, padding=0)
    return x + h


def cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, non-export branch (attention.py:217-250): fp32 scores, softmax(-1)."""
    q = F.linear(x, sd[p + "to_q.weight"])
    ctx = x if context is None else context
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    b, n, c = q.shape
    d = c // heads

    def split(t):  # 'b n (h d) -> (b h) n d'
        return t.view(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v)
    out = out.view(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return _lin(sd, p + "to_out.0.", out)


def cross_attention_fused(sd, p, x, context, heads):
    """The export branch's fused layout (attention.py:170,173,193-194,203-205): x @ cat([Wq,Wk,Wv]).T -> chunk(3);
    context @ cat([Wk,Wv]).T -> chunk(2). Same function as cross_attention; pinned by the reference's own script
    ldm_torch/modules/test_attention_onnx_torch_error.py:172-200 (allclose atol 1e-6)."""
    if context is None:
        qkv_w = torch.cat([sd[p + "to_q.weight"], sd[p + "to_k.weight"], sd[p + "to_v.weight"]]).transpose(0, 1)
        q, k, v = torch.matmul(x, qkv_w).chunk(3, dim=-1)
    else:
        q = F.linear(x, sd[p + "to_q.weight"])
        kv_w = torch.cat([sd[p + "to_k.weight"], sd[p + "to_v.weight"]]).transpose(0, 1)
        k, v = torch.matmul(context, kv_w).chunk(2, dim=-1)
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = (torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)).softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v).view(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return _lin(sd, p + "to_out.0.", out)


def transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward (attention.py:381-385) with GEGLU feed-forward (attention.py:49-76)."""
    c = x.shape[-1]

    def ln(i, t):
        return F.layer_norm(t, (c,), sd[p + f"norm{i}.weight"], sd[p + f"norm{i}.bias"], 1e-5)

    x = cross_attention(sd, p + "attn1.", ln(1, x), None, heads) + x
    x = cross_attention(sd, p + "attn2.", ln(2, x), context, heads) + x
    h = _lin(sd, p + "ff.net.0.proj.", ln(3, x))
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    return _lin(sd, p + "ff.net.2.", h) + x


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer.forward, use_linear=False (attention.py:431-450)."""
    b, c, hh, ww = x.shape
    x_in = x
    h = _gn(sd, p + "norm.", x, 1e-6)
    h = _conv(sd, p + "proj_in.", h, padding=0)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = transformer_block(sd, p + "transformer_blocks.0.", h, context, heads)
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    h = _conv(sd, p + "proj_out.", h, padding=0)
    return h + x_in


def run_block(sd, p, block: Block, x, emb, context):
    """TimestepEmbedSequential.forward (openaimodel.py:79-87)."""
    for j, (kind, a) in enumerate(block.layers):
        q = f"{p}{j}."
        if kind == "conv":
            x = _conv(sd, q, x)
        elif kind == "res":
            x = resblock(sd, q, x, emb)
        elif kind == "st":
            x = spatial_transformer(sd, q, x, context, a["heads"])
        elif kind == "down":
            x = _conv(sd, q + "op.", x, stride=2)          # Downsample (openaimodel.py:150-159)
        elif kind == "up":
            x = F.interpolate(x, scale_factor=2, mode="nearest")  # Upsample (openaimodel.py:108-118)
            x = _conv(sd, q + "conv.", x)
    return x


def _time_embed(sd, t, mc):
    emb = timestep_embedding(t, mc)
    return _lin(sd, "time_embed.2.", F.silu(_lin(sd, "time_embed.0.", emb)))


def hint_block(sd, hint):
    """input_hint_block (cldm/cldm.py:147-163): 8 convs, SiLU between, strides 1,1,2,1,2,1,2,1."""
    h = hint
    for i, s in enumerate(HINT_STRIDES):
        h = _conv(sd, f"input_hint_block.{2 * i}.", h, stride=s)
        if i != len(HINT_STRIDES) - 1:
            h = F.silu(h)
    return h


def controlnet_forward(sd, cfg: UNetConfig, x, hint, t, context):
    """ControlNet.forward (cldm/cldm.py:284-305) -> 13 tensors."""
    inp, mid, _, _ = unet_layout(cfg)
    emb = _time_embed(sd, t, cfg.model_channels)
    guided = hint_block(sd, hint)
    outs = []
    h = x
    for i, b in enumerate(inp):
        h = run_block(sd, f"input_blocks.{i}.", b, h, emb, context)
        if guided is not None:
            h = h + guided
            guided = None
        outs.append(_conv(sd, f"zero_convs.{i}.0.", h, padding=0))
    h = run_block(sd, "middle_block.", mid, h, emb, context)
    outs.append(_conv(sd, "middle_block_out.0.", h, padding=0))
    return outs


def unet_forward(sd, cfg: UNetConfig, x, t, context, control=None, only_mid_control=False):
    """ControlledUnetModel.forward (cldm/cldm.py:23-45); control=None gives UNetModel.forward
    (openaimodel.py:756-788). Does not mutate the caller's control list (the reference pops it)."""
    inp, mid, out, _ = unet_layout(cfg)
    control = list(control) if control is not None else None
    emb = _time_embed(sd, t, cfg.model_channels)
    hs = []
    h = x
    for i, b in enumerate(inp):
        h = run_block(sd, f"input_blocks.{i}.", b, h, emb, context)
        hs.append(h)
    h = run_block(sd, "middle_block.", mid, h, emb, context)
    if control is not None:
        h = h + control.pop()
    for i, b in enumerate(out):
        if only_mid_control or control is None:
            h = torch.cat([h, hs.pop()], dim=1)
        else:
            h = torch.cat([h, hs.pop() + control.pop()], dim=1)
        h = run_block(sd, f"output_blocks.{i}.", b, h, emb, context)
    h = F.silu(_gn(sd, "out.0.", h, 1e-5))
    return _conv(sd, "out.2.", h)


def apply_model(sd_unet, sd_cn, cfg, x_noisy, t, cond, control_scales=None, only_mid_control=False):
    """ControlLDM.apply_model (cldm/cldm.py:328-341)."""
    cond_txt = torch.cat(cond["c_crossattn"], 1)
    if cond["c_concat"] is None:
        return unet_forward(sd_unet, cfg, x_noisy, t, cond_txt, None, only_mid_control)
    scales = control_scales if control_scales is not None else [1.0] * 13
    control = controlnet_forward(sd_cn, cfg, x_noisy, torch.cat(cond["c_concat"], 1), t, cond_txt)
    control = [c * s for c, s in zip(control, scales)]
    return unet_forward(sd_unet, cfg, x_noisy, t, cond_txt, control, only_mid_control)


# ------------------------------------------------------------------------------------------------------------
# DDIM sampler
# ------------------------------------------------------------------------------------------------------------
def make_beta_schedule(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """util.py:21-25 ('linear' schedule; SD1.5 constants), float64."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def alphas_cumprod(n_timestep=1000):
    return np.cumprod(1.0 - make_beta_schedule(n_timestep), axis=0)


def ddim_schedule(S, eta=0.0, n_timestep=1000):
    """make_ddim_timesteps 'uniform' (util.py:46-60) + make_ddim_sampling_parameters (util.py:63-74) as used by
    DDIMSampler.make_schedule (ddim_hacked.py:23-52). The reference evaluates these on float32 alphas_cumprod
    (to_torch -> .cpu() numpy, ddim_hacked.py:28,42)."""
    c = n_timestep // S
    ts = np.asarray(list(range(0, n_timestep, c))) + 1
    ac = alphas_cumprod(n_timestep).astype(np.float32)
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return dict(timesteps=ts, alphas=alphas, alphas_prev=alphas_prev, sigmas=sigmas,
                sqrt_one_minus_alphas=np.sqrt(1.0 - alphas))


def ddim_update(x, e_t, a_t, a_prev, sigma_t, sqrt_one_minus_at, noise=None):
    """p_sample_ddim tail (ddim_hacked.py:208-231) for parameterization 'eps'."""
    pred_x0 = (x - sqrt_one_minus_at * e_t) / math.sqrt(a_t)
    dir_xt = math.sqrt(1.0 - a_prev - sigma_t ** 2) * e_t
    x_prev = math.sqrt(a_prev) * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma_t * noise
    return x_prev, pred_x0


def ddim_sample(eps_fn, x_T, cond, uncond, S=20, scale=9.0, eta=0.0, collect=False, noises=None, mask=None, x0=None,
                q_noises=None):
    """DDIMSampler.sample / ddim_sampling / p_sample_ddim (ddim_hacked.py:55-231).
    eps_fn(x, t, cond) plays model.apply_model; cond first, then uncond (ddim_hacked.py:190-191).
    eta > 0: noises[i] is the N(0,1) draw of step i (noise_like, ddim_hacked.py:227). mask / x0: the inpainting blend
    img = q_sample(x0, ts) * mask + (1 - mask) * img before every step (ddim_hacked.py:154-157) with q_sample(x0, t) =
    sqrt(abar_t) x0 + sqrt(1 - abar_t) q_noises[i] (LatentDiffusion.q_sample, ddpm.py -- absent from the reference
    checkout; the standard definition)."""
    assert eta == 0.0 or noises is not None, "eta > 0 needs the per-step noise draws"
    sch = ddim_schedule(S, eta)
    ac = alphas_cumprod().astype(np.float32)
    img = x_T
    b = x_T.shape[0]
    trace = []
    for i, step in enumerate(np.flip(sch["timesteps"])):
        index = S - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        if mask is not None:
            img_orig = math.sqrt(float(ac[int(step)])) * x0 + math.sqrt(1.0 - float(ac[int(step)])) * q_noises[i]
            img = img_orig * mask + (1.0 - mask) * img
        if uncond is None or scale == 1.0:
            e_t = eps_fn(img, ts, cond)
        else:
            e_c = eps_fn(img, ts, cond)
            e_u = eps_fn(img, ts, uncond)
            e_t = e_u + scale * (e_c - e_u)
        x_in = img
        img, pred_x0 = ddim_update(img, e_t, float(sch["alphas"][index]), float(sch["alphas_prev"][index]),
                                   float(sch["sigmas"][index]), float(sch["sqrt_one_minus_alphas"][index]),
                                   noise=noises[i] if eta != 0.0 else None)
        if collect:
            trace.append(dict(t=int(step), x_in=x_in, eps=e_t, x_prev=img))
    return img, trace


def ddim_encode(eps_fn, x0, cond, t_enc, S=20, scale=1.0, uncond=None):
    """DDIMSampler.encode (ddim_hacked.py:233-276), use_original_steps=False: deterministic DDIM inversion over the first
    t_enc DDIM timesteps. The reference's CFG branch concatenates conditionings with torch.cat, which only works for tensor
    conditionings; for the ControlNet dict conditioning the two branches are evaluated separately (same arithmetic)."""
    sch = ddim_schedule(S)
    alphas_next, alphas = sch["alphas"][:t_enc], sch["alphas_prev"][:t_enc]
    x_next = x0
    for i in range(t_enc):
        t = torch.full((x0.shape[0],), int(sch["timesteps"][i]), dtype=torch.long)
        if scale == 1.0:
            noise_pred = eps_fn(x_next, t, cond)
        else:
            e_u, e_c = eps_fn(x_next, t, uncond), eps_fn(x_next, t, cond)
            noise_pred = e_u + scale * (e_c - e_u)
        an, a = float(alphas_next[i]), float(alphas[i])
        x_next = math.sqrt(an / a) * x_next + math.sqrt(an) * (math.sqrt(1 / an - 1) - math.sqrt(1 / a - 1)) * noise_pred
    return x_next


def stochastic_encode(x0, t_index, S=20, noise=None):
    """DDIMSampler.stochastic_encode (ddim_hacked.py:278-292), use_original_steps=False: t_index [B] int64 indexes the
    DDIM tables."""
    sch = ddim_schedule(S)
    a = torch.tensor(sch["alphas"], dtype=torch.float32)[t_index].reshape(-1, 1, 1, 1)
    return a.sqrt() * x0 + (1.0 - a).sqrt() * noise


def ddim_decode(eps_fn, x_latent, cond, t_start, S=20, scale=1.0, uncond=None):
    """DDIMSampler.decode (ddim_hacked.py:294-317): p_sample_ddim over the first t_start DDIM timesteps, reversed."""
    sch = ddim_schedule(S)
    x = x_latent
    for i, step in enumerate(np.flip(sch["timesteps"][:t_start])):
        index = t_start - i - 1
        ts = torch.full((x.shape[0],), int(step), dtype=torch.long)
        if uncond is None or scale == 1.0:
            e_t = eps_fn(x, ts, cond)
        else:
            e_c, e_u = eps_fn(x, ts, cond), eps_fn(x, ts, uncond)
            e_t = e_u + scale * (e_c - e_u)
        x, _ = ddim_update(x, e_t, float(sch["alphas"][index]), float(sch["alphas_prev"][index]), 0.0,
                           float(sch["sqrt_one_minus_alphas"][index]))
    return x


# ------------------------------------------------------------------------------------------------------------
# VAE decoder
# ------------------------------------------------------------------------------------------------------------
def vae_resblock(sd, p, x):
    """ResnetBlock.forward with temb=None (model.py:129-149)."""
    h = _conv(sd, p + "conv1.", F.silu(_gn(sd, p + "norm1.", x, 1e-6)))
    h = _conv(sd, p + "conv2.", F.silu(_gn(sd, p + "norm2.", h, 1e-6)))
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(sd, p + "nin_shortcut.", x, padding=0)
    return x + h


def vae_attn(sd, p, x):
    """AttnBlock.forward (model.py:179-203): single head, d = C."""
    h = _gn(sd, p + "norm.", x, 1e-6)
    q, k, v = (_conv(sd, p + n + ".", h, padding=0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** -0.5)
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out.", h, padding=0)


def vae_decode(sd, cfg: VAEConfig, z):
    """decode_first_stage as described at canny2image_torch.py:63-67 (ddpm.py is absent from the reference):
    z / scale_factor -> post_quant_conv -> Decoder.forward (model.py:619-652)."""
    _, levels, _ = vae_layout(cfg)
    h = _conv(sd, "post_quant_conv.", z / cfg.scale_factor, padding=0)
    p = "decoder."
    h = _conv(sd, p + "conv_in.", h)
    h = vae_resblock(sd, p + "mid.block_1.", h)
    h = vae_attn(sd, p + "mid.attn_1.", h)
    h = vae_resblock(sd, p + "mid.block_2.", h)
    for i_level in reversed(range(len(cfg.ch_mult))):
        lv = levels[i_level]
        for j in range(cfg.num_res_blocks + 1):
            h = vae_resblock(sd, f"{p}up.{i_level}.block.{j}.", h)
        if lv["upsample"]:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"{p}up.{i_level}.upsample.conv.", h)
    h = F.silu(_gn(sd, p + "norm_out.", h, 1e-6))
    return _conv(sd, p + "conv_out.", h)


def vae_encoder_param_spec(cfg: VAEConfig):
    """Encoder.__init__ (model.py:452-512) + AutoencoderKL.quant_conv (autoencoder.py, absent: Conv2d(2 z, 2 embed, 1))."""
    nres = len(cfg.ch_mult)
    z = cfg.z_channels
    in_mult = (1,) + tuple(cfg.ch_mult)
    s = [("quant_conv.weight", (2 * z, 2 * z, 1, 1), "w"), ("quant_conv.bias", (2 * z,), "b"),
         ("encoder.conv_in.weight", (cfg.ch, 3, 3, 3), "w"), ("encoder.conv_in.bias", (cfg.ch,), "b")]
    block_in = cfg.ch
    for i_level in range(nres):
        block_in = cfg.ch * in_mult[i_level]
        block_out = cfg.ch * cfg.ch_mult[i_level]
        for j in range(cfg.num_res_blocks):
            s += _spec_vae_res(f"encoder.down.{i_level}.block.{j}.", block_in, block_out)
            block_in = block_out
        if i_level != nres - 1:
            s += [(f"encoder.down.{i_level}.downsample.conv.weight", (block_in, block_in, 3, 3), "w"),
                  (f"encoder.down.{i_level}.downsample.conv.bias", (block_in,), "b")]
    s += _spec_vae_res("encoder.mid.block_1.", block_in, block_in)
    a = "encoder.mid.attn_1."
    s += [(a + "norm.weight", (block_in,), "norm_w"), (a + "norm.bias", (block_in,), "norm_b")]
    for n in ("q", "k", "v", "proj_out"):
        s += [(a + n + ".weight", (block_in, block_in, 1, 1), "w"), (a + n + ".bias", (block_in,), "b")]
    s += _spec_vae_res("encoder.mid.block_2.", block_in, block_in)
    s += [("encoder.norm_out.weight", (block_in,), "norm_w"), ("encoder.norm_out.bias", (block_in,), "norm_b"),
          ("encoder.conv_out.weight", (2 * z, block_in, 3, 3), "w"), ("encoder.conv_out.bias", (2 * z,), "b")]
    return s


def vae_encode(sd, cfg: VAEConfig, x):
    """Encoder.forward (model.py:514-543) -> quant_conv: the moments [B, 2 z, h/8, w/8] = (mean | logvar) of the
    DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-35). Downsample = F.pad(x, (0,1,0,1)) +
    conv3x3 stride 2 without padding (model.py:78-84). The latent is scale_factor * mean (mode) or a sample."""
    p = "encoder."
    h = _conv(sd, p + "conv_in.", x)
    nres = len(cfg.ch_mult)
    for i_level in range(nres):
        for j in range(cfg.num_res_blocks):
            h = vae_resblock(sd, f"{p}down.{i_level}.block.{j}.", h)
        if i_level != nres - 1:
            h = _conv(sd, f"{p}down.{i_level}.downsample.conv.", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = vae_resblock(sd, p + "mid.block_1.", h)
    h = vae_attn(sd, p + "mid.attn_1.", h)
    h = vae_resblock(sd, p + "mid.block_2.", h)
    h = _conv(sd, p + "conv_out.", F.silu(_gn(sd, p + "norm_out.", h, 1e-6)))
    return _conv(sd, "quant_conv.", h, padding=0)


def to_uint8_image(x):
    """canny2image_torch.py:68 — 'b c h w -> b h w c', *127.5 + 127.5, clip, uint8."""
    y = (x.permute(0, 2, 3, 1) * 127.5 + 127.5).numpy().clip(0, 255).astype(np.uint8)
    return y


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs shared by the oracle, the golden generator and the CUDA tests
# ------------------------------------------------------------------------------------------------------------
def make_inputs(cfg: UNetConfig, batch: int, h: int, w: int, hint: Optional[Tensor] = None):
    """x_T seed 2946901 (compute_score_torch.py:37); contexts seeds 1 (cond) / 2 (uncond); hint: a binary edge map,
    three equal channels, values {0,1} (canny2image_torch.py:33-38) from seed 7 unless given."""
    g = torch.Generator().manual_seed(2946901)
    x_T = torch.randn((batch, cfg.in_channels, h, w), generator=g)
    ctx_c = torch.randn((batch, 77, cfg.context_dim), generator=torch.Generator().manual_seed(1))
    ctx_u = torch.randn((batch, 77, cfg.context_dim), generator=torch.Generator().manual_seed(2))
    if hint is None:
        r = torch.rand((batch, 1, 8 * h, 8 * w), generator=torch.Generator().manual_seed(7))
        hint = (r > 0.9).float().expand(-1, 3, -1, -1).contiguous()
    cond = {"c_concat": [hint], "c_crossattn": [ctx_c]}
    uncond = {"c_concat": [hint], "c_crossattn": [ctx_u]}
    return x_T, cond, uncond
