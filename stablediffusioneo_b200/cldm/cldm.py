"""ControlledUnetModel / ControlNet / ControlLDM with the reference's names, constructor arguments, forward
signatures and state-dict keys (cldm/cldm.py), running on libsdeo.so.

The hot entry point is ControlLDM.apply_model(x_noisy, t, cond) (cldm/cldm.py:328-341). Internally it reorders the
two independent halves — UNet encoder first, then ControlNet — so that every ControlNet zero-conv adds its scaled
output straight onto the UNet skip it controls (`hs.pop() + control.pop()`, cldm/cldm.py:41, and `h += control.pop()`,
:35) in the conv epilogue: no 13 scale kernels (:338), no 13 add kernels."""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..ldm.modules.attention import SpatialTransformer
from ..ldm.modules.diffusionmodules.openaimodel import (Downsample, ResBlock, TimestepEmbedSequential, UNetModel,
                                                        _check_supported)
from ..ldm.modules.diffusionmodules import util
from ..ldm.modules.diffusionmodules.util import (BF16, CatPair, SiLU, conv_nd, is_internal, linear, make_beta_schedule,
                                                 nchw_view, nhwc, operand, timestep_embedding, to_external, to_internal,
                                                 zero_module)
from ..ldm.modules.diffusionmodules.model import Decoder, Encoder


def _ctx_internal(context):
    return context.contiguous() if context.dtype == BF16 else ops.to_bf16(context.float())


class ControlledUnetModel(UNetModel):
    def run_encoder(self, x, emb, context):
        hs = []
        h = x
        for module in self.input_blocks:
            h = module.run(h, emb, context)
            hs.append(h)
        h = self.middle_block.run(h, emb, context)
        return hs, h

    def run_decoder(self, h, hs, emb, context, before_block=None):
        """before_block(k): optional hook called before output block k reads its skip tensor hs[len(hs) - 1 - k]
        (the step engine waits there for the zero conv that produced it on the side stream)."""
        hs = list(hs)
        for k, module in enumerate(self.output_blocks):
            if before_block is not None:
                before_block(k)
            h = module.run(CatPair(h, hs.pop()), emb, context)
        return self.run_out(h)

    def run(self, x, emb, context, control=None, only_mid_control=False):
        """cldm/cldm.py:23-45 on internal tensors; `control` (13 internal tensors, already scaled) is consumed."""
        hs, h = self.run_encoder(x, emb, context)
        if control is not None:
            h = ops_add(h, control.pop())
            if not only_mid_control:
                hs = [ops_add(s, c) for s, c in zip(hs, control)]
                del control[:]
        return self.run_decoder(h, hs, emb, context)

    def forward(self, x, timesteps=None, context=None, control=None, only_mid_control=False, **kwargs):
        ctx = _ctx_internal(context)
        emb = self.embed_time(timesteps)
        ctl = None
        if control is not None:
            ctl = [to_internal(c) for c in control]
            del control[:]  # the reference pops the caller's list empty (cldm/cldm.py:35-41)
        return to_external(self.run(to_internal(x), emb, ctx, ctl, only_mid_control), self.out_channels)


def ops_add(a, b):
    """a + b on internal tensors of either kind (the generic, unfused injection path): bf16 result."""
    return nchw_view(ops.add_scaled(nhwc(operand(a)), nhwc(operand(b)), 1.0))


class ControlNet(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, hint_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, use_checkpoint=False, use_fp16=False,
                 num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=False,
                 transformer_depth=1, context_dim=None, n_embed=None, legacy=True, disable_self_attentions=None,
                 num_attention_blocks=None, disable_middle_self_attn=False, use_linear_in_transformer=False):
        super().__init__()
        _check_supported(dims, None, use_scale_shift_norm, resblock_updown, use_spatial_transformer, context_dim,
                         n_embed, num_heads, num_head_channels, disable_self_attentions, num_attention_blocks,
                         use_linear_in_transformer, conv_resample)
        if type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)
        self.dims = dims
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.num_res_blocks = len(channel_mult) * [num_res_blocks] if isinstance(num_res_blocks, int) else list(num_res_blocks)
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.predict_codebook_ids = False

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), SiLU(),
                                        linear(time_embed_dim, time_embed_dim))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        self.zero_convs = nn.ModuleList([self.make_zero_conv(model_channels)])

        # hint encoder (cldm/cldm.py:147-163): 3->16->16->32(s2)->32->96(s2)->96->256(s2)->model_channels, SiLU between
        widths, strides = [16, 16, 32, 32, 96, 96, 256], [1, 1, 2, 1, 2, 1, 2]
        layers, cin = [], hint_channels
        for wd, s in zip(widths, strides):
            layers += [conv_nd(dims, cin, wd, 3, padding=1, stride=s), SiLU()]
            cin = wd
        layers.append(zero_module(conv_nd(dims, cin, model_channels, 3, padding=1)))
        self.input_hint_block = TimestepEmbedSequential(*layers)

        def res(cin_, cout_):
            return ResBlock(cin_, time_embed_dim, dropout, out_channels=cout_, dims=dims, use_checkpoint=use_checkpoint)

        def st(c):
            return SpatialTransformer(c, num_heads, c // num_heads, depth=transformer_depth, context_dim=context_dim,
                                      disable_self_attn=False, use_linear=False, use_checkpoint=use_checkpoint)

        self._feature_size = model_channels
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(self.num_res_blocks[level]):
                blk = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    blk.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*blk))
                self.zero_convs.append(self.make_zero_conv(ch))
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                self.zero_convs.append(self.make_zero_conv(ch))
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), st(ch), res(ch, ch))
        self.middle_block_out = self.make_zero_conv(ch)

    def make_zero_conv(self, channels):
        return TimestepEmbedSequential(zero_module(conv_nd(self.dims, channels, channels, 1, padding=0)))

    # ---- pieces --------------------------------------------------------------------------------------------
    def embed_time(self, timesteps):
        t_emb = timestep_embedding(timesteps, self.model_channels, repeat_only=False)
        return self.time_embed[2].run(self.time_embed[0].run(t_emb, act=ops.SDEO_ACT_SILU))

    def run_hint(self, hint):
        """input_hint_block on an internal hint tensor; each SiLU is fused into the preceding conv's epilogue.
        Depends only on the hint: callers hoist it out of the denoising loop (the reference recomputes it 40x/image)."""
        h = hint
        layers = list(self.input_hint_block)
        i = 0
        while i < len(layers):
            conv = layers[i]
            fused = i + 1 < len(layers) and isinstance(layers[i + 1], SiLU)
            last = i + (2 if fused else 1) >= len(layers)
            h = conv.run(h, act=ops.SDEO_ACT_SILU if fused else ops.SDEO_ACT_NONE,
                         out_fp32=last and util.STREAM_FP32)   # the result is a residual term of conv_in's epilogue
            i += 2 if fused else 1
        return h

    def run_body(self, x, guided_hint, emb, context):
        """The encoder copy without its zero convs: returns the 13 feature maps h_i the zero convs read
        (12 input blocks + middle block). Independent of the UNet, so it can run on its own stream."""
        feats = []
        h = x
        for i, module in enumerate(self.input_blocks):
            if i == 0:
                # h = conv_in(x) + guided_hint (cldm/cldm.py:294-297): residual add in the conv epilogue
                h = module[0].run(h, residual=guided_hint, stream=util.STREAM_FP32, gn_stats=True)
            else:
                h = module.run(h, emb, context)
            feats.append(h)
        feats.append(self.middle_block.run(h, emb, context))
        return feats

    def run_zero_convs(self, feats, scales=None, add_to=None, only_mid=False, order=None, after=None, rows=None):
        """13 outputs: scale_i * zero_conv_i(h_i) [+ add_to[i]] (only_mid: entries 0..11 are add_to[i] untouched).
        order: launch order of the 13 convs (default 0..12); after(i): hook called once output i has been enqueued.
        rows (guess mode in the step engine, needs add_to): the feature maps hold only the first `rows` samples of the
        batch; zero_conv_i(h_i) is added IN PLACE onto those rows of the stream tensor add_to[i] (and of its bf16 twin),
        the other rows stay as they are; outs[i] is add_to[i] itself (without epilogue statistics: the rows' producers
        differ)."""
        scales = [1.0] * (len(self.zero_convs) + 1) if scales is None else list(scales)
        convs = [z[0] for z in self.zero_convs] + [self.middle_block_out[0]]
        outs = [None] * len(convs)
        for i in (range(len(convs)) if order is None else order):
            conv, h = convs[i], feats[i]
            last = i == len(convs) - 1
            if only_mid and add_to is not None and not last:
                outs[i] = add_to[i]
            elif rows is not None:
                tgt = add_to[i]
                assert util.is_stream(tgt) and tgt._twin is not None and h.shape[0] == rows
                f32, twin = nhwc(tgt)[:rows], nhwc(tgt._twin)[:rows]
                conv.run(h, scale=scales[i], residual=nchw_view(f32), stream=True, out_f32=f32, out_twin=twin)
                tgt._gn_stats = None
                outs[i] = tgt
            else:
                outs[i] = conv.run(h, scale=scales[i], residual=add_to[i] if add_to is not None else None,
                                   stream=util.STREAM_FP32, gn_stats=add_to is not None)
            if after is not None:
                after(i)
        return outs

    def run(self, x, guided_hint, emb, context, scales=None, add_to=None, only_mid=False):
        """cldm/cldm.py:284-305 on internal tensors."""
        return self.run_zero_convs(self.run_body(x, guided_hint, emb, context), scales, add_to, only_mid)

    def forward(self, x, hint, timesteps, context, **kwargs):
        ctx = _ctx_internal(context)
        emb = self.embed_time(timesteps)
        guided = self.run_hint(to_internal(hint))
        outs = self.run(to_internal(x), guided, emb, ctx)
        return [to_external(o) for o in outs]


class DiffusionWrapper(nn.Module):
    """`model.diffusion_model` holder — keeps the checkpoint prefix `model.diffusion_model.*` (ddpm.py, absent)."""

    def __init__(self, diffusion_model):
        super().__init__()
        self.diffusion_model = diffusion_model


class FirstStage(nn.Module):
    """AutoencoderKL (ldm/models/autoencoder.py, absent from the reference checkout): the decode half
    `first_stage_model.post_quant_conv.*`, `first_stage_model.decoder.*` always; with_encoder adds
    `first_stage_model.encoder.*` and `first_stage_model.quant_conv.*` (img2img / inpainting callers)."""

    def __init__(self, ddconfig, embed_dim=4, with_encoder=False):
        super().__init__()
        if with_encoder:
            self.encoder = Encoder(**ddconfig)
            self.quant_conv = conv_nd(2, 2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = conv_nd(2, embed_dim, ddconfig["z_channels"], 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def encode_moments(self, x):
        """image fp32 NCHW in [-1, 1] -> moments fp32 [B, 2 embed, h/8, w/8] = (mean | logvar), i.e. the parameters of
        AutoencoderKL.encode's DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-35)."""
        if not hasattr(self, "encoder"):
            raise RuntimeError("FirstStage was built without the encoder: ControlLDM(first_stage_encoder=True)")
        m = self.quant_conv.run(self.encoder.run(to_internal(x)), out_fp32=True)
        return to_external(m, self.quant_conv.out_channels)


SD15_UNET_KW = dict(image_size=32, in_channels=4, model_channels=320, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                    channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                    context_dim=768, use_checkpoint=False, legacy=False)
SD15_VAE_KW = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                   num_res_blocks=2, attn_resolutions=[], dropout=0.0)


class ControlLDM(nn.Module):
    """The object canny2image_torch.py samples from. The reference derives it from ldm.models.diffusion.ddpm
    .LatentDiffusion, which is absent from the reference tree; the attributes DDIMSampler and the
    pipeline need are provided here: schedule buffers (linear, 1000 steps, 0.00085..0.012), `parameterization`,
    `scale_factor`, `apply_model`, `decode_first_stage`, `control_scales`, `only_mid_control`."""

    def __init__(self, unet_config=None, control_stage_config=None, first_stage_config=None, control_key="hint",
                 only_mid_control=False, timesteps=1000, linear_start=0.00085, linear_end=0.012, scale_factor=0.18215,
                 parameterization="eps", cond_stage_config=None, first_stage_encoder=False):
        super().__init__()
        unet_kw = dict(SD15_UNET_KW if unet_config is None else unet_config)
        cn_kw = dict(control_stage_config) if control_stage_config is not None else dict(unet_kw, hint_channels=3)
        self.model = DiffusionWrapper(ControlledUnetModel(out_channels=unet_kw.pop("out_channels", 4), **unet_kw))
        cn_kw.pop("out_channels", None)
        self.control_model = ControlNet(**cn_kw)
        self.first_stage_model = FirstStage(dict(SD15_VAE_KW if first_stage_config is None else first_stage_config),
                                            with_encoder=first_stage_encoder)
        # optional text encoder (`cond_stage_model.*` checkpoint keys): cond_stage_config = {} builds the CLIP ViT-L/14
        # text tower of cldm_v15.yaml; None (default) leaves prompt encoding to the caller
        self.cond_stage_model = None
        if cond_stage_config is not None:
            from ..ldm.modules.encoders.modules import FrozenCLIPEmbedder
            self.cond_stage_model = FrozenCLIPEmbedder(**dict(cond_stage_config))
        self.control_key = control_key
        self.only_mid_control = only_mid_control
        self.control_scales = [1.0] * 13
        # "bf16" (default: bf16 GEMM operands, fp32 accumulation / residual stream; 1e-2 gate) or "fp32" (split-operand
        # GEMMs + fp32 everything else, stablediffusioneo_b200/precise.py; 1e-4 gate, verification speed)
        self.precision = "bf16"
        self.parameterization = parameterization
        self.scale_factor = scale_factor
        self.channels = unet_kw["in_channels"]
        self.num_timesteps = int(timesteps)
        self._schedule_args = (int(timesteps), linear_start, linear_end)
        self.register_schedule()
        self._hint_cache = None
        def _drop_param_cache(module, incompatible_keys):  # (a post-hook must return None)
            module.__dict__.pop("_fp_params", None)
        self.register_load_state_dict_post_hook(_drop_param_cache)
        # first eager call of each layer shape picks its tile / split-K configuration (SDEO_NO_AUTOTUNE=1: heuristics only)
        ops.set_autotune(not os.environ.get("SDEO_NO_AUTOTUNE"))

    def register_schedule(self, device=None):
        """The DDPM schedule buffers DDIMSampler reads (ddpm.py's register_schedule: linear betas, cumulative alphas).
        Non-persistent, so they are not checkpoint keys; callable again after constructing the module on `meta`."""
        timesteps, linear_start, linear_end = self._schedule_args
        betas = make_beta_schedule("linear", timesteps, linear_start=linear_start, linear_end=linear_end)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        f32 = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
        self.register_buffer("betas", f32(betas), persistent=False)
        self.register_buffer("alphas_cumprod", f32(alphas_cumprod), persistent=False)
        self.register_buffer("alphas_cumprod_prev", f32(np.append(1.0, alphas_cumprod[:-1])), persistent=False)
        self.register_buffer("sqrt_alphas_cumprod", f32(np.sqrt(alphas_cumprod)), persistent=False)
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1.0 - alphas_cumprod)), persistent=False)

    @property
    def device(self):
        return self.betas.device

    def weights_fingerprint(self, first_stage=False):
        """Identity + version of every parameter a captured engine bakes in (packed / folded weights, time-embedding
        tables, hoisted K/V, the hint features): changes after load_state_dict(), an optimizer step or any other in-place
        update, and when a parameter is replaced. Engines and caches key on it so that they never replay stale weights."""
        cache = self.__dict__.get("_fp_params")
        if cache is None:
            # the parameter list is cached (walking the module tree costs ~1 ms per call); load_state_dict -- the one
            # supported way of replacing Parameter objects (assign=True) -- drops the cache through a post-hook
            cache = self.__dict__["_fp_params"] = (
                list(self.model.parameters()) + list(self.control_model.parameters()),
                list(self.first_stage_model.parameters()))
        params = cache[0] + cache[1] if first_stage else cache[0]
        return hash(tuple([p._version for p in params] + [p.data_ptr() for p in params]))

    # ---- hoisted, loop-invariant pieces ------------------------------------------------------------------
    def guided_hint(self, hint):
        """input_hint_block(hint), cached while the caller keeps passing the same (unmodified) hint tensor object."""
        c = self._hint_cache
        wkey = util._param_key(*self.control_model.input_hint_block.parameters())
        if c is None or c[0]() is not hint or c[1] != hint._version or c[3] != wkey:
            import weakref
            self._hint_cache = (weakref.ref(hint), hint._version, self.control_model.run_hint(to_internal(hint)), wkey)
        return self._hint_cache[2]

    def eps_internal(self, x, timesteps, ctx, guided_hint):
        """One ControlNet+UNet pass on internal tensors. x [N,8(4 used),h,w] bf16; returns eps fp32 NHWC-physical
        [N, 4, h, w]-shaped view. guided_hint None = UNet only (cond['c_concat'] is None, cldm/cldm.py:334-335)."""
        unet = self.model.diffusion_model
        emb_u = unet.embed_time(timesteps)
        hs, h = unet.run_encoder(x, emb_u, ctx)
        if guided_hint is not None:
            emb_c = self.control_model.embed_time(timesteps)
            outs = self.control_model.run(x, guided_hint, emb_c, ctx, scales=self.control_scales, add_to=hs + [h],
                                          only_mid=self.only_mid_control)
            hs, h = outs[:-1], outs[-1]
        return unet.run_decoder(h, hs, emb_u, ctx)

    def apply_model(self, x_noisy, t, cond, *args, **kwargs):
        """cldm/cldm.py:328-341: eps = UNet(x, t, ctx, control = scales * ControlNet(x, hint, t, ctx))."""
        assert isinstance(cond, dict)
        if self.precision == "fp32":
            from .. import precise
            return precise.eps(self, x_noisy, t, cond)
        if self.precision != "bf16":
            raise ValueError(f"ControlLDM.precision must be 'bf16' or 'fp32', got {self.precision!r}")
        cond_txt = cond["c_crossattn"]
        cond_txt = cond_txt[0] if len(cond_txt) == 1 else torch.cat(cond_txt, 1)
        ctx = _ctx_internal(cond_txt)
        guided = None
        if cond["c_concat"] is not None:
            hint = cond["c_concat"][0] if len(cond["c_concat"]) == 1 else torch.cat(cond["c_concat"], 1)
            guided = self.guided_hint(hint)
        eps = self.eps_internal(to_internal(x_noisy), t, ctx, guided)
        return to_external(eps, self.model.diffusion_model.out_channels)

    @torch.no_grad()
    def get_learned_conditioning(self, c):
        """Prompt(s) (token ids [B, 77], or strings when the tokenizer files are on disk) -> context [B, 77, 768]
        (ddpm.py's get_learned_conditioning -> cond_stage_model.encode, canny2image_torch.py:47,52)."""
        if self.cond_stage_model is None:
            raise RuntimeError("ControlLDM was built without a text encoder: pass cond_stage_config={}")
        return self.cond_stage_model.encode(c)

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """Forward diffusion sqrt(abar_t) x0 + sqrt(1 - abar_t) noise (LatentDiffusion.q_sample, ddpm.py -- absent from
        the reference checkout; called by DDIMSampler's mask blending, cldm/ddim_hacked.py:156). t int64 [B]."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return ops.axpby(x_start.float().contiguous(), noise.float().contiguous(),
                         self.sqrt_alphas_cumprod[t], self.sqrt_one_minus_alphas_cumprod[t])

    @torch.no_grad()
    def encode_first_stage(self, x, sample=False, noise=None):
        """image fp32 NCHW in [-1, 1] -> latent z = scale_factor * (mode or sample of the encoder posterior)
        (encode_first_stage + get_first_stage_encoding of ddpm.py; logvar clamped to [-30, 20] like
        DiagonalGaussianDistribution). Returns fp32 [B, 4, h/8, w/8]."""
        m = self.first_stage_model.encode_moments(x)
        zc = m.shape[1] // 2
        mean = m[:, :zc].contiguous()
        if not sample:
            return ops.axpby(mean, mean, self.scale_factor, 0.0)
        std = torch.exp(0.5 * m[:, zc:].clamp(-30.0, 20.0)).contiguous()
        noise = torch.randn_like(mean) if noise is None else noise.to(mean)
        return ops.axpby(mean, (std * noise).contiguous(), self.scale_factor, self.scale_factor)

    @torch.no_grad()
    def decode_first_stage(self, z):
        """z / scale_factor -> post_quant_conv -> Decoder (canny2image_torch.py:63-67). fp32 NCHW in [-1, 1]."""
        fs = self.first_stage_model
        zi = to_internal(z, scale=1.0 / self.scale_factor)
        return to_external(fs.decoder.run(fs.post_quant_conv.run(zi)), fs.decoder.out_ch)

    def _decode_u8_eager(self, z):
        fs = self.first_stage_model
        zi = to_internal(z, scale=1.0 / self.scale_factor)
        img = fs.decoder.run(fs.post_quant_conv.run(zi))
        return ops.image_to_u8(nhwc(img), fs.decoder.out_ch)

    def decode_first_stage_u8(self, z):
        """decode + 'b c h w -> b h w c' * 127.5 + 127.5, clip, uint8 (canny2image_torch.py:68) on the device.
        The decoder is ~65 kernel launches; from the third call with the same latent shape (and unchanged VAE weights) they
        are replayed as ONE captured CUDA graph (SDEO_NO_VAE_GRAPH=1: always eager). The result is a fresh tensor."""
        if os.environ.get("SDEO_NO_VAE_GRAPH") or not z.is_cuda or torch.cuda.is_current_stream_capturing():
            return self._decode_u8_eager(z)
        cache = self.__dict__.setdefault("_vae_graphs", {})
        fsp = self.__dict__.get("_fp_params")
        if fsp is None:
            self.weights_fingerprint(first_stage=True)
            fsp = self.__dict__["_fp_params"]
        wkey = hash(tuple([p._version for p in fsp[1]] + [p.data_ptr() for p in fsp[1]]))
        key = (tuple(z.shape), z.dtype, z.device.index)
        ent = cache.get(key)
        if ent is not None and ent["wkey"] != wkey:
            ent = None
        if ent is None:
            ent = cache[key] = {"wkey": wkey, "calls": 0, "graph": None}
        ent["calls"] += 1
        if ent["graph"] is None:
            if ent["calls"] < 3:   # first call packs weights / tunes the conv shapes, second confirms the steady state
                return self._decode_u8_eager(z)
            ent["z"] = z.detach().clone()
            torch.cuda.synchronize(z.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ent["out"] = self._decode_u8_eager(ent["z"])
            ent["graph"] = g
        ent["z"].copy_(z, non_blocking=True)
        ent["graph"].replay()
        return ent["out"].clone()
