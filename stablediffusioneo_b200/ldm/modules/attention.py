"""SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU / FeedForward with the reference's names,
constructor arguments and state-dict keys (ldm/modules/attention.py), running on libsdeo.so.

CrossAttention keeps the reference's fused layout — `qkv_w = cat([Wq, Wk, Wv]).T` for self-attention and
`kv_w = cat([Wk, Wv]).T` for cross-attention (attention.py:170,173,193-194,204-205) — as the ONE projection GEMM
per attention, whose epilogue scatters straight into the head-major q/k and transposed-v buffers the flash kernel
reads. Unlike the reference, the fused tensors are rebuilt from the *current* weights (the reference builds them
from init-time weights and never refreshes them — README.md:69-73 — a bug we do not reproduce)."""
import torch
import torch.nn as nn

from ... import ops
from .diffusionmodules import util
from .diffusionmodules.util import (BF16, Conv2d, GroupNorm32, LayerNorm, Linear, _param_key, boundary, nhwc,
                                    nchw_view, operand, tokens_boundary, zero_module)


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else (d() if callable(d) else d)


def Normalize(in_channels):
    return GroupNorm32(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class GEGLU(nn.Module):
    """proj -> chunk(2) -> x * gelu(gate) (attention.py:49-56); the product is fused into the GEMM epilogue."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = Linear(dim_in, dim_out * 2)
        self._cache = {}

    def _packed(self):
        key = _param_key(self.proj.weight, self.proj.bias)
        hit = self._cache.get("w")
        if hit is None or hit[0] != key:
            pw = ops.pack_conv_weight(self.proj.weight.detach(), geglu=True)
            pb = ops.pack_geglu_bias(self.proj.bias.detach(), pw.geglu_bn)
            self._cache["w"] = (key, pw, pb)
            hit = self._cache["w"]
        return hit[1], hit[2]

    def _packed_ln(self, ln):
        key = _param_key(self.proj.weight, self.proj.bias, ln.weight, ln.bias)
        hit = self._cache.get("ln")
        if hit is None or hit[0] != key:
            self._cache["ln"] = (key,) + util.fold_ln_weight(self.proj.weight, self.proj.bias, ln, geglu=True)
            hit = self._cache["ln"]
        return hit[1:]

    def run(self, x):
        if isinstance(x, util.DeferredLN):  # LayerNorm folded into this projection
            pw, pb, csum = self._packed_ln(x.ln)
            return ops.linear(x.raw(), pw, bias=pb, geglu=True, ln=x.fold(csum))
        pw, pb = self._packed()
        return ops.linear(x, pw, bias=pb, geglu=True)

    @tokens_boundary
    def forward(self, x):
        return self.run(x)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        if not glu:
            raise NotImplementedError("only the GEGLU feed-forward is on the ControlNet-SD1.5 path")
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), Linear(inner_dim, dim_out))

    def run(self, x, residual=None, stream=False):
        return self.net[2].run(self.net[0].run(x), residual=residual, stream=stream, twin=stream)

    @tokens_boundary
    def forward(self, x):
        return self.run(x)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.is_self = context_dim is None or context_dim == query_dim
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = Linear(query_dim, inner_dim, bias=False)
        self.to_k = Linear(context_dim, inner_dim, bias=False)
        self.to_v = Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self._cache = {}
        self.kv_static = None  # optional (context tensor, k, vt, nkv, ldv): hoisted K/V for a fixed context

    # -- the reference's fused weight attributes (attention.py:170,173), always derived from the live weights ----
    @property
    def qkv_w(self):
        return torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight]).transpose(0, 1).detach()

    @property
    def kv_w(self):
        return torch.cat([self.to_k.weight, self.to_v.weight]).transpose(0, 1).detach()

    def _packed(self, which):
        if which == "qkv":
            params = (self.to_q.weight, self.to_k.weight, self.to_v.weight)
        elif which == "kv":
            params = (self.to_k.weight, self.to_v.weight)
        else:
            params = (self.to_q.weight,)
        key = _param_key(*params)
        hit = self._cache.get(which)
        if hit is None or hit[0] != key:
            w = torch.cat([p.detach() for p in params], 0)  # rows = fused output features = (qkv_w).T
            self._cache[which] = (key, ops.pack_conv_weight(w))
            hit = self._cache[which]
        return hit[1]

    def _packed_ln(self, which, ln):
        params = (self.to_q.weight, self.to_k.weight, self.to_v.weight) if which == "qkv" else (self.to_q.weight,)
        key = _param_key(*params, ln.weight, ln.bias)
        hit = self._cache.get(which + "_ln")
        if hit is None or hit[0] != key:
            w = torch.cat([p.detach() for p in params], 0)
            self._cache[which + "_ln"] = (key,) + util.fold_ln_weight(w, None, ln)
            hit = self._cache[which + "_ln"]
        return hit[1:]

    def project_kv(self, context, k=None, vt=None):
        """k [B*heads, nkv, d], vt [B*heads, d, ldv] for a context [B, nkv, ctx_dim] (one GEMM on kv_w)."""
        b, nkv, _ = context.shape
        h, d = self.heads, self.dim_head
        ldv = (nkv + 7) // 8 * 8
        if k is None:
            k = torch.empty((b * h, nkv, d), dtype=BF16, device=context.device)
            vt = torch.empty((b * h, d, ldv), dtype=BF16, device=context.device)
        ops.qkv_project(context, self._packed("kv"), h, d, 1, k=k, vt=vt, ldv=ldv)
        return k, vt, nkv, ldv

    def run(self, x, context=None, residual=None, stream=False, feeds_ln=False):
        """x: [B, T, C] bf16 (or a DeferredLN: the LayerNorm is folded into the q / qkv projection); context
        [B, nkv, ctx_dim] bf16 or None (self-attention). Returns to_out(attn) (+ residual); `stream`: the result is an
        fp32 residual-stream tensor; feeds_ln: it feeds a LayerNorm (bf16 twin + row statistics from the epilogue)."""
        ln = x if isinstance(x, util.DeferredLN) else None
        if ln is not None:
            x = ln.raw()
        b, t, _ = x.shape
        h, d = self.heads, self.dim_head
        q = torch.empty((b * h, t, d), dtype=BF16, device=x.device)
        if context is None:
            ldv = (t + 7) // 8 * 8
            k = torch.empty_like(q)
            vt = torch.empty((b * h, d, ldv), dtype=BF16, device=x.device)
            if ln is not None:
                pw, bp, csum = self._packed_ln("qkv", ln.ln)
                ops.qkv_project(x, pw, h, d, 0, q=q, k=k, vt=vt, ldv=ldv, bias=bp, ln=ln.fold(csum))
            else:
                ops.qkv_project(x, self._packed("qkv"), h, d, 0, q=q, k=k, vt=vt, ldv=ldv)
            nkv = t
        else:
            if ln is not None:
                pw, bp, csum = self._packed_ln("q", ln.ln)
                ops.qkv_project(x, pw, h, d, 0, q=q, bias=bp, ln=ln.fold(csum))
            else:
                ops.qkv_project(x, self._packed("q"), h, d, 0, q=q)
            if self.kv_static is not None and self.kv_static[0] is context:
                _, k, vt, nkv, ldv = self.kv_static
            else:
                k, vt, nkv, ldv = self.project_kv(context)
        o = ops.attention(q, k, vt, b, h, t, nkv, d, ldv, self.scale)
        return self.to_out[0].run(o, residual=residual, stream=stream, twin=stream and feeds_ln,
                                  row_stats=stream and feeds_ln)

    def forward(self, x, context=None, mask=None):
        if exists(mask):
            raise NotImplementedError("attention masks are not on the ControlNet-SD1.5 path")
        if x.dtype == BF16:
            ctx = context.contiguous() if context is not None else None
            return self.run(x.contiguous(), ctx)
        ctx = ops.to_bf16(context.float()) if context is not None else None
        return ops.to_f32(self.run(ops.to_bf16(x.float()), ctx))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False):
        super().__init__()
        self.disable_self_attn = disable_self_attn
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                                    context_dim=context_dim if self.disable_self_attn else None)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1 = LayerNorm(dim)
        self.norm2 = LayerNorm(dim)
        self.norm3 = LayerNorm(dim)
        self.checkpoint = checkpoint

    def run(self, x, context=None):
        """attention.py:381-385; each residual add rides in the epilogue of the branch's last GEMM."""
        st = util.STREAM_FP32
        x = self.attn1.run(self.norm1.run(x), context if self.disable_self_attn else None, residual=x, stream=st,
                           feeds_ln=True)
        x = self.attn2.run(self.norm2.run(x), context, residual=x, stream=st, feeds_ln=True)
        x = self.ff.run(self.norm3.run(x), residual=x, stream=st)
        return x

    def forward(self, x, context=None):
        if x.dtype == BF16:
            return self.run(x.contiguous(), context)
        ctx = ops.to_bf16(context.float()) if context is not None else None
        y = self.run(ops.to_bf16(x.float()), ctx)
        return y if y.dtype == torch.float32 else ops.to_f32(y)


class SpatialTransformer(nn.Module):
    """GroupNorm -> 1x1 proj_in -> tokens -> transformer blocks -> 1x1 proj_out -> + x (attention.py:388-450).
    NHWC activations ARE the 'b (h w) c' token matrix, so both rearranges are free."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=True):
        super().__init__()
        if use_linear:
            raise NotImplementedError("use_linear SpatialTransformer is not on the ControlNet-SD1.5 path")
        if exists(context_dim) and not isinstance(context_dim, list):
            context_dim = [context_dim]
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                   disable_self_attn=disable_self_attn, checkpoint=use_checkpoint)
             for d in range(depth)])
        self.proj_out = zero_module(Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self.use_linear = use_linear

    def run(self, x, context=None):
        if not isinstance(context, list):
            context = [context]
        b, c, h, w = x.shape
        x_in = x
        st = util.STREAM_FP32
        t = self.proj_in.run(self.norm.run(x, silu=False, defer=True), out_fp32=st, stream=st, row_stats=st)  # the token stream
        tok = nhwc(t).reshape(b, h * w, t.shape[1])
        if getattr(t, "_sdeo_stream", False):  # [B,H,W,C] stream -> [B,T,C] stream (same memory), attributes carried over
            tok._sdeo_stream = True
            tok._twin = nhwc(t._twin).reshape(b, h * w, t.shape[1]) if t._twin is not None else None
            tok._gn_stats = None
            tok._row_stats = t._row_stats
        for i, block in enumerate(self.transformer_blocks):
            tok = block.run(tok, context[i])
        tok = operand(tok)                                                 # bf16 twin written by the last ff GEMM
        t = nchw_view(tok.reshape(b, h, w, tok.shape[-1]))
        return self.proj_out.run(t, residual=x_in, stream=st, gn_stats=True)  # feeds the next ResBlock's GroupNorm

    def forward(self, x, context=None):
        from .diffusionmodules.util import is_internal, to_external, to_internal
        if is_internal(x):
            return self.run(x, context)
        ctxs = context if isinstance(context, list) else [context]
        ctxs = [ops.to_bf16(c.float()) if (c is not None and c.dtype != BF16) else c for c in ctxs]
        return to_external(self.run(to_internal(x), ctxs))
