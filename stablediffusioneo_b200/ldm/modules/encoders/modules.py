"""FrozenCLIPEmbedder (ldm/modules/encoders/modules.py:89-144) on libsdeo.so: the CLIP ViT-L/14 text transformer that
turns a prompt's 77 token ids into the [B, 77, 768] context of every cross-attention (the first
"next" row after the denoising path). Parameter names follow transformers' CLIPTextModel, so the reference's
`cond_stage_model.transformer.*` checkpoint keys load unchanged.

Per layer: LayerNorm -> fused q|k|v projection (+bias) written head-major / V transposed -> causal flash attention ->
out_proj (+bias, + residual, fp32 stream) -> LayerNorm -> fc1 (+bias, quick_gelu in the epilogue) -> fc2 (+bias, +
residual). The tokenizer is host-side string processing and stays the reference's (transformers.CLIPTokenizer);
`forward` takes token ids."""
import torch
import torch.nn as nn

from .... import ops
from ..diffusionmodules.util import BF16, LayerNorm, Linear, _param_key


class CLIPAttention(nn.Module):
    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.k_proj = Linear(embed_dim, embed_dim)
        self.v_proj = Linear(embed_dim, embed_dim)
        self.q_proj = Linear(embed_dim, embed_dim)
        self.out_proj = Linear(embed_dim, embed_dim)
        self._cache = {}

    def _packed_qkv(self):
        ps = (self.q_proj, self.k_proj, self.v_proj)
        key = _param_key(*[p.weight for p in ps], *[p.bias for p in ps])
        hit = self._cache.get("qkv")
        if hit is None or hit[0] != key:
            w = torch.cat([p.weight.detach() for p in ps], 0)
            b = torch.cat([p.bias.detach() for p in ps], 0).float().contiguous()
            self._cache["qkv"] = (key, ops.pack_conv_weight(w), b)
            hit = self._cache["qkv"]
        return hit[1], hit[2]

    def run(self, x, residual):
        """x: bf16 [B, T, C] (LayerNorm output); residual: the fp32 stream. Causal self-attention + out_proj + residual."""
        b, t, _ = x.shape
        h, d = self.num_heads, self.head_dim
        ldv = (t + 7) // 8 * 8
        q = torch.empty((b * h, t, d), dtype=BF16, device=x.device)
        k = torch.empty_like(q)
        vt = torch.zeros((b * h, d, ldv), dtype=BF16, device=x.device)
        pw, bias = self._packed_qkv()
        ops.qkv_project(x, pw, h, d, 0, q=q, k=k, vt=vt, ldv=ldv, bias=bias)
        o = ops.attention(q, k, vt, b, h, t, t, d, ldv, self.scale, causal=True)
        return self.out_proj.run(o, residual=residual, stream=True)


class CLIPMLP(nn.Module):
    def __init__(self, hidden, intermediate):
        super().__init__()
        self.fc1 = Linear(hidden, intermediate)
        self.fc2 = Linear(intermediate, hidden)

    def run(self, x, residual):
        return self.fc2.run(self.fc1.run(x, act=ops.SDEO_ACT_QUICK_GELU), residual=residual, stream=True)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, hidden, heads, intermediate, eps):
        super().__init__()
        self.self_attn = CLIPAttention(hidden, heads)
        self.layer_norm1 = LayerNorm(hidden, eps=eps)
        self.mlp = CLIPMLP(hidden, intermediate)
        self.layer_norm2 = LayerNorm(hidden, eps=eps)

    def run(self, x):
        x = self.self_attn.run(ops.layernorm(x, self.layer_norm1.weight.detach(), self.layer_norm1.bias.detach(),
                                             self.layer_norm1.eps), residual=x)
        x = self.mlp.run(ops.layernorm(x, self.layer_norm2.weight.detach(), self.layer_norm2.bias.detach(),
                                       self.layer_norm2.eps), residual=x)
        return x


class CLIPEncoder(nn.Module):
    def __init__(self, hidden, heads, intermediate, layers, eps):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(hidden, heads, intermediate, eps) for _ in range(layers)])


class CLIPTextEmbeddings(nn.Module):
    def __init__(self, vocab, hidden, max_pos):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, hidden)
        self.position_embedding = nn.Embedding(max_pos, hidden)


class CLIPTextTransformer(nn.Module):
    def __init__(self, vocab, hidden, heads, intermediate, layers, max_pos, eps):
        super().__init__()
        self.embeddings = CLIPTextEmbeddings(vocab, hidden, max_pos)
        self.encoder = CLIPEncoder(hidden, heads, intermediate, layers, eps)
        self.final_layer_norm = LayerNorm(hidden, eps=eps)


class CLIPTextModel(nn.Module):
    """transformers.CLIPTextModel's parameter tree (`text_model.*`) for openai/clip-vit-large-patch14's text tower."""

    def __init__(self, vocab_size=49408, hidden_size=768, num_attention_heads=12, intermediate_size=3072,
                 num_hidden_layers=12, max_position_embeddings=77, layer_norm_eps=1e-5):
        super().__init__()
        self.text_model = CLIPTextTransformer(vocab_size, hidden_size, num_attention_heads, intermediate_size,
                                              num_hidden_layers, max_position_embeddings, layer_norm_eps)

    @torch.no_grad()
    def forward(self, input_ids):
        """input_ids int64 [B, T<=77] -> last_hidden_state fp32 [B, T, hidden] (after final_layer_norm)."""
        tm = self.text_model
        ids = input_ids.to(torch.int64).contiguous()
        x, _ = ops.embedding_add(ids, tm.embeddings.token_embedding.weight.detach().float().contiguous(),
                                 tm.embeddings.position_embedding.weight.detach().float().contiguous())
        for layer in tm.encoder.layers:
            x = layer.run(x)
        y = ops.layernorm(x, tm.final_layer_norm.weight.detach(), tm.final_layer_norm.bias.detach(), tm.final_layer_norm.eps)
        return ops.to_f32(y)


class FrozenCLIPEmbedder(nn.Module):
    """`cond_stage_model` of the reference: prompt -> [B, 77, 768] (layer="last"). Token ids in; a list of strings is
    tokenised with transformers.CLIPTokenizer if its vocabulary files are available locally."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, freeze=True, layer="last",
                 layer_idx=None):
        super().__init__()
        if layer != "last":
            raise NotImplementedError("only layer='last' is on the ControlNet-SD1.5 path (cldm_v15.yaml)")
        self.version, self.device, self.max_length = version, device, max_length
        self.transformer = CLIPTextModel(max_position_embeddings=max_length)
        self.tokenizer = None
        if freeze:
            self.freeze()

    def freeze(self):
        self.transformer = self.transformer.eval()
        for p in self.parameters():
            p.requires_grad = False

    def tokenize(self, text):
        if self.tokenizer is None:
            from transformers import CLIPTokenizer  # needs the vocabulary files on disk (no network here)
            self.tokenizer = CLIPTokenizer.from_pretrained(self.version)
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return enc["input_ids"]

    def forward(self, text):
        tokens = text if torch.is_tensor(text) else self.tokenize(text)
        return self.transformer(tokens.to(next(self.parameters()).device))

    def encode(self, text):
        return self(text)
