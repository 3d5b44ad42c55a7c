"""Torch-tensor front end of the C-ABI kernels. torch provides device memory and streams only; every op here
is one call into libsdeo.so on the current CUDA stream. Activations are bf16, physically NHWC: shape [N, H, W, C]
contiguous; token matrices are [rows, C] or [B, T, C] contiguous."""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ConvArgs, SDEO_ACT_NONE, SDEO_ACT_QUICK_GELU, SDEO_ACT_SILU, SDEO_EPI_GEGLU, SDEO_EPI_NORMAL, SDEO_EPI_QKV
from ._lib import check as _check

BF16 = torch.bfloat16

# Number of libsdeo kernels launched through this module (bench.py reports it as `gpu_launches`).
LAUNCHES = 0
_KERNELS_PER_CALL = {}  # GroupNorm is one cluster kernel on the denoiser tensors (two on the large VAE ones: undercounted)


def check(rc, what=""):
    global LAUNCHES
    _check(rc, what)
    LAUNCHES += _KERNELS_PER_CALL.get(what, 1)



def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.SdeoError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # ops launch on the CURRENT device's current stream; a tensor of another device would be dereferenced there
        raise _lib.SdeoError(f"{name}: tensor lives on {t.device} but the current CUDA device is "
                             f"cuda:{torch.cuda.current_device()} (wrap the call in torch.cuda.device(...))")
    if t.dtype != dtype:
        raise _lib.SdeoError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.SdeoError(f"{name}: expected a contiguous tensor")


class _Workspaces:
    """Per-(device, slot) scratch: split-K partial tiles + tile counters for conv2d, GroupNorm partial sums.
    Ops issued concurrently on different streams must use different slots (see `workspace_slot`)."""

    def __init__(self):
        self._conv = {}
        self._gn = {}
        self._retired = []
        self.slot = 0

    def conv(self, device):
        key = (device.index, self.slot)
        ws = self._conv.get(key)
        if ws is None:
            lib = _lib.load()
            nbytes = lib.sdeo_conv_workspace_bytes(None)
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)  # counters must start at zero
            self._conv[key] = ws
        return ws

    def gn(self, device, nbytes, kind="gn"):
        """kind: "gn" (engine GroupNorm, zero-initialised barrier words) / "gn_f16" (the plugin-contract kernels, which
        preset their slots to 0xFF) -- separate buffers. A workspace that has been handed out is never freed: captured CUDA
        graphs may hold its address; growing allocates a new buffer and keeps the old one alive."""
        key = (device.index, self.slot, kind)
        ws = self._gn.get(key)
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                self._retired.append(ws)
            ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)  # barrier words start at zero
            self._gn[key] = ws
        return ws


_workspaces = _Workspaces()


class workspace_slot:
    """Context manager selecting the scratch slot used by ops issued inside it (one slot per concurrent stream)."""

    def __init__(self, slot):
        self.slot = slot

    def __enter__(self):
        self.prev = _workspaces.slot
        _workspaces.slot = self.slot

    def __exit__(self, *exc):
        _workspaces.slot = self.prev


@dataclass
class PackedWeight:
    """bf16 K-major filter in the layout conv_gemm_kernel streams (see sdeo_pack_conv_weight)."""
    data: torch.Tensor
    cout: int
    ksize: int
    c1: int
    c2: int
    geglu_bn: int = 0


def pack_conv_weight(weight, c1=None, c2=0, geglu=False):
    """weight: fp32 [cout, cin, k, k] or [cout, cin] (Linear) on the GPU. c1/c2 split cin for fused concat inputs."""
    lib = _lib.load()
    if weight.dim() == 2:
        weight = weight[:, :, None, None]
    weight = weight.detach().to(torch.float32).contiguous()
    _req(weight, torch.float32, "weight")
    cout, cin, k, k2 = weight.shape
    assert k == k2 and k in (1, 2, 3)   # (2: a sub-pixel phase filter of upsample + conv3x3, see conv2d(up2_phase=...))
    if c1 is None:
        c1 = cin
    assert c1 + c2 == cin
    rows = lib.sdeo_packed_rows(cout)
    kp = lib.sdeo_packed_k(c1, c2, k)
    geglu_bn = 0
    if geglu:
        geglu_bn = lib.sdeo_pick_bn(rows, SDEO_EPI_GEGLU, 0)
        if geglu_bn <= 0:
            raise _lib.SdeoError(f"no GEGLU tile for {cout} rows")
    out = torch.empty((rows, kp), dtype=BF16, device=weight.device)
    check(lib.sdeo_pack_conv_weight(_ptr(weight), cout, c1, c2, k, geglu_bn, _ptr(out), _stream()), "pack_conv_weight")
    return PackedWeight(out, cout, k, c1, c2, geglu_bn)


def pack_geglu_bias(bias, geglu_bn):
    lib = _lib.load()
    bias = bias.detach().to(torch.float32).contiguous()
    out = torch.empty_like(bias)
    check(lib.sdeo_pack_geglu_bias(_ptr(bias), bias.numel(), geglu_bn, _ptr(out), _stream()), "pack_geglu_bias")
    return out


def conv2d(x, pw, x2=None, bias=None, emb=None, residual=None, scale=1.0, act=SDEO_ACT_NONE, stride=1, out=None,
           out_fp32=False, epi_mode=SDEO_EPI_NORMAL, qkv=None, twin=False, emb_step=None, gn_stats=False, row_stats=False,
           ln=None, pad_hi=0, gnf=None, up2_phase=0, out2=None):
    """x: [N,H,W,C1] bf16 (+ optional x2 [N,H,W,C2] = fused torch.cat along channels). Returns [N,Ho,Wo,cout].
    residual may be bf16 or fp32. out_fp32 + twin=True additionally writes a bf16 copy and returns (y_f32, y_bf16).
    emb: fp32 [N, cout] (row = sample), or with emb_step (int32 device scalar) a table [S, cout] whose row *emb_step is
    added to every sample. gn_stats (fp32 outputs): the epilogue also leaves per-channel partial statistics for the
    GroupNorm that consumes the result; they are attached to the returned fp32 tensor as `_gn_stats = (buffer, parts
    per sample)` when the kernel produced them (see groupnorm(stats=...)). row_stats (fp32 outputs feeding a LayerNorm):
    per-row partial statistics, attached as `_row_stats = (buffer, parts, rows)`. ln: LnFold -- this GEMM applies a
    LayerNorm to its input rows in the epilogue (x is the raw bf16 input; pw / bias carry gamma / beta).
    gnf: GnFold -- a GroupNorm (+ SiLU) applied to the RAW bf16 x (| x2) inside the operand path of this convolution
    (statistics from the producers' epilogues); gn_stats then also works with a bf16 output.
    up2_phase = 1 + 2a + b: pw is the 2x2 phase filter (upsample2x_conv_weights) of nearest-x2 upsampling + conv3x3; `out`
    is the full [N, 2H, 2W, cout] tensor (allocated when None), of which this call writes the pixels (2i + a, 2j + b).
    pad_hi=1 (3x3, stride 2): no leading padding, one trailing zero row / column = F.pad(x, (0,1,0,1)) + conv(padding=0),
    the VAE encoder's Downsample (model.py:78-84)."""
    lib = _lib.load()
    _req(x, BF16, "x")
    _req(x2, BF16, "x2")
    _req(bias, torch.float32, "bias")
    _req(emb, torch.float32, "emb")
    if residual is not None and residual.dtype != torch.float32:
        _req(residual, BF16, "residual")
    n, h, w, ld1 = x.shape
    c1 = pw.c1
    # the pixel stride may exceed the channels the filter consumes (zero-padded / ignored tail channels)
    assert ld1 >= c1 and (ld1 == c1 or ld1 - c1 < 8), f"conv2d: input has {ld1} channels, filter packed for {c1}"
    c2, ld2 = 0, 0
    if x2 is not None:
        assert x2.shape[:3] == x.shape[:3] and x2.shape[3] == pw.c2 and ld1 == c1
        c2 = ld2 = pw.c2
    else:
        assert pw.c2 == 0
    k = pw.ksize
    pad = 1 if k == 3 else 0
    if pad_hi:
        assert k == 3 and stride == 2 and pad_hi == 1
        pad = 0
    ho = (h + 2 * pad + pad_hi - k) // stride + 1
    wo = (w + 2 * pad + pad_hi - k) // stride + 1
    if up2_phase:
        assert k == 2 and stride == 1 and 1 <= up2_phase <= 4 and epi_mode == SDEO_EPI_NORMAL and residual is None
        ho, wo = 2 * h, 2 * w
    else:
        assert k != 2
    a = ConvArgs()
    a.up2_phase = up2_phase
    a.pad_hi = pad_hi
    a.x1, a.x2 = _ptr(x), _ptr(x2)
    a.n, a.h, a.w = n, h, w
    a.c1, a.ld1, a.c2, a.ld2 = c1, ld1, c2, ld2
    a.w_packed = _ptr(pw.data)
    a.cout, a.ksize, a.stride, a.pad = pw.cout, k, stride, pad
    a.epi_mode, a.act = epi_mode, act
    a.bias, a.emb = _ptr(bias), _ptr(emb)
    if emb_step is not None:
        _req(emb_step, torch.int32, "emb_step")
        a.emb_step = _ptr(emb_step)
    a.scale = float(scale)
    a.y_fp32 = 1 if out_fp32 else 0
    if epi_mode == SDEO_EPI_QKV:
        q, kk, vt, heads, dhead, tokens, ldv, first = qkv
        a.q, a.k, a.vt = _ptr(q), _ptr(kk), _ptr(vt)
        a.heads, a.dhead, a.tokens, a.ldv, a.qkv_first = heads, dhead, tokens, ldv, first
        out = None
    else:
        cols = pw.cout // 2 if epi_mode == SDEO_EPI_GEGLU else pw.cout
        if out is None:
            out = torch.empty((n, ho, wo, cols), dtype=torch.float32 if out_fp32 else BF16, device=x.device)
        else:
            assert out.shape[:3] == (n, ho, wo) and out.shape[3] >= cols and out.is_contiguous()
            assert out.dtype == (torch.float32 if out_fp32 else BF16)
        a.y, a.ldy = _ptr(out), out.shape[3]
        if residual is not None:
            assert residual.shape[:3] == (n, ho, wo) and residual.shape[3] >= cols and residual.is_contiguous()
            a.residual, a.ldr = _ptr(residual), residual.shape[3]
            a.residual_f32 = 1 if residual.dtype == torch.float32 else 0
    if not twin:
        assert out2 is None
    if twin:
        assert out_fp32 and epi_mode == SDEO_EPI_NORMAL
        if out2 is None:
            out2 = torch.empty(out.shape, dtype=BF16, device=x.device)
        else:
            assert out2.shape == out.shape and out2.dtype == BF16 and out2.is_contiguous()
        a.y2, a.ldy2 = _ptr(out2), out2.shape[3]
    ws = _workspaces.conv(x.device)
    a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
    if ln is not None:
        a.ln_stats, a.ln_parts, a.ln_ld = _ptr(ln.stats), ln.parts, ln.rows
        a.ln_c, a.ln_eps, a.ln_csum = ln.c, float(ln.eps), _ptr(ln.csum)
        assert ln.rows == n * ho * wo and ln.c == c1 and c2 == 0
    if gnf is not None:
        assert ln is None and gnf.gamma.numel() == c1 + c2 and (x2 is None) == (gnf.stats2 is None)
        _req(gnf.gamma, torch.float32, "gnf.gamma")
        _req(gnf.beta, torch.float32, "gnf.beta")
        a.gnf_stats1, a.gnf_parts1 = _ptr(gnf.stats1[0]), gnf.stats1[1]
        if x2 is not None:
            a.gnf_stats2, a.gnf_parts2 = _ptr(gnf.stats2[0]), gnf.stats2[1]
        a.gnf_gamma, a.gnf_beta = _ptr(gnf.gamma), _ptr(gnf.beta)
        a.gnf_groups, a.gnf_eps, a.gnf_silu = gnf.groups, float(gnf.eps), 1 if gnf.silu else 0
    rows_buf = None
    if row_stats and out_fp32 and epi_mode == SDEO_EPI_NORMAL:
        mx = ctypes.c_int32(0)
        _check(lib.sdeo_conv_row_stats_parts(ctypes.byref(a), ctypes.byref(mx), None), "conv_row_stats_parts")
        rows_buf = torch.empty((mx.value, n * ho * wo, 2), dtype=torch.float32, device=x.device)
        a.row_stats, a.row_stats_ld = _ptr(rows_buf), n * ho * wo
        gn_stats = False
    stats_buf = None
    if gn_stats and epi_mode == SDEO_EPI_NORMAL and (out_fp32 or out.shape[3] % 8 == 0):
        mx = ctypes.c_int32(0)
        _check(lib.sdeo_conv_gn_stats_slots(ctypes.byref(a), ctypes.byref(mx), None), "conv_gn_stats_slots")
        stats_buf = torch.empty((mx.value, pw.cout, 2), dtype=torch.float32, device=x.device)
        a.gn_stats = _ptr(stats_buf)
    check(lib.sdeo_conv2d(ctypes.byref(a), _stream()), "conv2d")
    if stats_buf is not None:
        parts = ctypes.c_int32(0)
        _check(lib.sdeo_conv_gn_stats_slots(ctypes.byref(a), None, ctypes.byref(parts)), "conv_gn_stats_slots")
        if parts.value > 0:
            out._gn_stats = (stats_buf, parts.value)
    if rows_buf is not None:
        parts = ctypes.c_int32(0)
        _check(lib.sdeo_conv_row_stats_parts(ctypes.byref(a), None, ctypes.byref(parts)), "conv_row_stats_parts")
        if parts.value > 0:
            out._row_stats = (rows_buf, parts.value, n * ho * wo)
    return (out, out2) if twin else out


GN_FOLD_MAX_PARTS = 48   # more partial slots per sample than this are folded by one small kernel first


def gn_stats_fold(stats, n, c):
    """(buffer, parts) of GroupNorm partial statistics -> the same with parts <= GN_FOLD_MAX_PARTS (sdeo_gn_stats_fold)."""
    lib = _lib.load()
    buf, parts = stats
    while parts > GN_FOLD_MAX_PARTS:
        op = ctypes.c_int32(0)
        _check(lib.sdeo_gn_stats_fold(None, None, n, parts, c, ctypes.byref(op), None), "gn_stats_fold")
        out = torch.empty((n * op.value, c, 2), dtype=torch.float32, device=buf.device)
        check(lib.sdeo_gn_stats_fold(_ptr(buf), _ptr(out), n, parts, c, ctypes.byref(op), _stream()), "gn_stats_fold")
        buf, parts = out, op.value
    return buf, parts


def upsample2x_conv_weights(weight):
    """3x3 filter [cout, cin, 3, 3] of a conv that follows nearest-x2 upsampling -> the four packed 2x2 phase filters
    (index 2a + b): along each axis phase 0 sees input pixels (i-1, i) with taps (w0, w1 + w2), phase 1 sees (i, i+1)
    with (w0 + w1, w2). Sums in fp32, one bf16 rounding when packed."""
    w = weight.detach().float()
    rows = (torch.stack([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 2), torch.stack([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], 2))
    out = []
    for a_ in range(2):
        r = rows[a_]                                     # [cout, cin, 2, 3]
        cols = (torch.stack([r[..., 0], r[..., 1] + r[..., 2]], 3), torch.stack([r[..., 0] + r[..., 1], r[..., 2]], 3))
        for b_ in range(2):
            out.append(pack_conv_weight(cols[b_].contiguous()))
    return out


def upsample2x_conv(x, phase_weights, bias=None, out=None):
    """conv3x3(nearest_x2(x)) as four sub-pixel phase convolutions over the low-resolution x [N,H,W,C] -> [N,2H,2W,cout]."""
    for ph in range(4):
        out = conv2d(x, phase_weights[ph], bias=bias, out=out, up2_phase=ph + 1)
    return out


@dataclass
class GnFold:
    """A GroupNorm (+ SiLU) folded into the convolution that consumes it: the producers' partial statistics
    ((buffer, parts per sample) for x and x2), affine parameters of the virtual concat, group count, eps."""
    stats1: tuple
    stats2: object
    gamma: torch.Tensor
    beta: torch.Tensor
    groups: int
    eps: float
    silu: bool


@dataclass
class LnFold:
    """A LayerNorm folded into the GEMM that consumes it: row statistics of the raw input (from its producer's
    epilogue), the normalised width, eps, and the column sums of the gamma-scaled packed weight."""
    stats: torch.Tensor
    parts: int
    rows: int
    c: int
    eps: float
    csum: torch.Tensor


def linear(x, pw, bias=None, residual=None, act=SDEO_ACT_NONE, out_fp32=False, geglu=False, twin=False, row_stats=False,
           ln=None):
    """x: [..., K] bf16 -> [..., cout] (GEGLU: [..., cout/2]). Runs the 1x1 case of the implicit-GEMM kernel."""
    lead = x.shape[:-1]
    rows = 1
    for s in lead:
        rows *= s
    x4 = x.reshape(1, 1, rows, x.shape[-1])
    res4 = residual.reshape(1, 1, rows, residual.shape[-1]) if residual is not None else None
    y = conv2d(x4, pw, bias=bias, residual=res4, act=act, out_fp32=out_fp32,
               epi_mode=SDEO_EPI_GEGLU if geglu else SDEO_EPI_NORMAL, twin=twin, row_stats=row_stats, ln=ln)
    y0 = y[0] if twin else y
    r0 = y0.reshape(*lead, y0.shape[-1])
    if getattr(y0, "_row_stats", None) is not None:
        r0._row_stats = y0._row_stats
    if twin:
        return r0, y[1].reshape(*lead, y[1].shape[-1])
    return r0


def qkv_project(x, pw, heads, dhead, first, q=None, k=None, vt=None, ldv=None, bias=None, ln=None):
    """x: [B, T, K] bf16. Packed rows hold consecutive blocks of heads*dhead columns for q/k/v starting at
    `first` (0=q, 1=k, 2=v). Writes q,k as [B*heads, T, dhead] and v transposed as [B*heads, dhead, ldv]."""
    b, t, kdim = x.shape
    x4 = x.reshape(1, 1, b * t, kdim)
    conv2d(x4, pw, bias=bias, epi_mode=SDEO_EPI_QKV, qkv=(q, k, vt, heads, dhead, t, ldv or 0, first), ln=ln)


def groupnorm(x, gamma, beta, eps, silu, x2=None, groups=32, out=None, stats=None, stats2=None):
    """x: [N,H,W,C1] (+ x2 [N,H,W,C2]); returns the normalised concat [N,H,W,C1+C2]. stats / stats2: the `_gn_stats`
    (buffer, parts per sample) the producing convolutions attached to x / x2 -- the tensor is then read only once."""
    lib = _lib.load()
    f32 = x.dtype == torch.float32
    _req(x, torch.float32 if f32 else BF16, "x")
    _req(x2, torch.float32 if f32 else BF16, "x2")
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    n, h, w, c1 = x.shape
    c2 = x2.shape[3] if x2 is not None else 0
    if out is None:
        out = torch.empty((n, h, w, c1 + c2), dtype=BF16, device=x.device)
    if stats is not None and (x2 is None or stats2 is not None):
        # (large feature maps leave one slot per 128-pixel tile: thousands per sample -- one small kernel folds them first)
        stats = gn_stats_fold(stats, n, c1)
        s2 = gn_stats_fold(stats2, n, c2) if x2 is not None else (None, 0)
        check(lib.sdeo_groupnorm_apply_stats(_ptr(x), _ptr(x2), 1 if f32 else 0, _ptr(stats[0]), stats[1], _ptr(s2[0]), s2[1],
                                             _ptr(gamma), _ptr(beta), _ptr(out), n, h * w, c1, c2, groups, float(eps),
                                             1 if silu else 0, _stream()), "groupnorm_apply_stats")
        return out
    nbytes = lib.sdeo_groupnorm_workspace_bytes(n, h * w, groups)
    ws = _workspaces.gn(x.device, nbytes)
    check(lib.sdeo_groupnorm_nhwc(_ptr(x), _ptr(x2), 1 if f32 else 0, _ptr(gamma), _ptr(beta), _ptr(out), n, h * w, c1, c2, groups,
                                  float(eps), 1 if silu else 0, _ptr(ws), ws.numel(), _stream()), "groupnorm")
    return out


def groupnorm_f16(x, gamma, beta, eps=1e-5, silu=False, groups=32):
    """The TensorRT plugin's contract (groupNormPlugin.cpp:136-160): x fp16 NHWC [N,H,W,C] -> fp16, fp32 gamma / beta."""
    lib = _lib.load()
    _req(x, torch.float16, "x")
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    ws = _workspaces.gn(x.device, lib.sdeo_groupnorm_f16_workspace_bytes(n, h * w, c, groups), kind="gn_f16")
    global LAUNCHES
    check(lib.sdeo_groupnorm_nhwc_f16(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), n, h * w, c, groups, float(eps),
                                      1 if silu else 0, _ptr(ws), ws.numel(), _stream()), "groupnorm_f16")
    LAUNCHES += 1
    return out


def layernorm(x, gamma, beta, eps=1e-5):
    """x: bf16 or fp32 (residual stream) [..., C] -> bf16."""
    lib = _lib.load()
    f32 = x.dtype == torch.float32
    _req(x, torch.float32 if f32 else BF16, "x")
    c = x.shape[-1]
    rows = x.numel() // c
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(lib.sdeo_layernorm(_ptr(x), 1 if f32 else 0, _ptr(gamma), _ptr(beta), _ptr(out), rows, c, float(eps), _stream()),
          "layernorm")
    return out


def attention(q, k, vt, batch, heads, nq, nkv, d, ldv, scale, out=None, causal=False):
    lib = _lib.load()
    if out is None:
        out = torch.empty((batch, nq, heads * d), dtype=BF16, device=q.device)
    if causal:
        assert nq == nkv
        check(lib.sdeo_attention_causal(_ptr(q), _ptr(k), _ptr(vt), _ptr(out), batch, heads, nq, d, ldv, float(scale),
                                        _stream()), "attention_causal")
        return out
    check(lib.sdeo_attention(_ptr(q), _ptr(k), _ptr(vt), _ptr(out), batch, heads, nq, nkv, d, ldv, float(scale),
                             _stream()), "attention")
    return out


def embedding_add(ids, tok, pos):
    """ids int64 [B, T]; tok fp32 [V, C]; pos fp32 [T, C] -> (fp32 [B, T, C], bf16 twin)."""
    lib = _lib.load()
    _req(ids, torch.int64, "ids")
    _req(tok, torch.float32, "tok")
    _req(pos, torch.float32, "pos")
    b, t = ids.shape
    c = tok.shape[1]
    y = torch.empty((b, t, c), dtype=torch.float32, device=ids.device)
    y2 = torch.empty((b, t, c), dtype=BF16, device=ids.device)
    check(lib.sdeo_embedding_add(_ptr(ids), _ptr(tok), _ptr(pos), _ptr(y), _ptr(y2), b * t, t, c, tok.shape[0], _stream()),
          "embedding_add")
    return y, y2


def softmax_rows(x, scale, out=None):
    """fp32 scores [rows, cols] -> bf16 softmax(x * scale) rows."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    check(lib.sdeo_softmax_rows(_ptr(x), _ptr(out), rows, cols, cols, cols, float(scale), _stream()), "softmax_rows")
    return out


def nchw_to_nhwc(x, ldy=None, scale=1.0):
    """fp32 [N,C,H,W] -> bf16 [N,H,W,ldy] = scale * x (channels zero-padded to ldy, default round_up(C, 8))."""
    lib = _lib.load()
    x = x.contiguous()
    _req(x, torch.float32, "x")
    n, c, h, w = x.shape
    if ldy is None:
        ldy = (c + 7) // 8 * 8
    out = torch.empty((n, h, w, ldy), dtype=BF16, device=x.device)
    check(lib.sdeo_nchw_to_nhwc_bf16(_ptr(x), _ptr(out), n, c, h * w, ldy, float(scale), _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, c=None):
    """bf16 or fp32 [N,H,W,ld] -> fp32 [N,c,H,W]."""
    lib = _lib.load()
    assert x.is_contiguous()
    n, h, w, ld = x.shape
    if c is None:
        c = ld
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    if x.dtype == BF16:
        check(lib.sdeo_nhwc_bf16_to_nchw(_ptr(x), _ptr(out), n, c, h * w, ld, _stream()), "nhwc_to_nchw")
    else:
        _req(x, torch.float32, "x")
        check(lib.sdeo_nhwc_f32_to_nchw(_ptr(x), _ptr(out), n, c, h * w, ld, _stream()), "nhwc_f32_to_nchw")
    return out


def upsample_nearest2x(x):
    lib = _lib.load()
    _req(x, BF16, "x")
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=BF16, device=x.device)
    check(lib.sdeo_upsample_nearest2x(_ptr(x), _ptr(out), n, h, w, c, _stream()), "upsample2x")
    return out


def add_scaled(a, b, alpha=1.0, out=None):
    lib = _lib.load()
    _req(a, BF16, "a")
    _req(b, BF16, "b")
    assert a.shape == b.shape
    if out is None:
        out = torch.empty_like(a)
    check(lib.sdeo_add_scaled(_ptr(a), _ptr(b), float(alpha), _ptr(out), a.numel(), _stream()), "add_scaled")
    return out


def timestep_embedding(t, n, dim, step_idx=None, max_period=10000.0):
    """t: int64 device tensor ([n], or a per-step table when step_idx is given) -> bf16 [n, dim], [cos | sin]."""
    lib = _lib.load()
    _req(t, torch.int64, "t")
    out = torch.empty((n, dim), dtype=BF16, device=t.device)
    check(lib.sdeo_timestep_embedding(_ptr(t), _ptr(step_idx), _ptr(out), n, dim, dim, float(max_period), _stream()),
          "timestep_embedding")
    return out


def silu(x):
    lib = _lib.load()
    _req(x, BF16, "x")
    out = torch.empty_like(x)
    check(lib.sdeo_silu(_ptr(x), _ptr(out), x.numel(), _stream()), "silu")
    return out


def to_bf16(x):
    lib = _lib.load()
    x = x.contiguous()
    _req(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(lib.sdeo_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "f32_to_bf16")
    return out


def to_f32(x):
    lib = _lib.load()
    _req(x, BF16, "x")
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(lib.sdeo_bf16_to_f32(_ptr(x), _ptr(out), x.numel(), _stream()), "bf16_to_f32")
    return out


def cfg_ddim_step(eps_c, eps_u, x, coef_table, step_idx=None, noise=None, x_prev=None, pred_x0=None, x_next=None,
                  dup=0, eps_nhwc=False, noise_table=None):
    """x: fp32 [N,C,H,W]. eps_*: fp32 NCHW, or NHWC [N,H,W,ld] when eps_nhwc. Returns (x_prev, pred_x0).
    noise_table: fp32 [S, N, C, H, W] of which row *step_idx is added (times sigma) instead of `noise`."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    n, c, h, w = x.shape
    ld_eps = eps_c.shape[-1] if eps_nhwc else 0
    if x_prev is None:
        x_prev = torch.empty_like(x)
    ldn = x_next.shape[-1] if x_next is not None else 0
    if noise_table is not None:
        assert noise is None and step_idx is not None and tuple(noise_table.shape[1:]) == (n, c, h, w)
        _req(noise_table, torch.float32, "noise_table")
        check(lib.sdeo_cfg_ddim_step_noise_table(_ptr(eps_c), _ptr(eps_u), 1 if eps_nhwc else 0, ld_eps, _ptr(x),
                                                 _ptr(noise_table), _ptr(x_prev), _ptr(pred_x0), _ptr(x_next), dup, ldn,
                                                 _ptr(coef_table), _ptr(step_idx), n, c, h * w, _stream()),
              "cfg_ddim_step_noise_table")
        return x_prev, pred_x0
    check(lib.sdeo_cfg_ddim_step(_ptr(eps_c), _ptr(eps_u), 1 if eps_nhwc else 0, ld_eps, _ptr(x), _ptr(noise),
                                 _ptr(x_prev), _ptr(pred_x0), _ptr(x_next), dup, ldn, _ptr(coef_table),
                                 _ptr(step_idx), n, c, h * w, _stream()), "cfg_ddim_step")
    return x_prev, pred_x0


def axpby(x, z, a, b, out=None):
    """fp32 a[s] * x + b[s] * z with per-sample coefficients: a, b python floats or fp32 device tensors [N]."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(z, torch.float32, "z")
    assert x.shape == z.shape
    n = x.shape[0]
    coef = lambda c: (c.to(device=x.device, dtype=torch.float32).reshape(n).contiguous() if torch.is_tensor(c)
                      else torch.full((n,), float(c), dtype=torch.float32, device=x.device))
    a, b = coef(a), coef(b)
    if out is None:
        out = torch.empty_like(x)
    check(lib.sdeo_axpby_f32(_ptr(x), _ptr(z), _ptr(a), _ptr(b), _ptr(out), x.numel(), x.numel() // n, _stream()), "axpby")
    return out


def mask_blend(x0, noise, img, mask, a, b):
    """mask * (a[s] * x0 + b[s] * noise) + (1 - mask) * img on fp32 [N, C, H, W]; mask [N, 1 or C, H, W]."""
    lib = _lib.load()
    for t, name in ((x0, "x0"), (noise, "noise"), (img, "img"), (mask, "mask")):
        _req(t, torch.float32, name)
    n, c, h, w = img.shape
    assert x0.shape == img.shape == noise.shape and mask.shape[0] == n and mask.shape[2:] == img.shape[2:]
    a = a.to(device=img.device, dtype=torch.float32).reshape(n).contiguous()
    b = b.to(device=img.device, dtype=torch.float32).reshape(n).contiguous()
    out = torch.empty_like(img)
    check(lib.sdeo_mask_blend_f32(_ptr(x0), _ptr(noise), _ptr(img), _ptr(mask), _ptr(a), _ptr(b), _ptr(out), n, c,
                                  mask.shape[1], h * w, _stream()), "mask_blend")
    return out


def mask_blend_table_(img, orig_table, mask, step_idx):
    """In place: img = mask * orig_table[*step_idx] + (1 - mask) * img (fp32 [N, C, H, W]; orig_table [S, N, C, H, W]; mask
    [N, 1 or C, H, W]; step_idx int32 device scalar) -- the inpainting blend as a node of the captured step graph."""
    lib = _lib.load()
    for t, name in ((orig_table, "orig_table"), (img, "img"), (mask, "mask")):
        _req(t, torch.float32, name)
    _req(step_idx, torch.int32, "step_idx")
    n, c, h, w = img.shape
    assert orig_table.shape[1:] == img.shape and mask.shape[0] == n and mask.shape[2:] == img.shape[2:]
    check(lib.sdeo_mask_blend_table_f32(_ptr(orig_table), _ptr(img), _ptr(mask), _ptr(img), _ptr(step_idx), n, c,
                                        mask.shape[1], h * w, _stream()), "mask_blend_table")
    return img


def set_autotune(enable=True):
    """Per-shape autotuning of the conv/linear kernel's N tile and K slices (see sdeo_conv_autotune)."""
    _check(_lib.load().sdeo_conv_autotune(1 if enable else 0), "conv_autotune")


class cta_budget:
    """Context manager: convolutions / linears enqueued inside use at most `n` CTAs each (see sdeo_conv_set_cta_budget)."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        _check(_lib.load().sdeo_conv_set_cta_budget(self.n), "conv_set_cta_budget")

    def __exit__(self, *exc):
        _check(_lib.load().sdeo_conv_set_cta_budget(0), "conv_set_cta_budget")


def memset(t, value=0):
    lib = _lib.load()
    assert t.is_contiguous()
    check(lib.sdeo_memset_async(_ptr(t), int(value), t.numel() * t.element_size(), _stream()), "memset")
    return t


def counter_add(ctr, delta=1):
    lib = _lib.load()
    check(lib.sdeo_counter_add(_ptr(ctr), int(delta), _stream()), "counter_add")


def image_to_u8(x, c=3):
    lib = _lib.load()
    _req(x, BF16, "x")
    n, h, w, ld = x.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
    check(lib.sdeo_image_to_u8(_ptr(x), _ptr(out), n * h * w, c, ld, _stream()), "image_to_u8")
    return out


# ---- fp32 ("precise") mode: split-term operands for the bf16 GEMM + plain fp32 passes (csrc/precise.cu) -------------
# term patterns, 2 bits per K block (0 = hi, 1 = mid, 2 = lo): activation side / weight side. The GEMM over the
# concatenated blocks sums a_level * w_level per block.
SPLIT_PATTERNS = {
    3: ((0, 1, 0), (0, 0, 1)),                       # hi.hi + mid.hi + hi.mid
    6: ((0, 1, 0, 1, 2, 0), (0, 0, 1, 1, 0, 2)),     # + mid.mid + lo.hi + hi.lo
}


def _pattern_bits(levels):
    bits = 0
    for i, lv in enumerate(levels):
        bits |= (lv & 3) << (2 * i)
    return bits


def split_terms(x, terms=3, nchw=False):
    """fp32 [..., C] (or NCHW [N, C, H, W] when nchw) -> bf16 [..., terms * Cp] activation-side split (Cp = C rounded up
    to 8; for nchw the result is NHWC [N, H, W, terms * Cp])."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    pat = _pattern_bits(SPLIT_PATTERNS[terms][0])
    if nchw:
        n, c, h, w = x.shape
        cp = (c + 7) // 8 * 8
        out = torch.empty((n, h, w, terms * cp), dtype=BF16, device=x.device)
        check(lib.sdeo_split_terms(_ptr(x), _ptr(out), n * h * w, c, cp, 0, h * w, terms, pat, _stream()), "split_terms")
        return out
    c = x.shape[-1]
    cp = (c + 7) // 8 * 8
    rows = x.numel() // c
    out = torch.empty(tuple(x.shape[:-1]) + (terms * cp,), dtype=BF16, device=x.device)
    check(lib.sdeo_split_terms(_ptr(x), _ptr(out), rows, c, cp, c, 0, terms, pat, _stream()), "split_terms")
    return out


def pack_conv_weight_split(weight, terms=3, c1=None, c2=0):
    """fp32 filter [cout, cin, k, k] / [cout, cin] -> PackedWeight over the weight-side split channels
    [terms * Cp1 (+ terms * Cp2)] matching split_terms() of the input(s)."""
    lib = _lib.load()
    if weight.dim() == 2:
        weight = weight[:, :, None, None]
    weight = weight.detach().to(torch.float32).contiguous()
    _req(weight, torch.float32, "weight")
    cout, cin, k, _ = weight.shape
    if c1 is None:
        c1 = cin
    assert c1 + c2 == cin
    pat = _pattern_bits(SPLIT_PATTERNS[terms][1])
    parts, widths = [], []
    for c0, cc in ((0, c1), (c1, c2)):
        if cc == 0:
            continue
        cp = (cc + 7) // 8 * 8
        buf = torch.empty((cout, terms * cp, k, k), dtype=torch.float32, device=weight.device)
        check(lib.sdeo_split_terms_weight(_ptr(weight), _ptr(buf), cout, cin, c0, cc, cp, k * k, terms, pat, _stream()),
              "split_terms_weight")
        parts.append(buf)
        widths.append(terms * cp)
    w3 = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
    return pack_conv_weight(w3, c1=widths[0], c2=widths[1] if len(widths) > 1 else 0)


def groupnorm_f32(x, gamma, beta, eps, silu, x2=None, groups=32):
    """fp32 NHWC [N,H,W,C1] (+ x2) -> fp32 [N,H,W,C1+C2]."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(x2, torch.float32, "x2")
    n, h, w, c1 = x.shape
    c2 = x2.shape[3] if x2 is not None else 0
    out = torch.empty((n, h, w, c1 + c2), dtype=torch.float32, device=x.device)
    check(lib.sdeo_groupnorm_f32(_ptr(x), _ptr(x2), _ptr(gamma), _ptr(beta), _ptr(out), n, h * w, c1, c2, groups, float(eps),
                                 1 if silu else 0, _stream()), "groupnorm_f32")
    return out


def layernorm_f32(x, gamma, beta, eps=1e-5):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    c = x.shape[-1]
    out = torch.empty_like(x)
    check(lib.sdeo_layernorm_f32(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), x.numel() // c, c, float(eps), _stream()),
          "layernorm_f32")
    return out


def attention_f32(q, k, v, heads, scale):
    """q [B, Nq, heads*d], k / v [B, Nkv, heads*d] fp32 -> [B, Nq, heads*d] fp32."""
    lib = _lib.load()
    for t, name in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.float32, name)
    b, nq, c = q.shape
    nkv = k.shape[1]
    d = c // heads
    out = torch.empty_like(q)
    check(lib.sdeo_attention_f32(_ptr(q), _ptr(k), _ptr(v), _ptr(out), b, heads, nq, nkv, d, c, c, c, c, float(scale),
                                 _stream()), "attention_f32")
    return out


def geglu_f32(x):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    inner = x.shape[-1] // 2
    out = torch.empty(tuple(x.shape[:-1]) + (inner,), dtype=torch.float32, device=x.device)
    check(lib.sdeo_geglu_f32(_ptr(x), _ptr(out), x.numel() // (2 * inner), inner, _stream()), "geglu_f32")
    return out


def silu_f32(x):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    out = torch.empty_like(x)
    check(lib.sdeo_silu_f32(_ptr(x), _ptr(out), x.numel(), _stream()), "silu_f32")
    return out


def timestep_embedding_f32(t, dim, max_period=10000.0):
    lib = _lib.load()
    _req(t, torch.int64, "t")
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    check(lib.sdeo_timestep_embedding_f32(_ptr(t), _ptr(out), t.shape[0], dim, float(max_period), _stream()),
          "timestep_embedding_f32")
    return out


# ---- hint preprocessing (csrc/canny.cu) -------------------------------------------------------------------------
def canny(img, low_threshold, high_threshold):
    """uint8 CUDA tensor [H, W] or [H, W, C] (C = 1 or 3) -> uint8 [H, W] edge map (0 / 255), bit-exact with cv2.Canny."""
    lib = _lib.load()
    _req(img, torch.uint8, "img")
    if img.dim() == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    edges = torch.empty((h, w), dtype=torch.uint8, device=img.device)
    ws = torch.empty(lib.sdeo_canny_workspace_bytes(h, w), dtype=torch.uint8, device=img.device)
    global LAUNCHES
    check(lib.sdeo_canny_u8(_ptr(img), h, w, c, float(low_threshold), float(high_threshold), _ptr(edges), _ptr(ws),
                            ws.numel(), _stream()), "canny")
    LAUNCHES += 2
    return edges


def edges_to_hint(edges, num_samples=1):
    """uint8 [H, W] edge map -> fp32 [num_samples, 3, H, W] = HWC3(map) / 255 (canny2image_torch.py:34-38)."""
    lib = _lib.load()
    _req(edges, torch.uint8, "edges")
    h, w = edges.shape
    hint = torch.empty((num_samples, 3, h, w), dtype=torch.float32, device=edges.device)
    check(lib.sdeo_edges_to_hint(_ptr(edges), _ptr(hint), num_samples, h * w, _stream()), "edges_to_hint")
    return hint
