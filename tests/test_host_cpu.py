"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/sdeo.h declares, the host-side
logic (schedules, module surface / state-dict names, planner queries, error paths) and the world_size-2 gloo path."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import O, ROOT, load_golden


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sdeo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdeo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdeo.h but not exported by libsdeo.so"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert lib.sdeo_version() >= 1


def test_conv_args_struct_matches_header():
    """Field order of the ctypes mirror == field order of struct sdeo_conv_args."""
    from stablediffusioneo_b200 import _lib
    src = open(os.path.join(ROOT, "include", "sdeo.h")).read()
    body = src[src.index("typedef struct sdeo_conv_args {"):src.index("} sdeo_conv_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if " " in decl else decl
        names = names.replace("void*", "").replace("float*", "").replace("const", "")
        for n in names.split(","):
            n = n.strip().lstrip("*").strip()
            n = n.split()[-1].lstrip("*") if n else n
            if n:
                fields.append(n)
    assert fields == [f[0] for f in _lib.ConvArgs._fields_]


def test_planner_queries_without_gpu():
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    assert lib.sdeo_packed_rows(320) == 320 and lib.sdeo_packed_rows(4) == 16 and lib.sdeo_packed_rows(3) == 16
    assert lib.sdeo_packed_k(320, 0, 3) == 9 * 320
    assert lib.sdeo_packed_k(1280, 640, 3) == 9 * 1920
    assert lib.sdeo_packed_k(4, 0, 3) == 9 * 64          # ragged channels pad to one 64-wide K chunk per tap
    assert lib.sdeo_packed_k(96, 0, 3) == 9 * 128
    for rows in (320, 640, 1280, 960, 1920, 3840, 2560, 5120, 10240, 512, 256, 128, 96, 32, 16):
        bn = lib.sdeo_pick_bn(rows, _lib.SDEO_EPI_NORMAL, 0)
        assert 16 <= bn <= 256 and bn % 16 == 0 and rows % bn == 0
    for rows in (2560, 5120, 10240, 512):
        bn = lib.sdeo_pick_bn(rows, _lib.SDEO_EPI_GEGLU, 0)
        assert bn % 32 == 0 and rows % bn == 0 and (rows // 2) % (bn // 2) == 0
    assert lib.sdeo_conv_workspace_bytes(None) >= 148 * 128 * 260 * 4  # one fp32 partial tile per resident CTA
    assert lib.sdeo_conv_counter_bytes() == 0                          # no tile counters: the K slices form a cluster
    assert lib.sdeo_groupnorm_workspace_bytes(2, 1536, 32) > 0


def test_errors_are_codes_not_crashes():
    """Bad arguments return SDEO_EINVAL with a message (the plugin contract: enqueue returns -1, never throws,
    groupNormPlugin.cpp:223-227)."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    a = _lib.ConvArgs()
    assert lib.sdeo_conv2d(ctypes.byref(a), None) == -22
    assert b"null" in lib.sdeo_last_error()
    assert lib.sdeo_groupnorm_nhwc(None, None, 0, None, None, None, 1, 1, 8, 0, 32, 1e-5, 0, None, 0, None) == -22
    assert lib.sdeo_attention(None, None, None, None, 1, 1, 1, 1, 8, 8, 1.0, None) == -22
    assert lib.sdeo_layernorm(None, 0, None, None, None, 1, 8, 1e-5, None) == -22


def test_ops_refuse_cpu_tensors():
    """There is no CPU fallback: ops raise on host tensors instead of computing on the CPU."""
    from stablediffusioneo_b200 import _lib, ops
    x = torch.zeros((1, 4, 4, 8), dtype=torch.bfloat16)
    with pytest.raises(_lib.SdeoError):
        ops.groupnorm(x, torch.ones(8), torch.zeros(8), 1e-5, True, groups=1)


def test_module_surface_and_state_dict_names():
    """Our ControlLDM has exactly the reference's parameter names/shapes (as restated by the oracle's spec, which
    make_golden.py checks against the real modules with strict load_state_dict)."""
    from stablediffusioneo_b200.cldm.cldm import ControlLDM, ControlNet, ControlledUnetModel
    from stablediffusioneo_b200.ldm.modules.attention import CrossAttention
    with torch.device("meta"):
        model = ControlLDM()
    names = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    spec = {}
    spec.update({"model.diffusion_model." + n: s for n, s, _ in O.unet_param_spec(O.SD15)})
    spec.update({"control_model." + n: s for n, s, _ in O.controlnet_param_spec(O.SD15)})
    spec.update({"first_stage_model." + n: s for n, s, _ in O.vae_param_spec(O.SD15_VAE)})
    assert names == spec
    assert isinstance(model.model.diffusion_model, ControlledUnetModel) and isinstance(model.control_model, ControlNet)
    assert model.control_scales == [1.0] * 13 and model.num_timesteps == 1000 and model.parameterization == "eps"
    att = CrossAttention(query_dim=64, heads=8, dim_head=8)
    assert att.qkv_w.shape == (64, 192)
    assert torch.equal(att.qkv_w[:, 64:128], att.to_k.weight.t())
    att2 = CrossAttention(query_dim=64, context_dim=96, heads=8, dim_head=8)
    assert att2.kv_w.shape == (96, 128)
    with pytest.raises(NotImplementedError):
        ControlNet(image_size=32, in_channels=4, model_channels=64, hint_channels=3, num_res_blocks=2,
                   attention_resolutions=[1], num_heads=8, use_spatial_transformer=True, context_dim=96,
                   use_scale_shift_norm=True)


def test_sampler_schedule_matches_reference():
    """DDIMSampler.make_schedule on a duck-typed CPU model object reproduces the reference's tables."""
    from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler
    g = load_golden("sd15_256x384")

    class M:
        num_timesteps = 1000
        device = torch.device("cpu")
        ac = np.cumprod(1.0 - O.make_beta_schedule(), axis=0)
        betas = torch.tensor(O.make_beta_schedule(), dtype=torch.float32)
        alphas_cumprod = torch.tensor(ac, dtype=torch.float32)
        alphas_cumprod_prev = torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32)

    s = DDIMSampler(M())
    s.make_schedule(20, ddim_eta=0.0, verbose=False)
    assert np.array_equal(np.asarray(s.ddim_timesteps), g["ddim_timesteps"].numpy())
    assert np.allclose(np.asarray(s.ddim_alphas, dtype=np.float64), g["ddim_alphas"].numpy(), rtol=1e-6)
    assert np.allclose(np.asarray(s.ddim_alphas_prev, dtype=np.float64), g["ddim_alphas_prev"].numpy(), rtol=1e-6)
    row = s._coef_row(19, 9.0)
    assert row[0] == 9.0 and abs(row[2] - 1.0 / np.sqrt(float(g["ddim_alphas"][19]))) < 1e-6 and row[5] == 0.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dist_worker(rank, world, port, n_items, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from stablediffusioneo_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    idx = parallel.shard_indices(n_items, rank, world)
    local = torch.stack([torch.full((4, 2, 3), float(i)) for i in idx]) if idx else torch.zeros((0, 4, 2, 3))
    full = parallel.gather_by_image(local, n_items)
    ok = all(bool((full[i] == float(i)).all()) for i in range(n_items)) and full.shape == (n_items, 4, 2, 3)
    out[rank] = int(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5, 1])
def test_replica_sharding_and_gather_gloo(n_items):
    """world_size-2 gloo run of the N>1 path: images shard r::world, the final latents are all-gathered in order."""
    from stablediffusioneo_b200 import parallel
    assert parallel.shard_indices(8, 1, 2) == [1, 3, 5, 7] and parallel.shard_indices(5, 1, 2) == [1, 3]
    assert sorted(parallel.shard_indices(5, 0, 2) + parallel.shard_indices(5, 1, 2)) == list(range(5))
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0, 0])
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, n_items, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def test_checkpoint_helpers_and_create_model(tmp_path):
    """cldm/model.py:8-28 mirrors: load_state_dict for .pth ('state_dict'-wrapped) and .safetensors, create_model from a
    cldm_v15.yaml-style config (PyYAML instead of OmegaConf); strict loading into the module tree proves the checkpoint
    prefixes (`control_model.`, `model.diffusion_model.`, `first_stage_model.`) are the module names."""
    import yaml
    from helpers import O, unet_kwargs, vae_kwargs
    from stablediffusioneo_b200.cldm.model import create_model, load_state_dict
    cfg = {"model": {"target": "cldm.cldm.ControlLDM", "params": {
        "linear_start": 0.00085, "linear_end": 0.0120, "timesteps": 1000, "control_key": "hint", "scale_factor": 0.18215,
        "only_mid_control": False,
        "control_stage_config": {"target": "cldm.cldm.ControlNet",
                                 "params": dict({k: v for k, v in unet_kwargs(O.TINY).items() if k != "out_channels"},
                                                hint_channels=3)},
        "unet_config": {"target": "cldm.cldm.ControlledUnetModel", "params": unet_kwargs(O.TINY)},
        "first_stage_config": {"target": "ldm.models.autoencoder.AutoencoderKL",
                               "params": {"embed_dim": 4, "ddconfig": vae_kwargs(O.TINY_VAE)}},
        "cond_stage_config": {"target": "ldm.modules.encoders.modules.FrozenCLIPEmbedder"}}}}
    path = tmp_path / "cldm_tiny.yaml"
    path.write_text(yaml.safe_dump(cfg))
    model = create_model(str(path))
    assert model.cond_stage_model is None and model.scale_factor == 0.18215
    sd = {k: torch.randn_like(v) for k, v in model.state_dict().items()}
    torch.save({"state_dict": sd}, tmp_path / "w.pth")
    import safetensors.torch
    safetensors.torch.save_file(sd, str(tmp_path / "w.safetensors"))
    for name in ("w.pth", "w.safetensors"):
        loaded = load_state_dict(str(tmp_path / name), location="cpu")
        assert set(loaded) == set(sd) and all(torch.equal(loaded[k], sd[k]) for k in sd)
        model2 = create_model(str(path))
        missing, unexpected = model2.load_state_dict(loaded, strict=True)
        assert not missing and not unexpected
        k0 = "control_model.input_hint_block.0.weight"
        assert torch.equal(model2.state_dict()[k0], sd[k0])
    assert len(create_model().state_dict()) == 1166  # SD1.5 defaults: the reference modules' 1166 tensors


def test_engine_surface_shapes():
    """Engine.infer's binding names / static shapes (Engine.py:66-90, onnx2trt_static_plugin.py:79-115)."""
    from stablediffusioneo_b200.Engine import Engine
    from stablediffusioneo_b200.cldm.cldm import ControlLDM
    with torch.device("meta"):
        model = ControlLDM()
    unet = Engine(model, "unet").shape_dict()
    assert list(unet)[:3] == ["x_noisy", "timestep", "context"] and list(unet)[-1] == "latent"
    want = [(1, 320, 32, 48)] * 3 + [(1, 320, 16, 24), (1, 640, 16, 24), (1, 640, 16, 24), (1, 640, 8, 12), (1, 1280, 8, 12),
                                      (1, 1280, 8, 12)] + [(1, 1280, 4, 6)] * 4
    assert [unet[f"control{i}"] for i in range(13)] == want
    cn = Engine(model, "controlnet").shape_dict()
    assert list(cn)[:4] == ["x_noisy", "hint", "timestep", "context"] and len(cn) == 17
    assert cn["hint"] == (1, 3, 256, 384) and cn["context"] == (1, 77, 768)
    assert Engine(model, "decoder").shape_dict()["images"] == (1, 3, 256, 384)


@pytest.mark.parametrize("n,c,h,w", [(2, 320, 32, 48), (2, 2560, 8, 12), (1, 128, 256, 384), (16, 256, 256, 256),
                                     (3, 64, 7, 5), (1, 32, 1, 1), (5, 1280, 33, 17)])
def test_groupnorm_f16_visit_schedule(n, c, h, w):
    """The streamed GroupNorm's visit schedule (csrc/groupnorm_stream.cu): every tile gets exactly one statistics and one apply
    visit, the per-CTA sequences never deadlock on the per-sample ready flags, apply visits are spread evenly over the CTAs,
    tiles cover the sample, and the buffers fit the shared-memory budget."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    plan = (ctypes.c_int32 * 7)()
    assert lib.sdeo_groupnorm_f16_plan(n, h * w, c, 32, 148, plan) == 0
    chunks, ppc, lag, grid, smem, stride, bufs = list(plan)
    hw = h * w
    assert (chunks - 1) * ppc < hw <= chunks * ppc
    assert stride % 128 == 0 and stride >= ppc * c * 2 and 2 <= bufs <= 4 and bufs * stride < smem <= 220 * 1024
    tiles = n * chunks
    assert 1 <= grid <= min(148, tiles) and chunks <= lag <= tiles
    # replay the kernel's schedule: CTA b walks its visits in order; an apply visit may only wait for statistics visits that
    # a round-robin over CTAs (one visit each, blocked applies skipped) eventually completes -> no deadlock, all tiles served
    cap = 2 * tiles // grid + 8
    seqs = []
    for b_ in range(grid):
        out = (ctypes.c_int32 * (2 * cap))()
        cnt = lib.sdeo_groupnorm_f16_visits(b_, grid, tiles, lag, out, cap)
        assert 0 < cnt <= cap
        seqs.append([(out[2 * i], out[2 * i + 1]) for i in range(cnt)])
    # partial slots: thread group g of CTA b (it takes every ng-th unit of the CTA) writes slot (tile - sample * chunks) mod
    # (ng * grid) once per sample it has statistics tiles of: the same slot for all its tiles of the sample, and the slots of
    # a sample are exactly 0 .. min(chunks, ng * grid) - 1, each written by one (CTA, group)
    for ng in (1, 2):
        for img in range(n):
            owner = {}
            for b_ in range(grid):
                for k_, t_ in seqs[b_]:
                    if k_ == 0 and t_ // chunks == img:
                        key = (b_, ((t_ - b_) // grid) % ng)
                        owner.setdefault(key, set()).add((t_ - img * chunks) % (ng * grid))
            assert all(len(v) == 1 for v in owner.values())
            assert sorted(next(iter(v)) for v in owner.values()) == list(range(min(chunks, ng * grid)))
    kinds = [sum(1 for v in s_ if v[0] == 1) for s_ in seqs]
    assert max(kinds) - min(kinds) <= 1, "apply visits are not balanced over the CTAs"
    pos = [0] * grid
    stats_done, applied, seen_stats = [0] * n, set(), set()
    progress = True
    while progress:
        progress = False
        for b_ in range(grid):
            if pos[b_] == len(seqs[b_]):
                continue
            kind, t = seqs[b_][pos[b_]]
            assert 0 <= t < tiles
            if kind == 0:
                assert t not in seen_stats
                seen_stats.add(t)
                stats_done[t // chunks] += 1  # a CTA publishes its partial right after its last statistics tile of the sample
            else:
                if stats_done[t // chunks] < chunks:
                    continue  # the control warp spins on a partial slot that is not written yet
                assert t not in applied
                applied.add(t)
            pos[b_] += 1
            progress = True
    assert len(seen_stats) == tiles and len(applied) == tiles, "schedule deadlocks or skips tiles"
    assert lib.sdeo_groupnorm_f16_workspace_bytes(n, hw, c, 32) >= n * (min(chunks, 2 * grid) + 1) * 32 * 8


@pytest.mark.parametrize("n,c,h,w", [(2, 320, 32, 48), (2, 640, 16, 24), (2, 1280, 8, 12), (2, 2560, 8, 12), (3, 64, 7, 5),
                                     (1, 32, 9, 11), (1, 4096, 5, 3), (1, 512, 48, 96), (1, 64, 1, 1)])
def test_groupnorm_f16_slab_index_model(n, c, h, w):
    """Index model of gn_slab_kernel (csrc/groupnorm_stream.cu) from sdeo_groupnorm_f16_slab_plan: the thread -> (row
    phase, column) map loads every 16-byte vector of a slab exactly once, the two-stage fold of the per-thread column sums
    ([T][16] -> stage A [P][W] -> one warp per group) adds every channel of a group exactly once, and the pieces of a split
    slab normalise disjoint row ranges that cover the slab. Checked by pushing exact integers through the maps."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    plan = (ctypes.c_int32 * 7)()
    hw, groups = h * w, 32
    assert lib.sdeo_groupnorm_f16_slab_plan(n, hw, c, groups, 200, plan) == 0
    sg, sv, slabs, split, rows_per, T, smem = list(plan)
    cpg = c // groups
    assert slabs * sg == groups and sv * 8 == sg * cpg and T in (256, 512) and sv * 16 <= T
    assert hw * sv * 16 + T * 16 * 4 + T * 4 <= smem <= 220 * 1024
    assert split in (1, 2, 4, 8) and rows_per * split >= hw and (split == 1 or hw * sv * 16 >= 48 * 1024)
    R = T // sv
    rng = np.random.default_rng(hw * 131 + c)
    slab = rng.integers(-8, 9, size=(hw, sv * 8)).astype(np.int64)      # one slab's channels, exact integers
    # ---- load phase / statistics phase: thread tid owns column tj and rows tr, tr + R, ...
    loaded = np.zeros((hw, sv), dtype=np.int64)
    part = np.zeros((T, 16), dtype=np.int64)                            # [tid][8 sums | 8 sums of squares]
    for tid in range(T):
        tr, tj = divmod(tid, sv)
        if tr >= R:
            continue
        for r in range(tr, hw, R):
            loaded[r, tj] += 1
            v = slab[r, tj * 8:(tj + 1) * 8]
            part[tid, :8] += v
            part[tid, 8:] += v * v
    assert (loaded == 1).all()
    # ---- stage A: part viewed as [R][W]; thread group p of P sums rows p, p + P, ... per column
    W, P = sv * 16, T // (sv * 16)
    flat = part.reshape(-1)                                             # index tid * 16 + k == tr * W + tj * 16 + k
    part2 = np.zeros((P, W), dtype=np.int64)
    for p in range(P):
        for col in range(W):
            part2[p, col] = sum(flat[r * W + col] for r in range(p, R, P))
    # ---- stage B: warp g folds group g: cpg channels x P parts (lane-strided, then a shuffle tree: any order is exact here)
    for g in range(sg):
        ss = qq = 0
        for i in range(cpg * P):
            p, ch = divmod(i, cpg)
            ch += g * cpg
            col = (ch >> 3) * 16 + (ch & 7)
            ss += part2[p, col]
            qq += part2[p, col + 8]
        block = slab[:, g * cpg:(g + 1) * cpg]
        assert ss == block.sum() and qq == (block * block).sum()
    # ---- normalise phase: piece q of a split slab takes rows [q * rows_per, min(hw, (q + 1) * rows_per)), same thread map
    done = np.zeros((hw, sv), dtype=np.int64)
    for piece in range(split):
        r_hi = min(hw, (piece + 1) * rows_per)
        for tid in range(T):
            tr, tj = divmod(tid, sv)
            if tr >= R:
                continue
            for r in range(piece * rows_per + tr, r_hi, R):
                done[r, tj] += 1
    assert (done == 1).all()
    # the channel -> group map of the normalise phase stays inside the slab
    assert max((tj * 8 + j) // cpg for tj in range(sv) for j in range(8)) == sg - 1


def test_silu_pair_formula():
    """The GroupNorm apply pass computes SiLU for two values with ONE reciprocal (norm.cu silu_pair): r = 1 / ((1 + e0)(1 + e1)),
    y0 = x0 * (1 + e1) * r, with the exponent clamped to 63 so the product of the denominators stays finite. Restated in
    float32 numpy: finite for every pair (including both values far below -88, where 1 + e^-x alone overflows), and within
    a few float32 ulps of x / (1 + e^-x) everywhere the result is not below 1e-17 in magnitude."""
    f32 = np.float32
    xs = np.array([-1e4, -200.0, -89.0, -88.0, -60.0, -44.0, -43.5, -30.0, -10.0, -3.0, -1.0, -1e-3, 0.0, 1e-3, 0.5, 2.0, 10.0,
                   50.0, 90.0, 1e4], dtype=f32)
    x0, x1 = [a.ravel() for a in np.meshgrid(xs, xs)]
    k = f32(-1.4426950408889634)
    with np.errstate(over="raise", invalid="raise"):
        e0 = np.exp2(np.minimum(x0 * k, f32(63.0))).astype(f32)
        e1 = np.exp2(np.minimum(x1 * k, f32(63.0))).astype(f32)
        d0, d1 = f32(1.0) + e0, f32(1.0) + e1
        r = f32(1.0) / (d0 * d1)
        y0, y1 = x0 * (d1 * r), x1 * (d0 * r)
    assert np.isfinite(y0).all() and np.isfinite(y1).all()
    for x, y in ((x0, y0), (x1, y1)):
        with np.errstate(over="ignore"):   # e^1e4 = inf in float64 too: gold = -0.0 there
            gold = x.astype(np.float64) / (1.0 + np.exp(-x.astype(np.float64)))
        big = np.abs(gold) > 1e-17
        assert np.all(np.abs(y[big] - gold[big]) <= 4e-6 * np.abs(gold[big]) + 1e-30)   # (fp32 exponent argument: ~1e-6 at x = -30)
        assert np.all(np.abs(y[~big]) < 1e-13)     # clamped tail: ~x * 2^-63, zero in any 16-bit output


def test_groupnorm_two_pass_plan_invariants():
    """sdeo_groupnorm_plan (the grid of the statistics + apply kernels): the chunks cover the sample exactly once, stay within
    the caps (384 per sample; two waves of three CTAs per SM over a big batch), the workspace query covers every row size,
    and the tuned points of the round-2 sweep come out (one CTA per 64 KB; 222 CTAs for small tensors)."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    plan = (ctypes.c_int32 * 2)()
    for n in (1, 2, 3, 5, 16, 40):
        for hw in (1, 7, 96, 1536, 4096, 6144, 24576, 98304, 262144):
            for row_bytes in (16, 256, 512, 1024, 2560, 5120, 10240):
                assert lib.sdeo_groupnorm_plan(n, hw, row_bytes, plan) == 0
                chunks, ppc = plan[0], plan[1]
                assert chunks >= 1 and ppc >= 1 and chunks * ppc >= hw and (chunks - 1) * ppc < hw
                assert chunks <= 384
                total = n * hw * row_bytes
                if total // (64 << 10) > 148 * 6:       # big: whole waves of the resident set
                    assert n * chunks <= 148 * 6 or chunks == 1
                assert lib.sdeo_groupnorm_workspace_bytes(n, hw, 32) >= n * chunks * 32 * 8
    assert lib.sdeo_groupnorm_plan(16, 512 * 512, 256, plan) == 0 and plan[0] == 55          # configs[4]: 880 CTAs
    assert lib.sdeo_groupnorm_plan(1, 256 * 384, 256, plan) == 0 and (plan[0], plan[1]) == (384, 256)   # 25 MB: 64 KB per CTA
    assert lib.sdeo_groupnorm_plan(1, 128 * 192, 512, plan) == 0 and plan[0] == 222          # 12.6 MB: the small-tensor floor
    assert lib.sdeo_groupnorm_plan(2, 32 * 48, 1280, plan) == 0 and (plan[0], plan[1]) == (96, 16)   # UNet: >= 16 rows per CTA
    assert lib.sdeo_groupnorm_plan(2, 64 * 96, 640, plan) == 0 and (plan[0], plan[1]) == (110, 56)   # ~1.5 CTAs per SM
    assert lib.sdeo_groupnorm_plan(0, 4, 16, plan) != 0


@pytest.mark.parametrize("c,elt", [(128, 2), (320, 4), (512, 2), (96, 2), (2560, 4), (4096, 2)])
def test_groupnorm_two_pass_row_batches_model(c, elt):
    """Row schedule of the pipelined two-pass kernels (norm.cu gn_load_batch / gn_apply_column): a CTA of 256 threads is
    (R row phases) x (cols channel vectors); thread (tr, tv) walks rows pp = p_begin + tr, + H * R, ... and each step takes the
    H rows pp + u * R that lie inside the chunk (the rest read as zero). Every (row, vector) of every chunk of the plan is
    taken exactly once, for the 16-bit (H = 4) and fp32 (H = 2) batch depths, including more channel vectors than threads."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    plan = (ctypes.c_int32 * 2)()
    H = 4 if elt == 2 else 2
    cv = c // 8
    cols = min(cv, 256)
    R = 256 // cols
    for n, hw in ((1, 37), (2, 1536), (1, 6144), (16, 4096), (3, 1000)):
        assert lib.sdeo_groupnorm_plan(n, hw, c * elt, plan) == 0
        chunks, ppc = plan[0], plan[1]
        seen = np.zeros((hw, cv), dtype=np.int32)
        for chunk in range(chunks):
            p_begin, p_end = chunk * ppc, min(hw, (chunk + 1) * ppc)
            for vbase in range(0, cv, cols):
                for tid in range(R * cols):
                    tr, tv = divmod(tid, cols)
                    v = vbase + tv
                    if v >= cv:
                        continue
                    pp = p_begin + tr
                    while pp < p_end:
                        for u in range(H):
                            if pp + u * R < p_end:
                                seen[pp + u * R, v] += 1
                        pp += H * R
        assert (seen == 1).all(), (n, hw, chunks, ppc)


def test_groupnorm_f16_variant_selection():
    """sdeo_groupnorm_f16_variant: UNet samples at 256x384 go to the slab kernel (one CTA per slab of whole groups); with the
    slab kernel ruled out they fit a cluster (resident kernel: the cluster covers the sample and fits the shared-memory
    budget); VAE-sized samples take two launches, the streamed kernel only when asked for."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    info = (ctypes.c_int32 * 3)()
    unet = [(2, 320, 32, 48), (2, 640, 16, 24), (2, 1280, 8, 12), (2, 2560, 8, 12), (1, 64, 1, 1), (3, 64, 7, 5)]
    saved = os.environ.pop("SDEO_GN_F16_VARIANT", None)
    saved_kb = os.environ.pop("SDEO_GN_F16_SLAB_KB", None)
    try:
        for n, c, h, w in unet:
            assert lib.sdeo_groupnorm_f16_variant(n, h * w, c, 32, 148, 8, info) == 3
            sg, sv, smem = list(info)
            cpg = c // 32
            assert sg in (1, 2, 4, 8) and (sg * cpg) % 8 == 0 and sv == sg * cpg // 8 and 32 % sg == 0
            assert h * w * sv * 16 <= 128 * 1024 and h * w * sv * 16 < smem <= 220 * 1024
        assert lib.sdeo_groupnorm_f16_variant(1, 48 * 96, 512, 32, 148, 16, info) == 1    # 147 KB slab: beyond the default limit
        os.environ["SDEO_GN_F16_SLAB_KB"] = "200"
        assert lib.sdeo_groupnorm_f16_variant(1, 48 * 96, 512, 32, 148, 16, info) == 3
        assert lib.sdeo_groupnorm_f16_variant(1, 64 * 96, 512, 32, 148, 16, info) == 1    # 196 KB + fold scratch: does not fit
        os.environ.pop("SDEO_GN_F16_SLAB_KB")
        os.environ["SDEO_GN_F16_VARIANT"] = "resident"
        for n, c, h, w in unet:
            assert lib.sdeo_groupnorm_f16_variant(n, h * w, c, 32, 148, 8, info) == 2
            cs, rpc, smem = list(info)
            assert cs in (1, 2, 4, 8) and rpc * cs >= h * w and (rpc - 1) * cs < h * w + cs and rpc * c * 2 < smem <= 220 * 1024
        # beyond a cluster: the two-launch grid by default, the streamed kernel only when asked for
        assert lib.sdeo_groupnorm_f16_variant(2, 32 * 48, 640, 32, 148, 8, info) == 1   # 1.9 MB per sample: needs 16 CTAs
        assert lib.sdeo_groupnorm_f16_variant(2, 32 * 48, 640, 32, 148, 16, info) == 2 and info[0] == 16
        assert lib.sdeo_groupnorm_f16_variant(16, 256 * 256, 256, 32, 148, 16, info) == 1
        os.environ.pop("SDEO_GN_F16_VARIANT")
        assert lib.sdeo_groupnorm_f16_variant(16, 256 * 256, 256, 32, 148, 16, info) == 1
        os.environ["SDEO_GN_F16_VARIANT"] = "stream"
        assert lib.sdeo_groupnorm_f16_variant(2, 32 * 48, 960, 32, 148, 8, info) == 0   # 2.9 MB: streamed
        assert lib.sdeo_groupnorm_f16_variant(16, 256 * 256, 256, 32, 148, 16, info) == 0
        assert lib.sdeo_groupnorm_f16_variant(1, 64, 8192, 32, 148, 8, info) == 1       # more channel vectors than threads
    finally:
        os.environ.pop("SDEO_GN_F16_VARIANT", None)
        os.environ.pop("SDEO_GN_F16_SLAB_KB", None)
        if saved is not None:
            os.environ["SDEO_GN_F16_VARIANT"] = saved
        if saved_kb is not None:
            os.environ["SDEO_GN_F16_SLAB_KB"] = saved_kb


def _gn_stream_protocol_model(lib, n, c, h, w, ng, bufs, sms=148, verbose=False):
    plan = (ctypes.c_int32 * 7)()
    assert lib.sdeo_groupnorm_f16_plan(n, h * w, c, 32, sms, plan) == 0
    chunks, ppc, lag, grid = plan[0], plan[1], plan[2], plan[3]
    tiles = n * chunks
    two_level = lag >= chunks + grid or lag == tiles
    W = ng * grid
    P = min(chunks, W)
    team = 16
    M = -(-P // team)
    tree = two_level and lag == tiles and P > 2 * team  # gs_plan's rule
    mids = [set() for _ in range(n)]
    next_team = list(range(grid))
    cap = 2 * tiles // grid + 8
    seqs = []
    for b in range(grid):
        out = (ctypes.c_int32 * (2 * cap))()
        cnt = lib.sdeo_groupnorm_f16_visits(b, grid, tiles, lag, out, cap)
        seqs.append([(out[2 * i], out[2 * i + 1]) for i in range(cnt)])
    unit = lambda kind, t: t + lag if kind else t
    published = [set() for _ in range(n)]
    finals = [False] * n
    full = [[False] * len(s) for s in seqs]
    released = [[False] * len(s) for s in seqs]
    kc = [0] * grid
    next_fold = list(range(grid))
    gl = [[[k for k, (kind, t) in enumerate(s) if ((unit(kind, t) - b) // grid) % ng == g] for g in range(ng)] for b, s in enumerate(seqs)]
    pg = [[0] * ng for _ in range(grid)]
    prev = [[None] * ng for _ in range(grid)]  # deferred release (ng == 1)
    progress = True
    while progress:
        progress = False
        for b in range(grid):
            s = seqs[b]
            # control warp
            while kc[b] < len(s):
                k = kc[b]
                kind, t = s[k]
                u = unit(kind, t)
                blocked = False
                while tree and next_team[b] < n * M and u >= (next_team[b] // M + 1) * chunks:
                    f, m = divmod(next_team[b], M)
                    if not all(j in published[f] for j in range(m * team, min(P, (m + 1) * team))):
                        blocked = True
                        break
                    mids[f].add(m)
                    next_team[b] += grid
                    progress = True
                while not blocked and two_level and next_fold[b] < n and u >= (next_fold[b] + 1) * chunks:
                    if (len(mids[next_fold[b]]) < M) if tree else (len(published[next_fold[b]]) < P):
                        blocked = True
                        break
                    finals[next_fold[b]] = True
                    next_fold[b] += grid
                    progress = True
                if blocked:
                    break
                if k >= bufs and not released[b][k - bufs]:
                    break
                if kind == 1:
                    img = t // chunks
                    if (two_level and not finals[img]) or (not two_level and len(published[img]) < P):
                        break
                full[b][k] = True
                kc[b] += 1
                progress = True
            # thread groups
            for g in range(ng):
                while pg[b][g] < len(gl[b][g]):
                    k = gl[b][g][pg[b][g]]
                    if not full[b][k]:
                        break
                    kind, t = s[k]
                    if kind == 0:
                        released[b][k] = True
                        img = t // chunks
                        nxt = t + W
                        if nxt >= tiles or nxt // chunks != img:
                            j = (t - img * chunks) % W
                            assert j not in published[img], (b, g, t, j)
                            published[img].add(j)
                    else:
                        if ng > 1:
                            released[b][k] = True
                    if ng == 1:
                        if prev[b][g] is not None:
                            released[b][prev[b][g]] = True
                        prev[b][g] = k if kind == 1 else None
                    pg[b][g] += 1
                    progress = True
    done = all(pg[b][g] == len(gl[b][g]) for b in range(grid) for g in range(ng))
    if not done and verbose:
        for b in range(grid):
            if kc[b] < len(seqs[b]):
                k = kc[b]; kind, t = seqs[b][k]
                print("cta", b, "control at k", k, kind, t, "unit", unit(kind, t), "next_fold", next_fold[b], "pg", pg[b], [gl[b][g][pg[b][g]] if pg[b][g] < len(gl[b][g]) else None for g in range(ng)])
                if b > 12: break
        print("published", [len(p) for p in published], "P", P, "finals", finals, "chunks", chunks, "lag", lag, "grid", grid)
    return done


@pytest.mark.parametrize("shape", [(2, 320, 32, 48), (1, 512, 64, 96), (3, 64, 7, 5), (40, 320, 16, 24), (1, 64, 1, 1),
                                   (5, 1280, 33, 17), (4, 256, 128, 128), (8, 256, 256, 256)])
def test_groupnorm_f16_protocol_model(shape):
    """Discrete-event model of the streamed GroupNorm's protocol (csrc/groupnorm_stream.cu) on the library's own plan and visit
    lists: an in-order control warp per CTA (team and folder duties, buffer recycling, statistics wait before an apply visit), one or two
    thread groups, partial slots published once per (CTA, group, sample). Every visit must complete for every buffer count:
    no deadlock, no slot written twice."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    for ng in (1, 2):
        for bufs in (2, 4):
            assert _gn_stream_protocol_model(lib, *shape, ng=ng, bufs=bufs), (shape, ng, bufs)


# ---- conv launch plans (host logic of csrc/gemm_conv.cu, no device work) ----------------------------------------
_PLAN_SHAPES = [
    # (n, h, w, c1, c2, cout, ksize, stride, epi_mode, stream)
    (2, 32, 48, 320, 0, 320, 3, 1, 0, 1), (2, 32, 48, 320, 0, 320, 1, 1, 0, 1), (2, 32, 48, 640, 320, 320, 3, 1, 0, 0),
    (2, 16, 24, 640, 0, 640, 3, 1, 0, 1), (2, 8, 12, 1280, 0, 1280, 3, 1, 0, 1), (2, 4, 6, 1280, 0, 1280, 3, 1, 0, 1),
    (2, 32, 48, 320, 0, 320, 3, 2, 0, 1), (1, 1, 3072, 320, 0, 2560, 1, 1, 1, 0), (1, 1, 192, 5120, 0, 1280, 1, 1, 0, 1),
    (8, 96, 96, 320, 0, 320, 3, 1, 0, 1), (8, 48, 48, 640, 0, 640, 3, 1, 0, 1), (4, 512, 512, 128, 0, 128, 3, 1, 0, 0),
    (4, 128, 128, 512, 0, 512, 3, 1, 0, 0), (1, 1, 73728, 320, 0, 2560, 1, 1, 1, 0), (16, 512, 512, 128, 0, 3, 3, 1, 0, 0),
]


def _describe(lib, _lib, shape, halo):
    n, h, w, c1, c2, cout, k, stride, epi, stream = shape
    a = _lib.ConvArgs()
    P = ctypes.c_void_p(256)
    a.x1, a.w_packed, a.y = P, P, P
    a.n, a.h, a.w, a.c1, a.ld1, a.c2, a.ld2 = n, h, w, c1, c1, c2, c2
    if c2:
        a.x2 = P
    a.cout, a.ksize, a.stride, a.pad, a.scale = cout, k, stride, k // 2, 1.0
    a.ldy = (cout // 2 if epi == 1 else cout + (-cout) % 8)
    a.epi_mode = epi
    if stream:
        a.residual, a.ldr, a.residual_f32, a.y_fp32, a.y2, a.ldy2 = P, cout, 1, 1, P, cout
    out = (ctypes.c_int32 * 16)()
    rc = lib.sdeo_conv_plan_describe(ctypes.byref(a), halo, out, 16)
    return rc, list(out)


@pytest.mark.parametrize("env", [{}, {"SDEO_PAIR": "0"}, {"SDEO_PAIR": "1"}, {"SDEO_OCC2": "1"}, {"SDEO_OCC2": "0", "SDEO_PAIR": "0"}])
def test_conv_plan_invariants(env, monkeypatch):
    """Every plan the host logic hands to conv_gemm_kernel keeps the kernel's structural assumptions: ring depth a multiple
    of the producer count (a ring slot is always refilled by the same thread), shared memory within the per-SM budget
    (half of it for two-CTA-per-SM plans), accumulator columns >= N tile, an even M-tile grid for CTA pairs."""
    from stablediffusioneo_b200 import _lib
    lib = _lib.load()
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    seen = 0
    for shape in _PLAN_SHAPES:
        for halo in (0, 1):
            rc, o = _describe(lib, _lib, shape, halo)
            if rc != 0:
                continue  # this tiling is not available for the shape (1x1, stride 2, forced pair with one M tile, ...)
            seen += 1
            bn, splits, flags, bn_, bh, bw, m_tiles, n_tiles, stages, a_stages, a_stage_bytes, smem, rows, tmem, pitch, cps = o
            is_halo, pair, occ2, nprod = flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, flags >> 4
            assert 1 <= nprod <= 4 and stages >= 2 and stages % nprod == 0, (shape, o)
            assert smem <= (112 if occ2 else 227) * 1024, (shape, o)
            assert tmem >= bn and tmem in (32, 64, 128, 256), (shape, o)
            assert 16 <= bn <= 256 and bn % 16 == 0 and rows <= 128 and 1 <= splits <= 8, (shape, o)
            assert is_halo == halo
            if is_halo:
                assert a_stages >= 2 and pitch == bw + 2 and a_stage_bytes % 1024 == 0 and a_stage_bytes >= (bh + 2) * pitch * 128
            if occ2:
                assert splits == 1
            if pair:
                assert splits <= 4
    assert seen >= 15


def _simulate_ring(stages, nprod, steps, rng):
    """Discrete model of the conv mainloop's ring: `nprod` producers (K step i belongs to producer i % nprod) and one
    consumer, synchronised by per-slot full / empty mbarriers that waiters test by PHASE PARITY (a waiter cannot tell a
    phase from the one two later). Returns None, or a description of the first hazard (a slot refilled before its previous
    contents were consumed, or consumed before it was filled). Random interleaving of the actors."""
    full_done = [0] * stages    # completed phases of full[s]
    empty_done = [0] * stages   # completed phases of empty[s]
    written = [0] * stages      # fills of slot s so far
    consumed = [0] * stages
    prod_next = [p for p in range(nprod)]   # next K step of producer p
    cons_next = 0
    while cons_next < steps:
        actors = []
        for p in range(nprod):
            i = prod_next[p]
            if i >= steps:
                continue
            s, rnd = i % stages, i // stages
            # wait(empty[s], parity (rnd & 1) ^ 1) when rnd > 0: passes while the barrier's phase parity differs
            if rnd == 0 or (empty_done[s] & 1) != ((rnd & 1) ^ 1):
                actors.append(("p", p))
        s, rnd = cons_next % stages, cons_next // stages
        if (full_done[s] & 1) != (rnd & 1):
            actors.append(("c", 0))
        if not actors:
            return f"deadlock at consumer step {cons_next}"
        kind, p = actors[rng.integers(len(actors))]
        if kind == "p":
            i = prod_next[p]
            s, rnd = i % stages, i // stages
            if consumed[s] != rnd:
                return f"producer {p} refills slot {s} for step {i} before step {i - stages} was consumed"
            written[s] += 1
            full_done[s] += 1
            prod_next[p] += nprod
        else:
            if written[s] != rnd + 1:
                return f"consumer reads slot {s} for step {cons_next} before it was filled"
            consumed[s] += 1
            empty_done[s] += 1
            cons_next += 1
    return None


def test_conv_ring_protocol_model():
    """No interleaving produces a hazard as long as the ring is at least as deep as the producer count (the in-order
    consumer then bounds how far a producer can run ahead to less than two phases of any slot) -- make_plan keeps
    nprod <= stages, and the depth a multiple of nprod so that a slot is always refilled by the same thread. With more
    producers than slots a producer two phases ahead passes its parity wait and overwrites live data."""
    rng = np.random.default_rng(7)
    for stages, nprod in ((2, 2), (3, 3), (4, 4), (8, 4), (12, 4), (4, 2), (6, 3), (4, 1), (5, 1), (5, 4), (7, 4), (5, 3)):
        for _ in range(200):
            assert _simulate_ring(stages, nprod, 61, rng) is None, (stages, nprod)
    for stages, nprod in ((2, 4), (3, 4), (2, 3)):
        bad = [_simulate_ring(stages, nprod, 61, rng) for _ in range(200)]
        assert any(b is not None for b in bad), (stages, nprod)
