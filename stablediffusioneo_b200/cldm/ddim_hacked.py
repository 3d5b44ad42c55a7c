"""DDIMSampler with the reference's interface (cldm/ddim_hacked.py): make_schedule, sample, ddim_sampling,
p_sample_ddim -> (samples, intermediates).

Two execution paths, same arithmetic:
  * generic: p_sample_ddim calls model.apply_model for cond and uncond (any duck-typed model) and runs the CFG
    combine + DDIM update as ONE fused kernel (the reference: ~12 elementwise launches + 4 host-syncing torch.full);
  * engine (model is our ControlLDM, eta == 0, standard options): cond and uncond run as one batch of 2B through
    ControlNet+UNet (one weight pass), hint features and cross-attention K/V are hoisted out of the loop, the whole
    step — timestep embedding to x_{t-1} — is one CUDA graph replayed per step, with the step's timestep and DDIM
    coefficients read from device tables through a device-side step counter (no host patching between replays).
"""
import math
import os

import numpy as np
import torch
from tqdm import tqdm

from .. import ops
from ..ldm.modules.diffusionmodules.util import (BF16, make_ddim_sampling_parameters, make_ddim_timesteps, noise_like,
                                                 extract_into_tensor, nhwc, to_internal)


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_engine = True     # set False to force the generic two-call path
        self.use_cuda_graph = True
        self._engine = None

    def register_buffer(self, name, attr):
        # the reference hard-codes cuda (ddim_hacked.py:17-21); follow the model's device instead
        if type(attr) == torch.Tensor and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        self.register_buffer('betas', to_torch(self.model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
        ac = alphas_cumprod.detach().cpu()
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(ac)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - ac)))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - ac)))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / ac)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / ac - 1)))
        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.register_buffer('ddim_sigmas', ddim_sigmas)
        self.register_buffer('ddim_alphas', ddim_alphas)
        self.register_buffer('ddim_alphas_prev', ddim_alphas_prev)
        self.register_buffer('ddim_sqrt_one_minus_alphas', np.sqrt(1. - ddim_alphas))
        acp = self.model.alphas_cumprod_prev.detach().cpu()
        sigmas_for_original_sampling_steps = ddim_eta * torch.sqrt((1 - acp) / (1 - ac) * (1 - ac / acp))
        self.register_buffer('ddim_sigmas_for_original_num_steps', sigmas_for_original_sampling_steps)
        # host copies of the per-step scalars (the reference reads them from device tensors with a sync per step)
        self._h = dict(alphas=np.asarray(ddim_alphas, dtype=np.float64), alphas_prev=np.asarray(ddim_alphas_prev, dtype=np.float64),
                       sigmas=np.asarray(ddim_sigmas, dtype=np.float64),
                       sqrt_one_minus_alphas=np.asarray(np.sqrt(1. - ddim_alphas), dtype=np.float64))

    # --------------------------------------------------------------------------------------------------------
    def _coef_row(self, index, scale, use_original_steps=False):
        """The six scalars of one step (ddim_hacked.py:208-230) as the fused kernel's coefficient row.
        use_original_steps (:203-206): the 1000-step DDPM tables instead of the DDIM subsequence. (The reference reads the
        sigmas from `self.model.ddim_sigmas_for_original_num_steps`, a buffer that lives on the SAMPLER -- its own
        use_original_steps path dies with an AttributeError; the sampler's buffer is what it meant.)"""
        if use_original_steps:
            a_t = float(self.alphas_cumprod[index])
            a_prev = float(self.alphas_cumprod_prev[index])
            sigma_t = float(self.ddim_sigmas_for_original_num_steps[index])
            sqrt_1m_at = float(self.sqrt_one_minus_alphas_cumprod[index])
            return [float(scale), sqrt_1m_at, 1.0 / math.sqrt(a_t), math.sqrt(a_prev),
                    math.sqrt(max(1.0 - a_prev - sigma_t ** 2, 0.0)), sigma_t, 0.0, 0.0]
        a_t = float(self._h["alphas"][index])
        a_prev = float(self._h["alphas_prev"][index])
        sigma_t = float(self._h["sigmas"][index])
        sqrt_1m_at = float(self._h["sqrt_one_minus_alphas"][index])
        return [float(scale), sqrt_1m_at, 1.0 / math.sqrt(a_t), math.sqrt(a_prev),
                math.sqrt(max(1.0 - a_prev - sigma_t ** 2, 0.0)), sigma_t, 0.0, 0.0]

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None, **kwargs):
        if conditioning is not None and isinstance(conditioning, dict):
            ctmp = conditioning[list(conditioning.keys())[0]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f'Data shape for DDIM sampling is {size}, eta {eta}')
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  dynamic_threshold=dynamic_threshold, ucg_schedule=ucg_schedule, verbose=verbose)

    def _engine_ok(self, cond, uncond, scale, mask, callback, img_callback, quantize_denoised, score_corrector,
                   dynamic_threshold, ucg_schedule, ddim_use_original_steps, timesteps, noise_dropout, temperature):
        from .cldm import ControlLDM
        if not (self.use_engine and isinstance(self.model, ControlLDM)) or self.model.precision != "bf16":
            return False
        if any(v is not None for v in (callback, img_callback, score_corrector, dynamic_threshold, ucg_schedule,
                                       timesteps)) or quantize_denoised or ddim_use_original_steps:
            return False
        # (mask / x0 inpainting blend runs in the engine: q_sample(x0, t) of every step is drawn up front into a table)
        if noise_dropout != 0.0:
            return False  # (eta > 0 itself runs in the engine: the per-step noise is drawn up front into a device table)
        if not isinstance(cond, dict) or self.model.parameterization != "eps":
            return False
        conds = [cond]
        if uncond is not None and scale != 1.0:
            if not isinstance(uncond, dict):
                return False
            if cond["c_concat"] is None and uncond["c_concat"] is not None:
                return False  # (a hint on the unconditional branch only: no caller does that -> generic path)
            # cond with a hint, uncond without = guess mode (canny2image_torch.py:48): the engine runs the ControlNet on the
            # conditional rows only and adds its outputs onto those rows of the UNet's skip tensors
            conds.append(uncond)
        for cd in conds:
            # the engine's static buffers are sized from ONE context and ONE hint tensor per conditioning; lists with
            # several entries (concatenated along dim 1 by apply_model, cldm/cldm.py:331-336) take the generic path
            if len(cd["c_crossattn"]) != 1 or (cd["c_concat"] is not None and len(cd["c_concat"]) != 1):
                return False
        return True

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None, verbose=True):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        if self._engine_ok(cond, unconditional_conditioning, unconditional_guidance_scale, mask, callback, img_callback,
                           quantize_denoised, score_corrector, dynamic_threshold, ucg_schedule,
                           ddim_use_original_steps, timesteps, noise_dropout, temperature):
            if mask is not None:
                assert x0 is not None
            return self._engine_sampling(cond, unconditional_conditioning, unconditional_guidance_scale, img,
                                         log_every_t, temperature, mask=mask, x0=x0)

        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif timesteps is not None and not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        # ddim_hacked.py:144-145: all ddpm_num_timesteps steps, latest first, when ddim_use_original_steps
        time_range = list(reversed(range(0, timesteps))) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        iterator = tqdm(time_range, desc='DDIM Sampler', total=total_steps, disable=not verbose)
        for i, step in enumerate(iterator):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                # inpainting blend (ddim_hacked.py:154-157): img = q_sample(x0, ts) * mask + (1 - mask) * img, one kernel
                assert x0 is not None
                img = self._mask_blend(img, x0, mask, ts)
            if ucg_schedule is not None:
                assert len(ucg_schedule) == len(time_range)
                unconditional_guidance_scale = ucg_schedule[i]
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, use_original_steps=ddim_use_original_steps,
                                              quantize_denoised=quantize_denoised, temperature=temperature,
                                              noise_dropout=noise_dropout, score_corrector=score_corrector,
                                              corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              dynamic_threshold=dynamic_threshold)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, dynamic_threshold=None):
        """ddim_hacked.py:181-231. eps from model.apply_model (cond, then uncond); CFG + x0 + x_{t-1} in one kernel."""
        if quantize_denoised or score_corrector is not None or dynamic_threshold is not None \
                or noise_dropout > 0. or self.model.parameterization != "eps":
            raise NotImplementedError("option not on the ControlNet-SD1.5 path")
        guided = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        e_c = self.model.apply_model(x, t, c).contiguous()
        e_u = self.model.apply_model(x, t, unconditional_conditioning).contiguous() if guided else None
        row = self._coef_row(index, unconditional_guidance_scale if guided else 1.0, use_original_steps)
        coef = torch.tensor([row], dtype=torch.float32, device=x.device)
        noise = None
        if row[5] != 0.0:
            noise = (noise_like(x.shape, x.device, repeat_noise) * temperature).contiguous()
        pred_x0 = torch.empty_like(x)
        x_prev, _ = ops.cfg_ddim_step(e_c, e_u, x.contiguous(), coef, noise=noise, pred_x0=pred_x0)
        return x_prev, pred_x0

    def _mask_blend(self, img, x0, mask, ts):
        """img_orig = model.q_sample(x0, ts) (the model's own forward-diffusion draw, as in the reference), then
        img_orig * mask + (1 - mask) * img as one pass."""
        dev = img.device
        img_orig = self.model.q_sample(x0.to(device=dev, dtype=torch.float32), ts).to(device=dev, dtype=torch.float32).contiguous()
        mask = mask.to(device=dev, dtype=torch.float32)
        if mask.dim() != 4 or mask.shape[0] != img.shape[0] or mask.shape[1] not in (1, img.shape[1]) \
                or mask.shape[2:] != img.shape[2:]:
            mask = mask.expand_as(img)
        one = torch.ones((img.shape[0],), dtype=torch.float32, device=dev)
        return ops.mask_blend(img_orig, img_orig, img.contiguous(), mask.contiguous(), one, torch.zeros_like(one))

    def _guided_eps(self, x, t, c, scale, uc):
        """eps for encode(): the reference concatenates (uncond, cond) into one batch with torch.cat, which only works for
        tensor conditionings (ddim_hacked.py:257-262); dict conditionings (ControlLDM) are evaluated as two calls."""
        if scale == 1.:
            return self.model.apply_model(x, t, c)
        assert uc is not None
        if isinstance(c, dict) or isinstance(uc, dict):
            e_u = self.model.apply_model(x, t, uc).contiguous()
            e_c = self.model.apply_model(x, t, c).contiguous()
        else:
            e_u, e_c = torch.chunk(self.model.apply_model(torch.cat((x, x)), torch.cat((t, t)), torch.cat((uc, c))), 2)
            e_u, e_c = e_u.contiguous(), e_c.contiguous()
        return ops.axpby(e_u, e_c, 1.0 - scale, scale)

    @torch.no_grad()
    def encode(self, x0, c, t_enc, use_original_steps=False, return_intermediates=None, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, callback=None):
        """Deterministic DDIM inversion (ddim_hacked.py:233-276): x_next = sqrt(a_next / a) x + sqrt(a_next) (sqrt(1/a_next
        - 1) - sqrt(1/a - 1)) eps, one fused pass per step. Needs make_schedule() first, like the reference."""
        timesteps = np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps
        assert t_enc <= timesteps.shape[0]
        num_steps = t_enc
        if use_original_steps:   # ddim_hacked.py:242-244
            alphas_next = self.alphas_cumprod[:num_steps].cpu().numpy().astype(np.float64)
            alphas = self.alphas_cumprod_prev[:num_steps].cpu().numpy().astype(np.float64)
        else:
            alphas_next, alphas = self._h["alphas"][:num_steps], self._h["alphas_prev"][:num_steps]
        x_next = x0.to(device=self.model.device, dtype=torch.float32).contiguous()
        intermediates, inter_steps = [], []
        for i in tqdm(range(num_steps), desc='Encoding Image', disable=True):
            t = torch.full((x0.shape[0],), int(timesteps[i]), device=self.model.device, dtype=torch.long)
            noise_pred = self._guided_eps(x_next, t, c, unconditional_guidance_scale, unconditional_conditioning)
            an, a = float(alphas_next[i]), float(alphas[i])
            x_next = ops.axpby(x_next, noise_pred.contiguous(), math.sqrt(an / a),
                               math.sqrt(an) * (math.sqrt(1 / an - 1) - math.sqrt(1 / a - 1)))
            if return_intermediates and i % (num_steps // return_intermediates) == 0 and i < num_steps - 1:
                intermediates.append(x_next)
                inter_steps.append(i)
            elif return_intermediates and i >= num_steps - 2:
                intermediates.append(x_next)
                inter_steps.append(i)
            if callback:
                callback(i)
        out = {'x_encoded': x_next, 'intermediate_steps': inter_steps}
        if return_intermediates:
            out.update({'intermediates': intermediates})
        return x_next, out

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """sqrt(a_t) x0 + sqrt(1 - a_t) noise with t [B] indexing the DDIM tables, or the 1000-step tables when
        use_original_steps (ddim_hacked.py:278-292)."""
        dev = self.model.device
        x0 = x0.to(device=dev, dtype=torch.float32).contiguous()
        a = (self.alphas_cumprod if use_original_steps else torch.as_tensor(self._h["alphas"], dtype=torch.float32)).to(dev)
        a = a[t.to(dev)]
        if noise is None:
            noise = torch.randn_like(x0)
        return ops.axpby(x0, noise.to(device=dev, dtype=torch.float32).contiguous(), a.sqrt(), (1.0 - a).sqrt())

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False, callback=None):
        """p_sample_ddim over the first t_start DDIM timesteps, latest first (ddim_hacked.py:294-317)."""
        timesteps = (np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps)[:t_start]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        x_dec = x_latent.to(device=self.model.device, dtype=torch.float32)
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((x_latent.shape[0],), int(step), device=x_dec.device, dtype=torch.long)
            x_dec, _ = self.p_sample_ddim(x_dec, cond, ts, index=index, use_original_steps=use_original_steps,
                                          unconditional_guidance_scale=unconditional_guidance_scale,
                                          unconditional_conditioning=unconditional_conditioning)
            if callback:
                callback(i)
        return x_dec

    # --------------------------------------------------------------------------------------------------------
    def _engine_sampling(self, cond, uncond, scale, x_T, log_every_t, temperature=1.0, mask=None, x0=None):
        guided = not (uncond is None or scale == 1.)
        S = int(self.ddim_timesteps.shape[0])
        eng = self._engine
        key = _EngineKey(self.model, x_T, cond, uncond if guided else None, S, self.use_cuda_graph, mask is not None)
        if eng is None or eng.key != key:
            eng = self._engine = _Engine(self.model, key, self.use_cuda_graph)
        time_range = np.flip(self.ddim_timesteps)
        rows = [self._coef_row(S - i - 1, scale if guided else 1.0) for i in range(S)]
        # the reference logs x_inter / pred_x0 when index % log_every_t == 0 or index == S-1 (ddim_hacked.py:174-176)
        log_at = [i for i in range(S) if (S - i - 1) % log_every_t == 0 or i == 0]
        blend = None
        if mask is not None:
            dev = x_T.device
            m = mask.to(device=dev, dtype=torch.float32)
            if m.dim() != 4 or m.shape[0] != x_T.shape[0] or m.shape[1] not in (1, x_T.shape[1]) or m.shape[2:] != x_T.shape[2:]:
                m = m.expand_as(x_T)
            blend = (m.contiguous(), x0.to(device=dev, dtype=torch.float32).contiguous(), self.model)
        img, pred_x0, logged = eng.run(x_T, cond, uncond if guided else None, [int(t) for t in time_range], rows, log_at,
                                       temperature=temperature, blend=blend)
        inter = {'x_inter': [x_T] + [l[0] for l in logged], 'pred_x0': [x_T] + [l[1] for l in logged]}
        return img, inter


class _EngineKey:
    def __init__(self, model, x_T, cond, uncond, S, graph, masked=False):
        hint = cond["c_concat"]
        self.masked = masked
        # weights_fingerprint: the captured graphs, time-embedding tables and hoisted K/V bake in buffers derived from the
        # parameters; after load_state_dict() or any in-place update the engine is rebuilt instead of replaying old weights
        self.t = (id(model), tuple(x_T.shape), tuple(cond["c_crossattn"][0].shape),
                  None if hint is None else tuple(hint[0].shape), uncond is not None,
                  uncond is not None and hint is not None and uncond["c_concat"] is None, masked, S, graph,
                  tuple(model.control_scales), model.only_mid_control, model.weights_fingerprint())

    def __eq__(self, other):
        return isinstance(other, _EngineKey) and self.t == other.t

    def __ne__(self, other):
        return not self.__eq__(other)


class _Engine:
    """Static buffers + one captured CUDA graph for a (model, shapes, S) combination."""

    def __init__(self, model, key, use_graph):
        self.model, self.key, self.use_graph = model, key, use_graph
        self.graph = None
        self.ready = False

    def _alloc(self, x_T, cond, uncond, S):
        dev = x_T.device
        b, c, h, w = x_T.shape
        self.dup = 2 if uncond is not None else 1
        nb = self.dup * b
        self.side_stream = torch.cuda.Stream(device=dev)
        self.x_lat = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        self.x_keep = torch.empty_like(self.x_lat)
        self.pred_x0 = torch.empty_like(self.x_lat)
        self.x_in = torch.empty((nb, h, w, 8), dtype=BF16, device=dev)
        self.ts_table = torch.zeros((S,), dtype=torch.int64, device=dev)
        self.coef = torch.zeros((S, 8), dtype=torch.float32, device=dev)
        self.step_ctr = torch.zeros((1,), dtype=torch.int32, device=dev)
        # eta > 0 (ddim_hacked.py:227-230): sigma_t * noise, one fresh tensor per step, read by the captured step graph as
        # row *step_ctr of this table
        self.noise = torch.zeros((S, b, c, h, w), dtype=torch.float32, device=dev)
        # inpainting (mask / x0, ddim_hacked.py:154-157): q_sample(x0, t_i) of every step, blended into the latent by the
        # first node of the step graph (row *step_ctr)
        self.masked = self.key.masked
        if self.masked:
            self.orig_table = torch.zeros((S, b, c, h, w), dtype=torch.float32, device=dev)
            self.mask = torch.zeros((b, c, h, w), dtype=torch.float32, device=dev)
        ctx_shape = cond["c_crossattn"][0].shape
        self.ctx = torch.empty((nb, ctx_shape[1], ctx_shape[2]), dtype=BF16, device=dev)
        self.has_hint = cond["c_concat"] is not None
        # guess mode: only the conditional rows go through the ControlNet
        self.guess = self.has_hint and uncond is not None and uncond["c_concat"] is None
        self.ctx_cn = self.ctx[:b] if self.guess else self.ctx   # (ONE view object: the hoisted K/V are keyed on its identity)
        if self.has_hint:
            hs = cond["c_concat"][0].shape
            self.hint = torch.empty((b if self.guess else nb, hs[1], hs[2], hs[3]), dtype=torch.float32, device=dev)

    def _load_inputs(self, x_T, cond, uncond, ts, rows, temperature=1.0, blend=None):
        b = x_T.shape[0]
        self.x_keep.copy_(x_T, non_blocking=True)
        if self.masked:
            mask, x0, model = blend
            self.mask.copy_(mask.expand_as(self.mask))
        for i, row in enumerate(rows):
            if self.masked:
                # (the step-by-step path draws q_sample's noise at the top of every step, before the step's own noise)
                t_i = torch.full((b,), int(ts[i]), device=x_T.device, dtype=torch.long)
                self.orig_table[i].copy_(model.q_sample(x0, t_i))
            if row[5] != 0.0:
                # the same draws, in the same order, as the step-by-step path (p_sample_ddim: one noise_like per step
                # whose sigma is not zero), so both paths produce the same trajectory from the same generator state
                self.noise[i].copy_(noise_like(tuple(self.noise[i].shape), self.noise.device, False) * temperature)
        self._ts_host = [int(v) for v in ts]
        self.ts_table.copy_(torch.tensor(ts, dtype=torch.int64), non_blocking=True)
        self.coef.copy_(torch.tensor(rows, dtype=torch.float32), non_blocking=True)
        conds = [cond] + ([uncond] if uncond is not None else [])
        for i, cd in enumerate(conds):
            ctx = cd["c_crossattn"][0] if len(cd["c_crossattn"]) == 1 else torch.cat(cd["c_crossattn"], 1)
            ctx = ctx.to(self.ctx.device)
            self.ctx[i * b:(i + 1) * b].copy_(ctx if ctx.dtype == BF16 else ops.to_bf16(ctx.float().contiguous()))
            if self.has_hint and cd["c_concat"] is not None:
                hint = cd["c_concat"][0] if len(cd["c_concat"]) == 1 else torch.cat(cd["c_concat"], 1)
                self.hint[i * b:(i + 1) * b].copy_(hint, non_blocking=True)
        self.reset_latent()

    def reset_latent(self):
        """Rewind to step 0 with the resident x_T: latent, step counter, and the bf16 NHWC network input
        (channels padded to 8, duplicated for cond/uncond). Device-side only."""
        b = self.x_lat.shape[0]
        self.x_lat.copy_(self.x_keep)
        ops.memset(self.step_ctr, 0)
        x0 = ops.nchw_to_nhwc(self.x_lat, 8)
        for i in range(self.dup):
            self.x_in[i * b:(i + 1) * b].copy_(x0)

    def _prologue(self):
        """Loop-invariant work of one image: time-embedding tables (eager, only when the timestep list changes), then
        the hint encoder and the cross-attention K/V projections -- replayed as ONE captured graph from the second image
        on (31 launches whose host-side cost, ~40 us each, would otherwise sit on every image's latency)."""
        self._prologue_emb()
        if self.use_graph and getattr(self, "pro_graph", None) is not None:
            self.pro_graph.replay()
            return
        self._prologue_cond()
        if self.use_graph and getattr(self, "pro_graph", None) is None and getattr(self, "_pro_warm", False):
            # second image: every buffer exists and every conv shape is tuned -> capture for the following images
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._prologue_cond()
            self.pro_graph = g
        self._pro_warm = True

    def _prologue_emb(self):
        """time-embedding tables: emb_layers outputs of every ResBlock for all S timesteps (depend only on t)."""
        from ..ldm.modules.diffusionmodules.openaimodel import StepEmb
        m = self.model
        unet = m.model.diffusion_model
        ts_key = tuple(self._ts_host)
        if getattr(self, "_emb_key", None) != ts_key:
            t_emb_all = ops.timestep_embedding(self.ts_table, self.S, unet.model_channels)
            self.emb_u = StepEmb.build(unet, t_emb_all, self.step_ctr, into=getattr(self, "emb_u", None))
            if self.has_hint:
                self.emb_c = StepEmb.build(m.control_model, t_emb_all, self.step_ctr, into=getattr(self, "emb_c", None))
            self._emb_key = ts_key

    def _prologue_cond(self):
        """Hint encoder and cross-attention K/V of every transformer block. Results live in buffers allocated once,
        because the captured graphs hold their addresses."""
        m = self.model
        if self.has_hint:
            g = m.control_model.run_hint(to_internal(self.hint))
            if getattr(self, "guided", None) is None:
                self.guided = g
            else:
                self.guided.copy_(g)
        else:
            self.guided = None
        from ..ldm.modules.attention import CrossAttention
        mods = [(mod, self.ctx) for mod in m.model.diffusion_model.modules()]
        if self.has_hint:
            mods += [(mod, self.ctx_cn) for mod in m.control_model.modules()]
        for mod, ctx in mods:
            if isinstance(mod, CrossAttention) and not mod.is_self:
                st = mod.kv_static
                if st is not None and st[0] is ctx:
                    mod.project_kv(ctx, k=st[1], vt=st[2])
                else:
                    k, vt, nkv, ldv = mod.project_kv(ctx)
                    mod.kv_static = (ctx, k, vt, nkv, ldv)

    def _step(self):
        m = self.model
        unet = m.model.diffusion_model
        nb = self.x_in.shape[0]
        b = self.x_lat.shape[0]
        if self.masked:
            ops.mask_blend_table_(self.x_lat, self.orig_table, self.mask, self.step_ctr)
            x_b = ops.nchw_to_nhwc(self.x_lat, 8)   # the network input of this step is the BLENDED latent
            for i in range(self.dup):
                self.x_in[i * b:(i + 1) * b].copy_(x_b)
        x = self.x_in.permute(0, 3, 1, 2)
        emb_u = self.emb_u  # per-image [S, Cout] tables, row selected on the device by step_ctr (see _prologue)
        if self.has_hint:
            # the UNet encoder and the ControlNet body are independent: run them on two streams (captured as two
            # branches of the step graph), then apply the 13 zero convs onto the UNet skips
            cn = m.control_model
            main = torch.cuda.current_stream()
            side = self.side_stream
            side.wait_stream(main)
            # while both branches are in flight every GEMM launch is limited to about half the SMs, so that the two
            # streams' kernels really run side by side (SDEO_BRANCH_CTAS, default 74; 0 = no limit)
            with ops.cta_budget(int(os.environ.get("SDEO_BRANCH_CTAS", "74"))):
                with torch.cuda.stream(side), ops.workspace_slot(1):
                    x_cn = self.x_in[:b].permute(0, 3, 1, 2) if self.guess else x
                    feats = cn.run_body(x_cn, self.guided, self.emb_c, self.ctx_cn)
                hs, h = unet.run_encoder(x, emb_u, self.ctx)
            # the 13 zero convs (+ control scale + add onto the UNet skips) stay on the side stream, launched in the
            # order the decoder consumes them; the decoder waits per tensor, so only the first ones are on its path
            side.wait_stream(main)
            n_out = len(hs) + 1
            if getattr(self, "_zc_events", None) is None:
                self._zc_events = [torch.cuda.Event() for _ in range(n_out)]
            evs = self._zc_events
            with torch.cuda.stream(side), ops.workspace_slot(1):
                outs = cn.run_zero_convs(feats, scales=m.control_scales, add_to=hs + [h], only_mid=m.only_mid_control,
                                         order=range(n_out - 1, -1, -1), after=lambda i: evs[i].record(side),
                                         rows=b if self.guess else None)
            main.wait_event(evs[n_out - 1])
            # the encoder outputs are read (as residuals) by the side stream: keep them alive until the step has been
            # enqueued, or main's allocator would hand their memory to the decoder while the zero convs still read it
            enc_keep = (hs, h)  # noqa: F841
            hs, h = outs[:-1], outs[-1]
            eps = nhwc(unet.run_decoder(h, hs, self.emb_u, self.ctx,
                                        before_block=lambda k: main.wait_event(evs[n_out - 2 - k])))
        else:
            hs, h = unet.run_encoder(x, emb_u, self.ctx)
            eps = nhwc(unet.run_decoder(h, hs, emb_u, self.ctx))          # fp32 [nb, h, w, 4]
        eps_c = eps[:b]
        eps_u = eps[b:] if self.dup == 2 else None
        ops.cfg_ddim_step(eps_c, eps_u, self.x_lat, self.coef, step_idx=self.step_ctr, noise_table=self.noise, x_prev=self.x_lat,
                          pred_x0=self.pred_x0, x_next=self.x_in, dup=self.dup, eps_nhwc=True)
        ops.counter_add(self.step_ctr, 1)

    def prepare(self, x_T, cond, uncond, ts, rows, temperature=1.0, blend=None):
        """Upload inputs, run the loop-invariant prologue and (first time) capture the per-step CUDA graph."""
        S = len(ts)
        if not self.ready:
            self._alloc(x_T, cond, uncond, S)
        self.S = S
        self._load_inputs(x_T, cond, uncond, ts, rows, temperature, blend)
        self._prologue()
        if self.use_graph and self.graph is None:
            # warm-up once eagerly (packs weights, sizes workspaces), then capture the step
            self._step()
            torch.cuda.synchronize()
            self.reset_latent()
            n0 = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
            self.launches_per_step = ops.LAUNCHES - n0
            self.graph = g
            self.reset_latent()
        self.ready = True

    def step(self):
        """One denoising step: graph replay (or the eager launch sequence)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()

    def run(self, x_T, cond, uncond, ts, rows, log_at=(), temperature=1.0, blend=None):
        """log_at: step numbers after which (x_{t-1}, pred_x0) are returned as copies (the reference's intermediates)."""
        self.prepare(x_T, cond, uncond, ts, rows, temperature, blend)
        log_at = set(log_at)
        logged = []
        for i in range(self.S):
            self.step()
            if i in log_at:
                logged.append((self.x_lat.clone(), self.pred_x0.clone()))
        return self.x_lat.clone(), self.pred_x0.clone(), logged
