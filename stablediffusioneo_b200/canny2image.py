"""The reference's image pipeline object (canny2image_torch.py:17-70): `hackathon().initialize()` then
`process(input_image, prompt, a_prompt, n_prompt, num_samples, image_resolution, ddim_steps, guess_mode, strength, scale,
seed, eta, low_threshold, high_threshold) -> [uint8 HWC image] * num_samples`, the call compute_score_torch.py:22-40 times
per image. Every stage after the host-side resize runs on the device: Canny + HWC3/255 hint (csrc/canny.cu), CLIP text
encoder (when built), the 20-step ControlNet+UNet engine (one CUDA graph per step), VAE decode and the uint8 image.

Differences that the environment forces, not the design: the reference reads ./models/cldm_v15.yaml and
control_sd15_canny.pth, neither of which ships with the reference checkout -- initialize() therefore takes optional paths
(or a ready ControlLDM); prompts may be strings (needs the CLIP tokenizer files on disk), token ids [B, 77], or context
tensors [B, 77, 768]."""
import random

import numpy as np
import torch

from .annotator.canny import CannyDetector
from .annotator.util import HWC3, resize_image
from .cldm.ddim_hacked import DDIMSampler
from .cldm.model import create_model, load_state_dict


def seed_everything(seed):
    """pytorch_lightning.seed_everything (canny2image_torch.py:42): python, numpy, torch (CPU + CUDA) generators."""
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


class hackathon:
    def initialize(self, config_path=None, ckpt_path=None, device="cuda", model=None, with_text_encoder=False):
        self.device = torch.device(device)
        self.apply_canny = CannyDetector(self.device)
        if model is None:
            kw = {"cond_stage_config": {}} if (with_text_encoder and config_path is None) else {}
            model = create_model(config_path, device=self.device, with_text_encoder=with_text_encoder, **kw)
            if ckpt_path is not None:
                # the reference loads strictly (canny2image_torch.py:20); here the checkpoint may carry keys this object
                # does not hold (cond_stage_model.* without a text encoder, the VAE encoder, position_ids), but every
                # parameter of the model must be in it -- a partial checkpoint would leave random / zero-conv weights
                res = model.load_state_dict(load_state_dict(ckpt_path, location=str(self.device)), strict=False)
                if res.missing_keys:
                    raise RuntimeError(f"checkpoint {ckpt_path} lacks {len(res.missing_keys)} parameters of the model, e.g. "
                                       f"{res.missing_keys[:5]}")
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)  # ops launch on the current device's stream
        self.model = model.to(self.device).eval()
        self.ddim_sampler = DDIMSampler(self.model)

    def _conditioning(self, prompt, num_samples):
        """strings / token ids -> get_learned_conditioning; a float tensor [B or 1, 77, C] is used as the context itself."""
        if torch.is_tensor(prompt) and prompt.is_floating_point():
            ctx = prompt.to(self.device)
            return ctx.expand(num_samples, -1, -1).contiguous() if ctx.shape[0] == 1 and num_samples > 1 else ctx
        if torch.is_tensor(prompt):
            ids = prompt.to(self.device)
            ids = ids.expand(num_samples, -1) if ids.shape[0] == 1 and num_samples > 1 else ids
            return self.model.get_learned_conditioning(ids)
        return self.model.get_learned_conditioning([prompt] * num_samples)

    @torch.no_grad()
    def process(self, input_image, prompt, a_prompt, n_prompt, num_samples, image_resolution, ddim_steps, guess_mode, strength,
                scale, seed, eta, low_threshold, high_threshold):
        img = resize_image(HWC3(input_image), image_resolution)
        H, W, C = img.shape
        control, detected_map = self.apply_canny.hint(img, low_threshold, high_threshold, num_samples)
        self.detected_map = detected_map
        if seed == -1:
            seed = random.randint(0, 65535)
        seed_everything(seed)
        positive = prompt + ', ' + a_prompt if isinstance(prompt, str) else prompt  # tensors carry the full prompt
        cond = {"c_concat": [control], "c_crossattn": [self._conditioning(positive, num_samples)]}
        un_cond = {"c_concat": None if guess_mode else [control], "c_crossattn": [self._conditioning(n_prompt, num_samples)]}
        shape = (4, H // 8, W // 8)
        self.model.control_scales = ([strength * (0.825 ** float(12 - i)) for i in range(13)] if guess_mode
                                     else [strength] * 13)
        samples, _ = self.ddim_sampler.sample(ddim_steps, num_samples, shape, cond, verbose=False, eta=eta,
                                              unconditional_guidance_scale=scale, unconditional_conditioning=un_cond)
        x_samples = self.model.decode_first_stage_u8(samples).cpu().numpy()  # 'b h w c' uint8 (canny2image_torch.py:68)
        return [x_samples[i] for i in range(num_samples)]
