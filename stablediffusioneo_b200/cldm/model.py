"""Checkpoint / config helpers with the reference's names (cldm/model.py:8-28): get_state_dict, load_state_dict,
create_model. The SD checkpoint prefixes are the module tree's own (`control_model.*`, `model.diffusion_model.*`,
`first_stage_model.*`, `cond_stage_model.transformer.*`), so `model.load_state_dict(load_state_dict(path), strict=False)`
is the whole loading story, as in canny2image_torch.py:20-22. OmegaConf is not required: the YAML is read with PyYAML."""
import os

import torch


def get_state_dict(d):
    return d.get('state_dict', d)


def load_state_dict(ckpt_path, location='cpu'):
    """.safetensors through safetensors.torch, anything else through torch.load; unwraps a 'state_dict' key."""
    _, extension = os.path.splitext(ckpt_path)
    if extension.lower() == ".safetensors":
        import safetensors.torch
        state_dict = safetensors.torch.load_file(ckpt_path, device=location)
    else:
        state_dict = get_state_dict(torch.load(ckpt_path, map_location=torch.device(location), weights_only=True))
    state_dict = get_state_dict(state_dict)
    print(f'Loaded state_dict from [{ckpt_path}]')
    return state_dict


def _params(node):
    return dict((node or {}).get("params", {}) or {})


def create_model(config_path=None, device=None, **overrides):
    """ControlLDM from a cldm_v15.yaml-style config (`model.params.{unet_config, control_stage_config, first_stage_config,
    cond_stage_config}.params`, `scale_factor`, `linear_start/end`, `timesteps`, `only_mid_control`, `control_key`);
    config_path None = the SD1.5 defaults recovered in SURVEY 8 a-0 (the yaml itself is absent from the reference
    checkout). `target:` class paths are ignored -- the classes are this package's mirrors. The text encoder is built only
    when the config names a cond_stage_config AND `with_text_encoder=True` is passed (its tokenizer files are external)."""
    from .cldm import ControlLDM
    kw = {}
    with_text = overrides.pop("with_text_encoder", False)
    if config_path is not None:
        import yaml
        with open(config_path) as f:
            cfg = yaml.safe_load(f)
        p = _params(cfg.get("model"))
        if "unet_config" in p:
            kw["unet_config"] = _params(p["unet_config"])
        if "control_stage_config" in p:
            kw["control_stage_config"] = _params(p["control_stage_config"])
        if "first_stage_config" in p:
            fs = _params(p["first_stage_config"])
            kw["first_stage_config"] = dict(fs.get("ddconfig", fs))
        if with_text and "cond_stage_config" in p:
            kw["cond_stage_config"] = _params(p["cond_stage_config"])
        for name in ("control_key", "only_mid_control", "timesteps", "linear_start", "linear_end", "scale_factor",
                     "parameterization"):
            if name in p:
                kw[name] = p[name]
        print(f'Loaded model config from [{config_path}]')
    kw.update(overrides)
    if device is not None:
        with torch.device(device):
            return ControlLDM(**kw)
    return ControlLDM(**kw)
