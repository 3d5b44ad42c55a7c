"""sdeo — B200-native ControlNet-SD1.5 denoising engine (hand-written sm_100a CUDA behind a C ABI).

Package layout mirrors the reference's operator surface for the denoising hot path:
  stablediffusioneo_b200.ldm.modules.attention            <- ldm/modules/attention.py
  stablediffusioneo_b200.ldm.modules.diffusionmodules.*   <- ldm/modules/diffusionmodules/{openaimodel,util,model}.py
  stablediffusioneo_b200.cldm.cldm / .cldm.ddim_hacked    <- cldm/cldm.py, cldm/ddim_hacked.py
"""
__version__ = "0.1.0"
