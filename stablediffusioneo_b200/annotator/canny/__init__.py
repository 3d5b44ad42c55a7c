"""CannyDetector with the reference's interface (annotator/canny/__init__.py:4-6), running on the device
(csrc/canny.cu, bit-exact with cv2.Canny for uint8 1-/3-channel images, aperture 3, L1 gradient).
numpy in -> numpy out (like cv2); a CUDA uint8 tensor in -> a CUDA uint8 tensor out (no host round trip)."""
import numpy as np
import torch

from ... import ops


class CannyDetector:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def __call__(self, img, low_threshold, high_threshold):
        if torch.is_tensor(img):
            return ops.canny(img, low_threshold, high_threshold)
        x = torch.from_numpy(np.ascontiguousarray(img)).to(self.device, non_blocking=True)
        return ops.canny(x, low_threshold, high_threshold).cpu().numpy()

    def hint(self, img, low_threshold, high_threshold, num_samples=1):
        """Image (numpy or CUDA uint8 HWC) -> (control hint fp32 [num_samples, 3, H, W] in {0, 1} on the device,
        detected_map uint8 [H, W] on the device): canny2image_torch.py:33-38 without leaving the GPU."""
        x = img if torch.is_tensor(img) else torch.from_numpy(np.ascontiguousarray(img)).to(self.device, non_blocking=True)
        edges = ops.canny(x, low_threshold, high_threshold)
        return ops.edges_to_hint(edges, num_samples), edges
