"""annotator/util.py mirrors (annotator/util.py:9-38): HWC3 and resize_image on uint8 HWC images (host side, numpy).
resize_image needs OpenCV only when the size actually changes (Lanczos / area resampling are cv2's)."""
import numpy as np


def HWC3(x):
    assert x.dtype == np.uint8
    if x.ndim == 2:
        x = x[:, :, None]
    assert x.ndim == 3
    H, W, C = x.shape
    assert C == 1 or C == 3 or C == 4
    if C == 3:
        return x
    if C == 1:
        return np.concatenate([x, x, x], axis=2)
    color = x[:, :, 0:3].astype(np.float32)
    alpha = x[:, :, 3:4].astype(np.float32) / 255.0
    y = color * alpha + 255.0 * (1.0 - alpha)
    return y.clip(0, 255).astype(np.uint8)


def resize_image(input_image, resolution):
    H, W, C = input_image.shape
    k = float(resolution) / min(float(H), float(W))
    Hn = int(np.round(H * k / 64.0)) * 64
    Wn = int(np.round(W * k / 64.0)) * 64
    if (Hn, Wn) == (H, W):
        return input_image
    import cv2
    return cv2.resize(input_image, (Wn, Hn), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)
