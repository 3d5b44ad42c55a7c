"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

numpy restatement of cv2.Canny(img, low, high) for 8-bit 1- or 3-channel images with the defaults the reference uses
(apertureSize=3, L2gradient=False): annotator/canny/__init__.py:4-6 calls cv2.Canny, canny2image_torch.py:33 feeds it the
resized HWC3 uint8 image. The algorithm lives in OpenCV (third-party; the reference pins opencv-contrib-python 4.3.0.36 in
requirements.txt, this image has cv2 4.13): modules/imgproc/src/canny.cpp -- Sobel 3x3 with BORDER_REPLICATE, per pixel the
channel with the largest L1 magnitude (strict >, first channel wins ties), non-maximum suppression with the fixed-point
tangent tests (CANNY_SHIFT 15, TG22 = 13573) on a zero-bordered magnitude map, double threshold, 8-connected hysteresis.

Pinned: tests/test_oracle.py checks it against cv2.Canny itself on random and structured images, and against the
committed fixture tests/golden/canny_bird0.npy (cv2.Canny of pictures_croped/bird_0.jpg, thresholds 100 / 200)."""
import numpy as np


def sobel_select(img):
    """img uint8 [H, W] or [H, W, C] -> (dx, dy, mag) int32 [H, W] of the max-L1 channel."""
    if img.ndim == 2:
        img = img[:, :, None]
    p = np.pad(img.astype(np.int32), ((1, 1), (1, 1), (0, 0)), mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    mag = np.abs(dx) + np.abs(dy)
    k = np.argmax(mag, axis=2)  # first maximum = strict '>' scan
    take = lambda a: np.take_along_axis(a, k[:, :, None], axis=2)[:, :, 0]
    return take(dx), take(dy), take(mag)


def canny(img, low_threshold, high_threshold):
    if low_threshold > high_threshold:
        low_threshold, high_threshold = high_threshold, low_threshold
    low, high = int(np.floor(low_threshold)), int(np.floor(high_threshold))
    dx, dy, mag = sobel_select(np.asarray(img, dtype=np.uint8))
    h, w = mag.shape
    mp = np.pad(mag, 1)  # zero border
    c = mp[1:-1, 1:-1]
    ax = np.abs(dx).astype(np.int64)
    ay = np.abs(dy).astype(np.int64) << 15
    tg22x = ax * 13573
    tg67x = tg22x + (ax << 16)
    horiz = ay < tg22x
    vert = (~horiz) & (ay > tg67x)
    diag = ~(horiz | vert)
    s_neg = (dx ^ dy) < 0  # s = -1
    keep_h = (c > mp[1:-1, :-2]) & (c >= mp[1:-1, 2:])
    keep_v = (c > mp[:-2, 1:-1]) & (c >= mp[2:, 1:-1])
    # s = +1: prev row at j-1, next row at j+1; s = -1: prev row at j+1, next row at j-1
    keep_dp = (c > mp[:-2, :-2]) & (c > mp[2:, 2:])
    keep_dn = (c > mp[:-2, 2:]) & (c > mp[2:, :-2])
    keep = (horiz & keep_h) | (vert & keep_v) | (diag & np.where(s_neg, keep_dn, keep_dp))
    cand = (c > low) & keep
    strong = cand & (c > high)
    weak = cand & ~strong
    # hysteresis: 8-connected growth of strong through weak
    edges = strong.copy()
    stack = list(zip(*np.nonzero(strong)))
    while stack:
        y, x = stack.pop()
        for yy in range(max(y - 1, 0), min(y + 2, h)):
            for xx in range(max(x - 1, 0), min(x + 2, w)):
                if weak[yy, xx] and not edges[yy, xx]:
                    edges[yy, xx] = True
                    stack.append((yy, xx))
    return (edges.astype(np.uint8)) * 255
