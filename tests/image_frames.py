"""Procedural 8-bit grayscale test frames (deterministic across machines: numpy's legacy RandomState), shared by
oracle/make_image_golden.py and the image-pipeline tests so that no image files need to be committed."""
import numpy as np


def make_frame(H, W, seed):
    rs = np.random.RandomState(1000 + seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = 128 + 60 * np.sin(xx / (11.0 + seed)) * np.cos(yy / (17.0 + 2 * seed)) + 30 * ((xx // 37 + yy // 29) % 2)
    img += rs.randn(H, W) * 20
    cy, cx, r = H * 0.4, W * 0.55, min(H, W) * 0.2
    img[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] += 45          # a bright disc: sharp edges exercise the negative lobes
    return np.clip(img, 0, 255).astype(np.uint8)
