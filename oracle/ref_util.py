"""Oracle: ``save_videos_grid``'s frame preparation (animatediff/utils/util.py:18-27) without the GIF writer.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``make_grid`` restates torchvision.utils.make_grid (the dependency the reference
calls; torchvision 0.26 here, its grid algorithm is unchanged since 0.2) for (b, 3, H, W) inputs with the reference's arguments
(nrow = n_rows, padding 2, pad_value 0, no normalisation); tests/test_oracle_golden.py pins it against the installed torchvision.
"""
import math

import numpy as np
import torch


def make_grid(x, nrow, padding=2):
    b, c, h, w = x.shape
    if b == 1:
        return x[0]
    xmaps = min(nrow, b)
    ymaps = int(math.ceil(float(b) / xmaps))
    height, width = h + padding, w + padding
    grid = x.new_full((c, height * ymaps + padding, width * xmaps + padding), 0.0)
    k = 0
    for yy in range(ymaps):
        for xx in range(xmaps):
            if k >= b:
                break
            grid[:, yy * height + padding: yy * height + padding + h, xx * width + padding: xx * width + padding + w] = x[k]
            k += 1
    return grid


def video_frames_uint8(videos, rescale=False, n_rows=6):
    """util.py:19-27: (b, c, t, h, w) -> list over t of uint8 (Hg, Wg, c) arrays."""
    out = []
    for x in videos.permute(2, 0, 1, 3, 4):
        x = make_grid(x, n_rows).transpose(0, 1).transpose(1, 2)
        if rescale:
            x = (x + 1.0) / 2.0
        out.append((x * 255).numpy().astype(np.uint8))
    return out
