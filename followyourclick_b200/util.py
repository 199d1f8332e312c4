"""Output side of the clip: ``save_videos_grid`` with the reference's signature (animatediff/utils/util.py:18-30).

The reference moves the fp32 video (b, 3, F, H, W) to the host (50 MB at cfg2), tiles every frame with torchvision's make_grid,
converts to uint8 with numpy and hands the frames to imageio.  Here the tiling and the 8-bit conversion are one kernel on the device
(fyc_video_grid_u8), a quarter of the bytes crosses PCIe through a pinned buffer, and only the GIF container is written on the host
(PIL; imageio's writer is a PIL plugin too).  Frame pixels are identical to the reference's; the GIF palette quantisation is the
writer's business and is not part of the hot path."""
import os

import torch

from . import ops


@torch.no_grad()
def video_to_uint8_frames(videos, rescale=False, n_rows=6, device=None):
    """videos (b, 3, F, H, W) float in [0, 1] (device or host tensor) -> uint8 host array [F, Hg, Wg, 3] (pinned-memory backed)."""
    v = videos
    if not v.is_cuda:
        v = v.to(device or "cuda", non_blocking=True)       # a host tensor (the reference API's return type) goes through the device once
    grid = ops.video_grid_u8(v.to(torch.float32).contiguous(), nrow=n_rows, padding=2, rescale=rescale)
    host = torch.empty(grid.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(grid, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def save_videos_grid(videos, path, rescale=False, n_rows=6, fps=8):
    """Same call as the reference: writes an animated GIF of the clips tiled n_rows per row."""
    frames = video_to_uint8_frames(videos, rescale=rescale, n_rows=n_rows)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    from PIL import Image
    imgs = [Image.fromarray(f) for f in frames]
    imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=int(round(1000.0 / fps)), loop=0)
    return path
