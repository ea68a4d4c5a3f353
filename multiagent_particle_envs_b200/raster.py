"""Headless rasteriser for `MultiAgentEnv.render('rgb_array')`.

The reference draws with pyglet/OpenGL (multiagent/rendering.py, environment.py:200-263): filled circles of
radius `entity.size` in `entity.color` (agents half transparent), camera = [-1, 1]^2 around the origin (or around
agent i when `shared_viewer=False`), 700 x 700 pixels.  This is a NumPy restatement of that picture for ONE
world -- a debugging aid, never on the hot path (positions are read back from the device).
"""
import numpy as np


def draw_world(positions, sizes, colors, alphas, center=(0.0, 0.0), cam_range=1.0, pixels=700):
    """positions [E,2], sizes [E], colors [E,3] in 0..1, alphas [E] -> uint8 image [pixels, pixels, 3]
    (white background, y axis pointing up, later entities drawn over earlier ones)."""
    img = np.ones((pixels, pixels, 3), dtype=np.float32)
    xs = center[0] + (np.arange(pixels, dtype=np.float32) + 0.5) / pixels * 2 * cam_range - cam_range
    ys = center[1] + cam_range - (np.arange(pixels, dtype=np.float32) + 0.5) / pixels * 2 * cam_range
    X, Y = np.meshgrid(xs, ys)
    for p, s, c, al in zip(positions, sizes, colors, alphas):
        if c is None or not np.all(np.isfinite(p)):
            continue
        mask = (X - p[0]) ** 2 + (Y - p[1]) ** 2 <= s * s
        col = np.asarray(c, dtype=np.float32)[:3]
        img[mask] = (1.0 - al) * img[mask] + al * col
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
