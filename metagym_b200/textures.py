"""Texture sets for the MetaMaze raycaster.

The reference loads nine 64x64 PNGs at import (metagym/metamaze/envs/maze_task.py:19-35): index 0 = ground, 1..6 = wall
textures, plus one ceiling image; `pygame.surfarray.array3d` yields x-major (W, H, 3) uint8 arrays that it keeps as
float32.  The renderer only ever indexes `texture[int(u * W), int(v * H)]`, so any [n_tex, 64, 64, 3] uint8 stack works.

`synthetic_textures` builds a deterministic procedural set (stone / brick / moss style value noise) so that the engine,
its tests and its benchmark do not depend on the reference's image files; `load_texture_dir` reads a directory laid out
like the reference's `img/` folder when a user wants the original look.
"""
import os

import numpy as np


def _value_noise(rs, size, cells):
    g = rs.rand(cells + 1, cells + 1)
    x = np.linspace(0, cells, size, endpoint=False)
    i = x.astype(int)
    f = x - i
    f = f * f * (3 - 2 * f)
    a = g[i][:, i]
    b = g[i + 1][:, i]
    c = g[i][:, i + 1]
    d = g[i + 1][:, i + 1]
    fx = f[:, None]
    fy = f[None, :]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synthetic_textures(n_tex=7, size=64, seed=0):
    """-> (grounds uint8 [n_tex, size, size, 3], ceil uint8 [size, size, 3])."""
    rs = np.random.RandomState(seed)
    palettes = [(120, 120, 120), (70, 110, 160), (150, 90, 60), (90, 130, 80), (170, 60, 50), (160, 140, 90),
                (110, 80, 140), (60, 60, 70), (140, 120, 100)]
    out = []
    for k in range(n_tex + 1):
        base = np.array(palettes[k % len(palettes)], dtype=np.float64)
        n1 = _value_noise(rs, size, 4)
        n2 = _value_noise(rs, size, 16)
        lum = 0.55 + 0.35 * n1 + 0.25 * (n2 - 0.5)
        img = lum[:, :, None] * base[None, None, :]
        if k not in (0, n_tex):                      # wall textures get mortar lines (brick courses)
            rows = (np.arange(size) // 8) % 2
            img[:, np.arange(size) % 8 == 0, :] *= 0.55
            for r in range(size // 8):
                off = 0 if r % 2 == 0 else 8
                cols = (np.arange(size) + off) % 16 == 0
                img[np.ix_(cols, np.arange(r * 8, r * 8 + 8))] *= 0.6
            del rows
        img += rs.randint(-6, 7, size=img.shape)
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return np.stack(out[:n_tex]), out[n_tex]


def load_texture_dir(texture_dir):
    """maze_task.py:19-35: sorted file names; 'ground' -> index 0, every 'wall' appended, 'ceil' -> ceiling."""
    from PIL import Image

    def load(path):
        a = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
        return np.ascontiguousarray(a.transpose(1, 0, 2))        # pygame surfarray is x-major

    grounds = [None]
    ceil = None
    for name in sorted(os.listdir(texture_dir)):
        path = os.path.join(texture_dir, name)
        if name.find("wall") >= 0:
            grounds.append(load(path))
        if name.find("ground") >= 0:
            grounds[0] = load(path)
        if name.find("ceil") >= 0:
            ceil = load(path)
    if grounds[0] is None or ceil is None:
        raise ValueError("texture directory needs a *ground* and a *ceil* image")
    return np.stack(grounds), ceil
