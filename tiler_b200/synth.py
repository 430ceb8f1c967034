"""Seeded synthetic clips and tile helpers (the reference ships no generator and no source clips).

Frames are int32 images in the tile pixel format 0x00BBGGRR (utils.pas:243-246), i.e. after the loader's SwapRB
(tilingencoder.pas:1315).  Pure numpy: this is input generation, not part of the measured path.
"""
import numpy as np

SEED = 0x42381337  # the reference's own CRandomSeed (extern.pas:226)


def _value_noise(rng, h, w, octaves=3):
    out = np.zeros((h, w), dtype=np.float32)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        cell = max(2, 32 >> o)
        gh, gw = h // cell + 2, w // cell + 2
        g = rng.random((gh, gw), dtype=np.float32)
        ys, xs = np.arange(h) / cell, np.arange(w) / cell
        y0, x0 = ys.astype(int), xs.astype(int)
        fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
        a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
        out += amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy)
        tot += amp
        amp *= 0.5
    return out / tot


def make_clip(width, height, n_frames, cut_every=0, seed=SEED, n_sprites=12, noise=0.02):
    """-> uint8 [n_frames, height, width, 3] (R, G, B).  Per scene: a smooth colour-gradient canvas translating
    1-3 px/frame, textured sprites (3-octave value noise) on linear paths, `noise` fraction of uniformly random pixels;
    a hard scene cut every `cut_every` frames (0 = none)."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n_frames, height, width, 3), dtype=np.uint8)
    scene = None
    for f in range(n_frames):
        if scene is None or (cut_every and f % cut_every == 0):
            n_left = n_frames - f if not cut_every else min(cut_every, n_frames - f)
            vel = rng.integers(1, 4, size=2) * rng.choice([-1, 1], size=2)
            pad_y, pad_x = abs(int(vel[1])) * n_left + 1, abs(int(vel[0])) * n_left + 1
            ch, cw = height + pad_y, width + pad_x
            yy, xx = np.mgrid[0:ch, 0:cw].astype(np.float32)
            base = rng.random(3) * 255
            gx, gy = (rng.random(3) - 0.5) * 300 / width, (rng.random(3) - 0.5) * 300 / height
            canvas = np.empty((ch, cw, 3), dtype=np.float32)
            for c in range(3):
                canvas[..., c] = base[c] + gx[c] * xx + gy[c] * yy
            canvas += (_value_noise(rng, ch, cw, 2)[..., None] - 0.5) * 40.0
            canvas = np.mod(np.abs(canvas), 510.0)
            canvas = np.where(canvas > 255, 510.0 - canvas, canvas).astype(np.uint8)
            sprites = []
            for _ in range(n_sprites):
                sh, sw = int(rng.integers(16, 96)), int(rng.integers(16, 96))
                tex = (_value_noise(rng, sh, sw)[..., None] * (rng.random(3) * 255 + 40)).clip(0, 255).astype(np.uint8)
                pos = rng.random(2) * [height, width]
                sv = (rng.random(2) - 0.5) * 8
                sprites.append((tex, pos, sv))
            scene = (canvas, vel, sprites, f, pad_y, pad_x)
        canvas, vel, sprites, f0, pad_y, pad_x = scene
        t = f - f0
        oy = vel[1] * t if vel[1] > 0 else pad_y - 1 + vel[1] * t
        ox = vel[0] * t if vel[0] > 0 else pad_x - 1 + vel[0] * t
        out = canvas[oy:oy + height, ox:ox + width].copy()
        for tex, pos, sv in sprites:
            sh, sw = tex.shape[:2]
            y = int(pos[0] + sv[0] * t) % height
            x = int(pos[1] + sv[1] * t) % width
            y1, x1 = min(height, y + sh), min(width, x + sw)
            out[y:y1, x:x1] = tex[: y1 - y, : x1 - x]
        if noise > 0:
            n_noise = int(noise * height * width)
            pos = rng.integers(0, height * width, size=n_noise)
            out.reshape(-1, 3)[pos] = rng.integers(0, 256, size=(n_noise, 3), dtype=np.uint8)
        frames[f] = out
    return frames


def pack_rgb(frames_u8):
    """uint8 [..., 3] (R,G,B) -> int32 0x00BBGGRR."""
    f = frames_u8.astype(np.int32)
    return f[..., 0] | (f[..., 1] << 8) | (f[..., 2] << 16)


def frame_to_tiles(packed):
    """int32 [h, w] -> tiles [th*tw, 64]; tilemap rounds UP to whole tiles, missing pixels stay 0
    (tilingencoder.pas:1776, 1310)."""
    h, w = packed.shape
    th, tw = (h - 1) // 8 + 1, (w - 1) // 8 + 1
    buf = np.zeros((th * 8, tw * 8), dtype=np.int32)
    buf[:h, :w] = packed
    return buf.reshape(th, 8, tw, 8).transpose(0, 2, 1, 3).reshape(th * tw, 64)


def clip_to_tiles(frames_u8):
    """uint8 [n, h, w, 3] -> int32 [n, tiles_per_frame, 64]."""
    packed = pack_rgb(frames_u8)
    return np.stack([frame_to_tiles(p) for p in packed])


def tiles_to_frame(tiles, height, width):
    th, tw = (height - 1) // 8 + 1, (width - 1) // 8 + 1
    return tiles.reshape(th, tw, 8, 8).transpose(0, 2, 1, 3).reshape(th * 8, tw * 8)[:height, :width]


def random_features(n, seed, adversarial=False):
    """Synthetic int16[192] vectors: realistic-ish (Laplacian, decaying with frequency) or the SURVEY 8d adversarial set
    (uniform +-13000 on coefficient 0 of each plane... coefficient 0 only, Laplacian(b=40) elsewhere)."""
    rng = np.random.default_rng(seed)
    if adversarial:
        f = rng.laplace(0.0, 40.0, size=(n, 192))
        f[:, 0] = rng.uniform(-13000, 13000, size=n)
    else:
        scale = np.concatenate([2000.0 / (1.0 + np.arange(64)) ** 1.2] * 3)
        f = rng.laplace(0.0, 1.0, size=(n, 192)) * scale
        f[:, 0] += rng.uniform(0, 13000, size=n)
    return np.clip(np.rint(f), -32768, 32767).astype(np.int16)
