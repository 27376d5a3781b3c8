"""Deterministic synthetic clips: textured global pan + an
independently moving rectangle + per-frame noise + one scene cut, YUV 4:2:0 / 4:2:2 /
4:4:4 / gray, 8 or 16 bit.  numpy only; used by tests/ and bench.py (there is no
network, so there are no real clips)."""
import numpy as np


def _box_blur(a, k=9):
    a = a.astype(np.float64)
    c = np.cumsum(np.pad(a, ((k // 2 + 1, k // 2), (0, 0)), mode="edge"), axis=0)
    a = (c[k:] - c[:-k]) / k
    c = np.cumsum(np.pad(a, ((0, 0), (k // 2 + 1, k // 2)), mode="edge"), axis=1)
    return (c[:, k:] - c[:, :-k]) / k


def _texture(seed, h, w):
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.integers(0, 256, size=(h + 256, w + 256)).astype(np.float64)
    t = _box_blur(_box_blur(t))
    t = (t - t.min()) / max(t.max() - t.min(), 1e-9)
    return (16 + t * (235 - 16)).astype(np.float64)


def make_clip(width, height, nframes, bits=8, ssw=1, ssh=1, gray=False, seed=0, scene_cut=True,
              noise_sigma=3.0, pure_noise=False):
    """Returns a list of frames; each frame is a list of planes (np.uint8 / np.uint16,
    C-contiguous).  `bits` in 8..16."""
    dtype = np.uint8 if bits <= 8 else np.uint16
    nplanes = 1 if gray else 3
    dims = [(height, width)] + [(height >> ssh, width >> ssw)] * (nplanes - 1)
    tex_seeds = [1234 + seed, 2345 + seed, 3456 + seed]
    cut_seeds = [9999 + seed, 9998 + seed, 9997 + seed]
    tex = [_texture(tex_seeds[p], *dims[p]) for p in range(nplanes)]
    tex2 = [_texture(cut_seeds[p], *dims[p]) for p in range(nplanes)] if scene_cut else None
    frames = []
    for t in range(nframes):
        planes = []
        use2 = scene_cut and t >= nframes // 2 and nframes >= 4
        for p in range(nplanes):
            h, w = dims[p]
            sx = 1 if p == 0 else (1 << ssw)
            sy = 1 if p == 0 else (1 << ssh)
            if pure_noise:
                rng = np.random.Generator(np.random.PCG64(42 + 131 * t + p + seed))
                img = rng.integers(0, 256, size=(h, w)).astype(np.float64)
            else:
                T = (tex2 if use2 else tex)[p]
                ox = 128 + int(round(3 * t * np.cos(0.1 * t))) // sx
                oy = 128 + int(round(2 * t * np.sin(0.07 * t))) // sy
                ox = min(max(ox, 0), 256)
                oy = min(max(oy, 0), 256)
                img = T[oy:oy + h, ox:ox + w].copy()
                # second motion: a rectangle (25 % area) moving (-2t, +t) luma pixels
                rh, rw = h // 2, w // 2
                ry = (h // 4 + t // sy) % max(h - rh, 1)
                rx = (w // 4 - (2 * t) // sx) % max(w - rw, 1)
                T2 = (tex if use2 or tex2 is None else tex2)[p]
                img[ry:ry + rh, rx:rx + rw] = 255 - T2[64:64 + rh, 64:64 + rw] * 0.8
                rng = np.random.Generator(np.random.PCG64(5678 + t * 7 + p + seed))
                img += rng.normal(0.0, noise_sigma, size=img.shape)
            img8 = np.clip(np.rint(img), 0, 255).astype(np.uint16)
            if bits <= 8:
                planes.append(np.ascontiguousarray(img8.astype(np.uint8)))
            else:
                rng = np.random.Generator(np.random.PCG64(777 + t * 5 + p + seed))
                low = rng.integers(0, 1 << (bits - 8), size=img8.shape).astype(np.uint16)
                v = (img8 << (bits - 8)) | low
                planes.append(np.ascontiguousarray(v.astype(dtype)))
        frames.append(planes)
    return frames
