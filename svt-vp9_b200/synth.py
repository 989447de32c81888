"""Deterministic synthetic YUV420 8-bit clips.

Same bytes feed the reference encoder (oracle/_ref) and the B200 path, so parity checks and
benchmarks are reproducible without any dataset.  texture = uniform u8 noise, k passes of a
5-point box blur, contrast stretch; frame n = crop of the texture at a linearly moving offset
(+ optional uniform integer noise), U = half-res texture crop, V = 255 - U.
"""
import numpy as np

CONFIGS = {
    # name: (width, height, seed, k, g, a, b, m, noise, distinct_frames)
    "A": (640, 480, 1234, 3, 3, 3, 2, 32, 0, 60),
    "B": (1920, 1080, 7, 4, 4, 3, 2, 64, 3, 40),
    "C": (3840, 2160, 11, 5, 5, 5, 3, 96, 3, 24),
}


def _texture(h, w, rng, k, g):
    t = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32)
    for _ in range(k):
        t = (t + np.roll(t, 1, 0) + np.roll(t, -1, 0) + np.roll(t, 1, 1) + np.roll(t, -1, 1)) / 5.0
    t = (t - t.mean()) * g + 128.0
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def make_clip(width, height, n_frames, seed=7, k=4, g=4, a=3, b=2, m=64, noise=3):
    """Return a list of (Y, U, V) uint8 arrays (planar 4:2:0)."""
    rng = np.random.default_rng(seed)
    tex = _texture(height + 2 * m + 32, width + 2 * m + 32, rng, k, g)
    texc = tex[::2, ::2]
    frames = []
    for n in range(n_frames):
        dx, dy = (a * n) % m, (b * n) % m
        y = tex[dy:dy + height, dx:dx + width].astype(np.int16)
        if noise:
            y = y + rng.integers(-noise, noise + 1, size=y.shape, dtype=np.int16)
        y = np.clip(y, 0, 255).astype(np.uint8)
        u = np.ascontiguousarray(texc[dy // 2:dy // 2 + height // 2, dx // 2:dx // 2 + width // 2])
        v = (255 - u).astype(np.uint8)
        frames.append((np.ascontiguousarray(y), u, v))
    return frames


def make_config(name, n_frames=None):
    w, h, seed, k, g, a, b, m, noise, distinct = CONFIGS[name]
    n = distinct if n_frames is None else min(n_frames, distinct)
    return w, h, make_clip(w, h, n, seed, k, g, a, b, m, noise)


def write_yuv(path, frames):
    with open(path, "wb") as f:
        for y, u, v in frames:
            f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())


if __name__ == "__main__":
    import sys
    name, path = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else None
    w, h, fr = make_config(name, n)
    write_yuv(path, fr)
    print(w, h, len(fr))
