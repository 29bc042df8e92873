"""Deterministic synthetic frame pairs: analytic textures displaced by an
analytic flow, so frame 2 needs no interpolation.  Used by the tests and bench.py only."""
import numpy as np


def texture_params(seed, n=64):
    rng = np.random.default_rng(seed)
    wavelength = np.exp(rng.uniform(np.log(6.0), np.log(96.0), n))
    theta = rng.uniform(0, 2 * np.pi, n)
    fx = np.cos(theta) / wavelength
    fy = np.sin(theta) / wavelength
    amp = 1.0 / np.sqrt(np.arange(1, n + 1))
    phase = rng.uniform(0, 2 * np.pi, n)
    return fx, fy, amp, phase


def texture(x, y, tp, contrast=1.0):
    fx, fy, amp, phase = tp
    acc = np.zeros_like(x, dtype=np.float64)
    for k in range(len(fx)):
        acc += amp[k] * np.sin(2 * np.pi * (fx[k] * x + fy[k] * y) + phase[k])
    scale = 119.5 / (2.2 * np.sqrt((amp ** 2).sum() / 2))  # ~[8, 247] for unit contrast
    return 127.5 + contrast * scale * acc


def flow_field(x, y, U0, U1, L, seed):
    rng = np.random.default_rng(seed + 7919)
    q = rng.uniform(0, 2 * np.pi, 4)
    u = U0[0] + U1 * np.sin(2 * np.pi * x / L + q[0]) * np.cos(2 * np.pi * y / L + q[1])
    v = U0[1] + U1 * np.sin(2 * np.pi * x / L + q[2]) * np.cos(2 * np.pi * y / L + q[3])
    return u, v


def make_pair(w, h, seed, U0=(0.3, -0.2), U1=0.5, L=256.0, contrast=1.0, noise=0.0):
    """Returns (frame0, frame1, u_true, v_true) as float32 (h, w) arrays; frame1(x) = I(x - flow(x))."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    tp = texture_params(seed)
    u, v = flow_field(x, y, U0, U1, L, seed)
    f0 = texture(x, y, tp, contrast)
    f1 = texture(x - u, y - v, tp, contrast)
    if noise > 0:
        rng = np.random.default_rng(seed + 104729)
        f0 = f0 + rng.normal(0, noise, f0.shape)
        f1 = f1 + rng.normal(0, noise, f1.shape)
    return f0.astype(np.float32), f1.astype(np.float32), u.astype(np.float32), v.astype(np.float32)


def smooth_random(w, h, seed, lo=-1.0, hi=1.0, cells=6):
    """Smooth random field (bilinear upsampling of a coarse grid): test flows / increments."""
    rng = np.random.default_rng(seed)
    g = rng.uniform(lo, hi, (cells + 1, cells + 1))
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    gx, gy = x * cells / max(w - 1, 1), y * cells / max(h - 1, 1)
    x0, y0 = np.minimum(gx.astype(int), cells - 1), np.minimum(gy.astype(int), cells - 1)
    ax, ay = gx - x0, gy - y0
    out = (g[y0, x0] * (1 - ax) * (1 - ay) + g[y0, x0 + 1] * ax * (1 - ay) +
           g[y0 + 1, x0] * (1 - ax) * ay + g[y0 + 1, x0 + 1] * ax * ay)
    return out.astype(np.float32)


def make_pair_torch(w, h, seed, device, U0=(0.3, -0.2), U1=0.5, L=256.0, contrast=1.0, n=24):
    """Same construction as make_pair, evaluated with torch on `device` (for frames too large for numpy
    in reasonable time, e.g. 8192x8192).  Returns float32 torch tensors (frame0, frame1) on `device`."""
    import torch
    fx, fy, amp, phase = texture_params(seed, n)
    ys = torch.arange(h, dtype=torch.float64, device=device).view(h, 1)
    xs = torch.arange(w, dtype=torch.float64, device=device).view(1, w)
    rng = np.random.default_rng(seed + 7919)
    q = rng.uniform(0, 2 * np.pi, 4)
    u = U0[0] + U1 * torch.sin(2 * np.pi * xs / L + q[0]) * torch.cos(2 * np.pi * ys / L + q[1])
    v = U0[1] + U1 * torch.sin(2 * np.pi * xs / L + q[2]) * torch.cos(2 * np.pi * ys / L + q[3])
    scale = 119.5 / (2.2 * np.sqrt((amp ** 2).sum() / 2))

    def tex(x, y):
        acc = torch.zeros((h, w), dtype=torch.float64, device=device)
        for k in range(n):
            acc += amp[k] * torch.sin(2 * np.pi * (fx[k] * x + fy[k] * y) + phase[k])
        return (127.5 + contrast * scale * acc).to(torch.float32)
    f0 = tex(xs.expand(h, w), ys.expand(h, w))
    f1 = tex(xs - u, ys - v)
    return f0, f1
