"""Synthetic Blender-style scene and ray batches.

Host-side data generation for tests and benchmarks: pinhole cameras on a radius-4 sphere looking at the
origin with the ray formulas of ScratchNerf/ScratchNerf/Dataset.cs:111-176 (direction, origin, radius), near=2,
far=6, lossMult=1, white background (SN/TrainState.cs:67-71); target colours come from an analytic
emission-absorption scene (Gaussian blobs, fp64 quadrature) so they lie in [0,1].  Also writes / reads the
64-byte ``train_data.bin`` record of SN/BinDataset.cs:35-49.
"""
from __future__ import annotations

import numpy as np

def view_pose(view=0, width=800, n_views=8, scene_seed=1234):
    """(c2w [3,4] float32 row-major [R | t], focal) of camera `view` of the synthetic scene: what
    `nerf_mipnerf_render_view` takes to render that view with the rays generated on the device."""
    rng = np.random.default_rng(scene_seed)
    focal = 0.5 * width / np.tan(0.5 * 0.6911112070083618)
    th = rng.uniform(0, 2 * np.pi, n_views)
    ph = rng.uniform(0.15 * np.pi, 0.5 * np.pi, n_views)
    cam = 4.0 * np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], -1)[view]
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    c2w = np.concatenate([np.stack([right, upv, -fwd], -1), cam[:, None]], 1)
    return c2w.astype(np.float32), float(focal)


def synthetic_rays(n_rays, width=800, height=800, n_views=8, seed=2024, scene_seed=1234, near=2.0, far=6.0):
    """Blender-style synthetic batch: pinhole cameras on a radius-4 sphere looking at the origin,
    ray formulas of SN/Dataset.cs:111-176, analytic blob colours in [0,1], white background.
    Returns (rays dict, pixels[R,3]) as float32.  Deterministic in (seed, scene_seed)."""
    rng = np.random.default_rng(scene_seed)
    focal = 0.5 * width / np.tan(0.5 * 0.6911112070083618)
    th = rng.uniform(0, 2 * np.pi, n_views)
    ph = rng.uniform(0.15 * np.pi, 0.5 * np.pi, n_views)
    cam = 4.0 * np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], -1)
    fwd = -cam / np.linalg.norm(cam, axis=-1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    upv = np.cross(right, fwd)
    c2w = np.stack([right, upv, -fwd], -1)  # columns: x, y, z(back)
    blobs_c = rng.uniform(-0.8, 0.8, (5, 3))
    blobs_s = rng.uniform(0.25, 0.5, 5)
    blobs_rgb = rng.uniform(0.1, 0.9, (5, 3))

    rb = np.random.default_rng(seed)
    vi = rb.integers(0, n_views, n_rays)
    px = rb.integers(0, width, n_rays)
    py = rb.integers(0, height, n_rays)

    def cam_dir(x, y):
        return np.stack([(x - width * 0.5 + 0.5) / focal, -(y - height * 0.5 + 0.5) / focal, -np.ones_like(x, dtype=np.float64)], -1)

    d0 = np.einsum("rij,rj->ri", c2w[vi], cam_dir(px.astype(np.float64), py.astype(np.float64)))
    nx = np.minimum(px + 1, width - 1).astype(np.float64)
    d1 = np.einsum("rij,rj->ri", c2w[vi], cam_dir(nx, py.astype(np.float64)))
    radii = np.linalg.norm(d0 - d1, axis=-1) * 2 / np.sqrt(12.0)
    edge = px == width - 1
    if edge.any():  # SN/Dataset.cs:151 gives 0 at the last column; use the left neighbour instead
        d2 = np.einsum("rij,rj->ri", c2w[vi], cam_dir(px.astype(np.float64) - 1, py.astype(np.float64)))
        radii = np.where(edge, np.linalg.norm(d0 - d2, axis=-1) * 2 / np.sqrt(12.0), radii)
    o = cam[vi]
    # analytic colour: emission-absorption through Gaussian blobs, 64-point quadrature in fp64
    ts = np.linspace(near, far, 65)
    tm = 0.5 * (ts[1:] + ts[:-1])
    pts = o[:, None, :] + d0[:, None, :] * tm[None, :, None]
    dens = np.zeros(pts.shape[:2])
    col = np.zeros(pts.shape)
    for c, s, k in zip(blobs_c, blobs_s, blobs_rgb):
        g = 8.0 * np.exp(-0.5 * np.sum((pts - c) ** 2, -1) / s**2)
        dens += g
        col += g[..., None] * k
    col = col / np.maximum(dens[..., None], 1e-12)
    dl = np.linalg.norm(d0, axis=-1, keepdims=True) * (ts[1] - ts[0])
    alpha = 1 - np.exp(-dens * dl)
    T = np.cumprod(np.concatenate([np.ones((n_rays, 1)), 1 - alpha[:, :-1]], 1), 1)
    w = alpha * T
    pix = np.sum(w[..., None] * col, 1) + (1 - w.sum(1, keepdims=True))
    rays = dict(origins=o.astype(np.float32), directions=d0.astype(np.float32), radii=radii.astype(np.float32),
                nears=np.full(n_rays, near, np.float32), fars=np.full(n_rays, far, np.float32),
                loss_mults=np.ones(n_rays, np.float32))
    return rays, np.clip(pix, 0, 1).astype(np.float32)


RECORD_FLOATS = 16  # o(3) d(3) viewdir(3) radius near far lossmult rgb(3)  — SN/BinDataset.cs:40-49


def pack_records(rays, pixels) -> np.ndarray:
    """[R,16] float32 records in the train_data.bin layout."""
    d = rays["directions"]
    view = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return np.concatenate([rays["origins"], d, view, rays["radii"][:, None], rays["nears"][:, None],
                           rays["fars"][:, None], rays["loss_mults"][:, None], pixels], 1).astype(np.float32)


def unpack_records(rec):
    rec = np.asarray(rec, np.float32).reshape(-1, RECORD_FLOATS)
    rays = dict(origins=rec[:, 0:3].copy(), directions=rec[:, 3:6].copy(), radii=rec[:, 9].copy(),
                nears=rec[:, 10].copy(), fars=rec[:, 11].copy(), loss_mults=rec[:, 12].copy())
    return rays, rec[:, 13:16].copy()
