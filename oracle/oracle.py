"""ctypes/numpy front end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs — never by the product package.  Parity is
unpinned by the reference's own tests (it has none); see oracle/oracle.h for the pins used.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_REF_PATH = _HERE / "_ref" / "libref_kernels.so"


class OrcConfig(C.Structure):
    """Mirror of ``orc_config`` (oracle.h)."""

    _fields_ = [
        ("n_samples", C.c_int), ("n_levels", C.c_int), ("net_depth", C.c_int), ("net_width", C.c_int),
        ("net_depth_condition", C.c_int), ("net_width_condition", C.c_int), ("skip_layer", C.c_int),
        ("deg_point", C.c_int), ("deg_view", C.c_int), ("white_bkgd", C.c_int),
        ("adam_eps_mode", C.c_int), ("last_sample_mode", C.c_int), ("randomized", C.c_int),
        ("reserved", C.c_int), ("density_bias", C.c_double), ("rgb_padding", C.c_double),
        ("coarse_loss_mult", C.c_double), ("resample_padding", C.c_double),
    ]


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists) via oracle/Makefile."""
    srcs = [_HERE / "oracle.c", _HERE / "oracle_impl.h", _HERE / "oracle.h"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale or (Path("/root/reference").exists() and not _REF_PATH.exists()):
        subprocess.run(["make", "-C", str(_HERE)], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_num_params.restype = C.c_long
        _lib.orc_learning_rate_decay.restype = C.c_float
        _lib.orc_learning_rate_decay.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]
        for suf in ("_f32", "_f64"):
            getattr(_lib, "orc_train_gradient" + suf).restype = C.c_double
    return _lib


def default_config(**kw) -> OrcConfig:
    c = OrcConfig()
    lib().orc_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dt(prec):
    return (np.float32, "_f32") if prec in ("f32", np.float32) else (np.float64, "_f64")


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_layers(cfg):
    return lib().orc_num_layers(C.byref(cfg))


def layer_sizes(cfg):
    L = num_layers(cfg)
    s = (C.c_int * (2 * L))()
    lib().orc_layer_sizes(C.byref(cfg), s)
    return list(s)


def layer_shapes(cfg):
    L = num_layers(cfg)
    o, a, b = (C.c_int * L)(), (C.c_int * L)(), (C.c_int * L)()
    lib().orc_layer_shapes(C.byref(cfg), o, a, b)
    return list(o), list(a), list(b)


def num_params(cfg):
    return int(lib().orc_num_params(C.byref(cfg)))


def act_stride(cfg):
    return cfg.net_depth * cfg.net_width + cfg.net_depth_condition * cfg.net_width_condition


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def sampling_uniforms(seed, step, level, ray0, n_rays, n):
    u = np.empty((n_rays, n), np.float32)
    lib().orc_sampling_uniforms(C.c_uint64(seed), C.c_uint32(step), C.c_uint32(level), C.c_uint32(ray0),
                                int(n_rays), int(n), _p(u))
    return u


def init_params(cfg, seed):
    p = np.empty(num_params(cfg), np.float32)
    lib().orc_init_params(C.byref(cfg), C.c_uint64(seed), _p(p))
    return p


def learning_rate_decay(step, lr_init=5e-4, lr_final=5e-6, max_steps=1000000, lr_delay_steps=2500,
                        lr_delay_mult=0.01):
    return float(lib().orc_learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult))


def generate_rays(c2w, focal, width, height, near=2.0, far=6.0, edge_mode=1, first=0, n=None):
    """Dataset.GenerateRays (SN/Dataset.cs:111-176) for one camera; c2w 3x4 row-major [R | t]."""
    c = _arr(np.asarray(c2w).reshape(12), np.float32)
    n = width * height - first if n is None else n
    out = dict(origins=np.empty((n, 3), np.float32), directions=np.empty((n, 3), np.float32), radii=np.empty(n, np.float32),
               nears=np.empty(n, np.float32), fars=np.empty(n, np.float32))
    fn = lib().orc_generate_rays
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_long, C.c_long] + [C.c_void_p] * 5
    fn(_p(c), focal, width, height, near, far, edge_mode, first, n, *[_p(out[k]) for k in ("origins", "directions", "radii", "nears", "fars")])
    return out


def ssim(a, b, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """(mean, map) of ComputeSsim / ComputeSsimAverage (SN/MipHelpers.cs:688-737); images [H, W, 3] float32."""
    a, b = _arr(a, np.float32), _arr(b, np.float32)
    H, W = a.shape[:2]
    m = np.empty_like(a)
    fn = lib().orc_ssim
    fn.restype = C.c_double
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    mean = fn(_p(a), _p(b), W, H, max_val, filter_size, filter_sigma, k1, k2, _p(m))
    return float(mean), m


# ---------------------------------------------------------------- per-stage functions


def sample_t_vals(nears, fars, u, S, randomized=1, prec="f32"):
    dt, suf = _dt(prec)
    nears, fars = _arr(nears, dt), _arr(fars, dt)
    R = nears.shape[0]
    u = _arr(u if u is not None else np.zeros((R, S + 1)), np.float32)
    t = np.empty((R, S + 1), dt)
    getattr(lib(), "orc_sample_t_vals" + suf)(_p(nears), _p(fars), _p(u), R, S, int(randomized), _p(t))
    return t


def resample_t_vals(t, w, u, padding=0.01, randomized=1, prec="f32"):
    dt, suf = _dt(prec)
    t, w = _arr(t, dt), _arr(w, dt)
    R, S = w.shape
    u = _arr(u if u is not None else np.zeros((R, S + 1)), np.float32)
    out = np.empty((R, S + 1), dt)
    fn = getattr(lib(), "orc_resample_t_vals" + suf)
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_float if dt == np.float32 else C.c_double, C.c_int, C.c_void_p]
    fn(_p(t), _p(w), _p(u), R, S, padding, int(randomized), _p(out))
    return out


def cast_rays(t, o, d, radii, prec="f32"):
    dt, suf = _dt(prec)
    t, o, d, radii = _arr(t, dt), _arr(o, dt), _arr(d, dt), _arr(radii, dt)
    R, S = t.shape[0], t.shape[1] - 1
    mean, cov = np.empty((R, S, 3), dt), np.empty((R, S, 3), dt)
    getattr(lib(), "orc_cast_rays" + suf)(_p(t), _p(o), _p(d), _p(radii), R, S, _p(mean), _p(cov))
    return mean, cov


def encode_position(mean, cov, deg=16, prec="f32"):
    dt, suf = _dt(prec)
    mean, cov = _arr(mean, dt).reshape(-1, 3), _arr(cov, dt).reshape(-1, 3)
    M = mean.shape[0]
    enc = np.empty((M, 6 * deg), dt)
    getattr(lib(), "orc_encode_position" + suf)(_p(mean), _p(cov), C.c_long(M), deg, _p(enc))
    return enc


def encode_direction(d, deg=4, prec="f32"):
    dt, suf = _dt(prec)
    d = _arr(d, dt).reshape(-1, 3)
    enc = np.empty((d.shape[0], 3 + 6 * deg), dt)
    getattr(lib(), "orc_encode_direction" + suf)(_p(d), d.shape[0], deg, _p(enc))
    return enc


def mlp_forward(cfg, params, enc_pos, enc_dir, want_acts=True, prec="f32"):
    dt, suf = _dt(prec)
    params, enc_pos, enc_dir = _arr(params, dt), _arr(enc_pos, dt), _arr(enc_dir, dt)
    M = enc_pos.shape[0]
    acts = np.empty((M, act_stride(cfg)), dt) if want_acts else None
    rd, rr = np.empty(M, dt), np.empty((M, 3), dt)
    getattr(lib(), "orc_mlp_forward" + suf)(C.byref(cfg), _p(params), _p(enc_pos), _p(enc_dir), C.c_long(M),
                                            _p(acts), _p(rd), _p(rr))
    return rd, rr, acts


def mlp_backward(cfg, params, enc_pos, enc_dir, acts, d_raw_density, d_raw_rgb, prec="f32"):
    dt, suf = _dt(prec)
    a = [_arr(x, dt) for x in (params, enc_pos, enc_dir, acts, d_raw_density, d_raw_rgb)]
    M = a[1].shape[0]
    g = np.zeros(num_params(cfg), dt)
    getattr(lib(), "orc_mlp_backward" + suf)(C.byref(cfg), *[_p(x) for x in a], C.c_long(M), _p(g))
    return g


def output_activations(cfg, raw_density, raw_rgb, prec="f32"):
    dt, suf = _dt(prec)
    rd, rr = _arr(raw_density, dt).reshape(-1), _arr(raw_rgb, dt).reshape(-1, 3)
    den, rgb = np.empty_like(rd), np.empty_like(rr)
    getattr(lib(), "orc_output_activations" + suf)(C.byref(cfg), _p(rd), _p(rr), C.c_long(rd.shape[0]), _p(den), _p(rgb))
    return den, rgb


def output_activations_grad(cfg, raw_density, raw_rgb, d_density, d_rgb, prec="f32"):
    dt, suf = _dt(prec)
    rd, rr = _arr(raw_density, dt).reshape(-1), _arr(raw_rgb, dt).reshape(-1, 3)
    dd, dr = _arr(d_density, dt).reshape(-1), _arr(d_rgb, dt).reshape(-1, 3)
    o1, o2 = np.empty_like(rd), np.empty_like(rr)
    getattr(lib(), "orc_output_activations_grad" + suf)(C.byref(cfg), _p(rd), _p(rr), _p(dd), _p(dr),
                                                         C.c_long(rd.shape[0]), _p(o1), _p(o2))
    return o1, o2


def volumetric_rendering(rgb, density, t, d, white_bkgd=1, prec="f32"):
    dt, suf = _dt(prec)
    rgb, density, t, d = _arr(rgb, dt), _arr(density, dt), _arr(t, dt), _arr(d, dt)
    R, S = density.shape
    out = dict(comp_rgb=np.empty((R, 3), dt), depth=np.empty(R, dt), acc=np.empty(R, dt),
               weights=np.empty((R, S), dt), alpha=np.empty((R, S), dt), transmittance=np.empty((R, S), dt))
    getattr(lib(), "orc_volumetric_rendering" + suf)(_p(rgb), _p(density), _p(t), _p(d), R, S, int(white_bkgd),
                                                     *[_p(out[k]) for k in ("comp_rgb", "depth", "acc", "weights", "alpha", "transmittance")])
    return out


def output_gradient(comp_rgb, pixels, loss_mults, loss_mult_sum, level_mult, prec="f32"):
    dt, suf = _dt(prec)
    comp_rgb, pixels, loss_mults = _arr(comp_rgb, dt), _arr(pixels, dt), _arr(loss_mults, dt)
    g = np.empty_like(comp_rgb)
    fn = getattr(lib(), "orc_output_gradient" + suf)
    ft = C.c_float if dt == np.float32 else C.c_double
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int, ft, ft, C.c_void_p]
    fn(_p(comp_rgb), _p(pixels), _p(loss_mults), comp_rgb.shape[0], loss_mult_sum, level_mult, _p(g))
    return g


def volumetric_rendering_gradient(g, rgb, density, t, d, white_bkgd=1, last_sample_mode=0, prec="f32"):
    dt, suf = _dt(prec)
    g, rgb, density, t, d = [_arr(x, dt) for x in (g, rgb, density, t, d)]
    R, S = density.shape
    d_rgb, d_den = np.empty((R, S, 3), dt), np.empty((R, S), dt)
    getattr(lib(), "orc_volumetric_rendering_gradient" + suf)(_p(g), _p(rgb), _p(density), _p(t), _p(d), R, S,
                                                              int(white_bkgd), int(last_sample_mode), _p(d_rgb), _p(d_den))
    return d_rgb, d_den


def adam_step(p, g, m, v, lr, iteration, eps_mode=0, prec="f32"):
    """In place on copies; returns (p, m, v)."""
    dt, suf = _dt(prec)
    p, m, v = [np.array(x, dtype=dt, copy=True) for x in (p, m, v)]
    g = _arr(g, dt)
    fn = getattr(lib(), "orc_adam_step" + suf)
    ft = C.c_float if dt == np.float32 else C.c_double
    fn.argtypes = [C.c_void_p] * 4 + [C.c_long, ft, C.c_int, C.c_int]
    fn(_p(p), _p(g), _p(m), _p(v), p.shape[0], lr, int(iteration), int(eps_mode))
    return p, m, v


def train_gradient(cfg, params, rays, pixels, u, with_backward=True, prec="f32"):
    """rays: dict(origins[R,3], directions[R,3], radii, nears, fars, loss_mults); u: [L,R,S+1] float32."""
    dt, suf = _dt(prec)
    R, S, L = rays["origins"].shape[0], cfg.n_samples, cfg.n_levels
    a = [_arr(x, dt) for x in (params, rays["origins"], rays["directions"], rays["radii"], rays["nears"],
                               rays["fars"], rays["loss_mults"], pixels)]
    u = _arr(u, np.float32)
    assert u.shape == (L, R, S + 1)
    out = dict(grads=np.zeros(num_params(cfg), dt) if with_backward else None,
               comp_rgb=np.empty((L, R, 3), dt), depth=np.empty((L, R), dt), acc=np.empty((L, R), dt),
               t_vals=np.empty((L, R, S + 1), dt), weights=np.empty((L, R, S), dt), loss=np.empty(L, dt))
    total = getattr(lib(), "orc_train_gradient" + suf)(
        C.byref(cfg), *[_p(x) for x in a], _p(u), R, int(with_backward),
        *[_p(out[k]) for k in ("grads", "comp_rgb", "depth", "acc", "t_vals", "weights", "loss")])
    out["total_loss"] = float(total)
    return out


# ---------------------------------------------------------------- synthetic scene


def synthetic_rays(*a, **kw):
    """Data generation only (not part of the checked algorithm): nerf_or_nothing_b200/scene.py."""
    from nerf_or_nothing_b200.scene import synthetic_rays as f

    return f(*a, **kw)


# ---------------------------------------------------------------- reference kernels (pin P1, GPU only)


def ref_lib() -> C.CDLL:
    """oracle/_ref/libref_kernels.so — the reference's own CUDA kernels (needs a GPU to call)."""
    if not _REF_PATH.exists():
        build()
    if not _REF_PATH.exists():
        raise FileNotFoundError(f"{_REF_PATH} missing: build it where /root/reference is present")
    lib_ = C.CDLL(str(_REF_PATH))
    return lib_
