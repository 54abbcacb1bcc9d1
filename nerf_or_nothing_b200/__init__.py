"""nerf_or_nothing_b200 — Python host-side mirror of the reference's operator interface over the C ABI of
``libnerfb200.so`` (include/nerfb200.h).

The reference's host is C# calling five C++/CLI classes of ``AcceleratedNeRFUtils``; neither toolchain exists
here, so the classes below keep the reference's names, argument meaning and call order
(ScratchNerf/ScratchNerf/Program.cs:21-62) on top of ctypes:

    AcceleratedMipNeRF            ANU/AcceleratedMipNeRF.h:10-41   (.mlp -> AcceleratedMLP, ANU/AcceleratedMLP.h:7-45)
    AcceleratedAdamOptimizer      ANU/AcceleratedAdamOptimizer.h:5-20
    AcceleratedGradientCalculator ANU/AcceleratedGradientCalculator.h:8-17
    OutputRetriever               ANU/OutputRetriever.h:7-11

There is NO CPU fallback and nothing here imports the test oracle: if the CUDA library is missing or no
sm_100 GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libnerfb200.so"

PRECISION_FP32 = 0
PRECISION_FP32_TC = 1
PRECISION_BF16_TC = 2
PRECISIONS = {"fp32": PRECISION_FP32, "fp32_tc": PRECISION_FP32_TC, "bf16": PRECISION_BF16_TC, "bf16_tc": PRECISION_BF16_TC}
COMM_ID_BYTES = 128
# nerf_config.engine_flags (include/nerfb200.h)
FLAG_NO_FUSED_FORWARD, FLAG_NO_FUSED_TRAIN_FORWARD, FLAG_NO_FUSED_DGRAD, FLAG_NO_DEFERRED_REDUCE, FLAG_NO_FUSED_ENCODE = 1, 2, 4, 8, 16
FLAG_FUSED_ENCODE_TRAIN, FLAG_NO_WEIGHT_MULTICAST, FLAG_WGRAD_FP16, FLAG_NO_FP8_CORRECTIONS, FLAG_PAIR_MMA = 32, 64, 128, 256, 512


class NerfError(RuntimeError):
    pass


class NerfConfig(C.Structure):
    """Mirror of ``nerf_config`` (include/nerfb200.h)."""

    _fields_ = [
        ("n_rays", C.c_int), ("n_samples", C.c_int), ("n_levels", C.c_int), ("net_depth", C.c_int),
        ("net_width", C.c_int), ("net_depth_condition", C.c_int), ("net_width_condition", C.c_int),
        ("skip_layer", C.c_int), ("deg_point", C.c_int), ("deg_view", C.c_int), ("white_bkgd", C.c_int),
        ("randomized", C.c_int), ("adam_eps_mode", C.c_int), ("last_sample_mode", C.c_int),
        ("precision", C.c_int), ("device", C.c_int), ("chunk_rays", C.c_int),
        ("density_bias", C.c_float), ("rgb_padding", C.c_float), ("coarse_loss_mult", C.c_float),
        ("resample_padding", C.c_float), ("seed", C.c_uint64), ("engine_flags", C.c_uint32),
    ]


CALLBACK = C.CFUNCTYPE(C.c_uint64, C.c_uint64, C.c_int, C.c_float, C.c_uint64, C.c_void_p)

_lib = None

_VP, _I, _L, _F = C.c_void_p, C.c_int, C.c_long, C.c_float
_SIGNATURES = {
    # name: argtypes (restype is int unless listed in _RESTYPES)
    "nerf_default_config": [C.POINTER(NerfConfig)],
    "nerf_device_count": [C.POINTER(_I)],
    "nerf_mipnerf_create": [C.POINTER(NerfConfig), C.POINTER(_VP)],
    "nerf_mipnerf_destroy": [_VP],
    "nerf_mipnerf_num_tensors": [_VP, C.POINTER(_I)],
    "nerf_mipnerf_get_layer_sizes": [_VP, C.POINTER(_I), C.POINTER(_I)],
    "nerf_mipnerf_all_params": [_VP, C.POINTER(_VP)],
    "nerf_mipnerf_all_gradients": [_VP, C.POINTER(_VP)],
    "nerf_mipnerf_flat_params": [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_L)],
    "nerf_mipnerf_set_params": [_VP, _VP, _L],
    "nerf_mipnerf_get_params": [_VP, _VP, _L],
    "nerf_mipnerf_get_gradients": [_VP, _VP, _L],
    "nerf_mipnerf_set_pixels": [_VP, _VP, _I],
    "nerf_mipnerf_set_sampling_uniforms": [_VP, _VP, _I],
    "nerf_mipnerf_set_step": [_VP, C.c_uint32],
    "nerf_mipnerf_get_gradient": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, C.POINTER(_VP)],
    "nerf_mipnerf_get_gradient_dev": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I],
    "nerf_mipnerf_render": [_VP, _VP, _VP, _VP, _VP, _VP, _L, _VP, _VP, _VP],
    "nerf_mipnerf_render_dev": [_VP, _VP, _VP, _VP, _VP, _VP, _L, _VP, _VP, _VP],
    "nerf_mipnerf_level_outputs": [_VP, _I] + [C.POINTER(C.c_uint64)] * 5,
    "nerf_mipnerf_render_view": [_VP, _VP, _F, _I, _I, _F, _F, _I, _L, _L, _VP, _VP, _VP, _I],
    "nerf_generate_rays": [_VP, _F, _I, _I, _F, _F, _I, _L, _L, _VP, _VP, _VP, _VP, _VP],
    "nerf_mipnerf_get_loss": [_VP, C.POINTER(_F), C.POINTER(_F)],
    "nerf_mipnerf_synchronize": [_VP],
    "nerf_mipnerf_launch_count": [_VP, C.POINTER(_L)],
    "nerf_mipnerf_stream": [_VP, C.POINTER(C.c_uint64)],
    "nerf_mipnerf_set_profiling": [_VP, _I],
    "nerf_mipnerf_read_profile": [_VP, _I, C.POINTER(_I), C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(_L), _I],
    "nerf_mlp_get_output": [_VP, _VP, _VP, _I, _I, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    "nerf_mlp_get_gradient": [_VP, _VP, _VP, _I, C.POINTER(_VP)],
    "nerf_mlp_reset_gradients": [_VP, _I],
    "nerf_mlp_relu_bits": [_VP, _I, _I, C.POINTER(C.c_uint64), C.POINTER(_I)],
    "nerf_adam_create": [C.POINTER(_I), _I, _I, _I, C.POINTER(_VP)],
    "nerf_adam_step": [_VP, C.POINTER(_VP), C.POINTER(_VP), _F],
    "nerf_adam_state": [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_L), C.POINTER(_I)],
    "nerf_adam_set_state": [_VP, _VP, _VP, _L, _I],
    "nerf_adam_destroy": [_VP],
    "nerf_gradcalc_create": [_I, _I, _F, _I, C.POINTER(_VP)],
    "nerf_gradcalc_get_output_gradient": [_VP, C.c_uint64, _VP, _I, C.c_uint64, _F, _I, C.POINTER(C.c_uint64)],
    "nerf_gradcalc_destroy": [_VP],
    "nerf_retrieve_output": [C.c_uint64, _I, _VP],
    "nerf_mipnerf_train_step": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _F, C.POINTER(_F)],
    "nerf_mipnerf_train_step_dev": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _F, C.POINTER(_F)],
    "nerf_comm_get_unique_id": [_VP],
    "nerf_mipnerf_comm_init": [_VP, _VP, _I, _I],
    "nerf_mipnerf_allreduce_gradients": [_VP],
    "nerf_mipnerf_comm_destroy": [_VP],
    "nerf_get_sample_t_vals": [_VP, _VP, _VP, _I, _I, _I, _VP],
    "nerf_get_resampled_t_vals": [_VP, _VP, _VP, _I, _I, _F, _I, _VP],
    "nerf_cast_rays": [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I],
    "nerf_encode_input_data": [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I],
    "nerf_apply_layer": [_VP, _VP, _VP, _VP, _VP, _VP, _L, _I, _I, _I, _I],
    "nerf_backpropagate_layer": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _L, _I, _I, _I, _I],
    "nerf_volumetric_rendering": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I],
    "nerf_get_output_gradient": [_VP, _VP, _VP, _VP, _F, _F, _I],
    "nerf_volumetric_rendering_gradient": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I],
    "nerf_adam_optimizer_step": [_VP, _VP, _VP, _VP, _F, _F, _F, _F, _F, _L, _I],
    "nerf_volumetric_rendering_async": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _F, _F, _VP],
    "nerf_volumetric_rendering_gradient_async": [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _F, _F, _VP],
    "nerf_dataset_create": [_VP, _L, _I, C.POINTER(_VP)],
    "nerf_dataset_load": [C.c_char_p, _I, C.POINTER(_VP)],
    "nerf_dataset_size": [_VP, C.POINTER(_L)],
    "nerf_dataset_destroy": [_VP],
    "nerf_dataset_draw_indices": [_VP, C.c_uint64, C.c_uint32, C.c_uint32, _I, _VP],
    "nerf_dataset_gather": [_VP, _VP, C.c_uint64, C.c_uint32, C.c_uint32, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP],
    "nerf_mipnerf_train_step_dataset": [_VP, _VP, _VP, _I, C.c_uint64, _F, C.POINTER(_F)],
    "nerf_image_error": [_VP, _VP, _L, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "nerf_image_ssim": [_VP, _VP, _I, _I, _F, _I, _F, _F, _F, _I, C.POINTER(C.c_double), _VP],
    "nerf_learning_rate_decay": [_I, _F, _F, _I, _I, _F],
    "nerf_checkpoint_save": [_VP, _VP, C.c_char_p],
    "nerf_checkpoint_load": [_VP, _VP, C.c_char_p],
    "nerf_version": [],
    "nerf_last_error": [],
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load libnerfb200.so (built by ``python -m nerf_or_nothing_b200.build``).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NerfError(f"{LIB_PATH} not built: run `python -m nerf_or_nothing_b200.build` (there is no CPU fallback)")
        l = C.CDLL(str(LIB_PATH))
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        l.nerf_last_error.restype = C.c_char_p
        l.nerf_learning_rate_decay.restype = C.c_float
        l.nerf_default_config.restype = None
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise NerfError(f"libnerfb200 error {status}: {lib().nerf_last_error().decode(errors='replace')}")


def default_config(**kw) -> NerfConfig:
    c = NerfConfig()
    lib().nerf_default_config(C.byref(c))
    for k, v in kw.items():
        if k == "precision" and isinstance(v, str):
            v = PRECISIONS[v]
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def _host(a, dtype=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _hp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(x):
    """Device pointer of a torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.c_void_p(x.data_ptr())


class AcceleratedMLP:
    """``AcceleratedMipNeRF.mlp`` (ANU/AcceleratedMLP.h:7-45)."""

    def __init__(self, model: "AcceleratedMipNeRF"):
        self._m = model

    @property
    def allParams(self):  # ANU/AcceleratedMLP.h:25
        n = self._m.num_tensors
        p = (C.c_void_p * n)()
        check(lib().nerf_mipnerf_all_params(self._m._h, p))
        return [int(x or 0) for x in p]

    @property
    def allGradients(self):  # ANU/AcceleratedMLP.h:24
        n = self._m.num_tensors
        p = (C.c_void_p * n)()
        check(lib().nerf_mipnerf_all_gradients(self._m._h, p))
        return [int(x or 0) for x in p]

    def get_layer_sizes(self):  # ANU/AcceleratedMLP.cpp:131-154
        return self._m.GetLayerSizes()

    def get_output(self, dev_encoded_position, dev_encoded_direction, level, n_rays):
        """ANU/AcceleratedMLP.cpp:214-255 -> (density_dev_ptr, rgb_dev_ptr) (named, SURVEY A-D2)."""
        d, r = C.c_uint64(), C.c_uint64()
        check(lib().nerf_mlp_get_output(self._m._h, _dp(dev_encoded_position), _dp(dev_encoded_direction), level,
                                        n_rays, C.byref(d), C.byref(r)))
        return d.value, r.value

    def get_gradient(self, color_gradient, density_gradient, level):  # ANU/AcceleratedMLP.cpp:256-321
        check(lib().nerf_mlp_get_gradient(self._m._h, _dp(color_gradient), _dp(density_gradient), level, None))
        return self.allGradients

    def reset_gradients(self, level=0):  # ANU/AcceleratedMLP.cpp:113-129
        check(lib().nerf_mlp_reset_gradients(self._m._h, level))

    def relu_bits(self, level, layer):
        """(device pointer, words per row) of the ReLU bit plane of hidden layer `layer` cached by the last forward of `level`."""
        p, w = C.c_uint64(), C.c_int()
        check(lib().nerf_mlp_relu_bits(self._m._h, level, layer, C.byref(p), C.byref(w)))
        return p.value, w.value


class AcceleratedMipNeRF:
    """ANU/AcceleratedMipNeRF.h:10-41.  ``AcceleratedMipNeRF()`` with no arguments reproduces the reference's
    compile-time configuration (1024 rays, 128+128 samples, 8x256 MLP)."""

    def __init__(self, config: NerfConfig | None = None, **kw):
        self.cfg = config if config is not None else default_config(**kw)
        h = C.c_void_p()
        check(lib().nerf_mipnerf_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        n = C.c_int()
        check(lib().nerf_mipnerf_num_tensors(self._h, C.byref(n)))
        self.num_tensors = n.value
        self.mlp = AcceleratedMLP(self)
        self._cb_keep = None

    def close(self):
        if getattr(self, "_h", None):
            lib().nerf_mipnerf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface
    def GetLayerSizes(self):  # ANU/AcceleratedMipNeRF.cpp:146-149
        s = (C.c_int * self.num_tensors)()
        n = C.c_int()
        check(lib().nerf_mipnerf_get_layer_sizes(self._h, s, C.byref(n)))
        return list(s)

    def GetGradient(self, origins, directions, radii, nears, fars, lossMultipliers, getOutputGradient=None):
        """ANU/AcceleratedMipNeRF.cpp:52-144.  ``getOutputGradient(comp_rgb_dev, level, loss_mult_sum,
        loss_mults_dev) -> device pointer`` or None for the built-in MSE against ``set_pixels``."""
        a = [_host(x) for x in (origins, directions, radii, nears, fars, lossMultipliers)]
        n = a[0].shape[0]
        cb = None
        if getOutputGradient is not None:
            def _tramp(comp, level, lms, lm_dev, _user):
                return int(getOutputGradient(comp, level, lms, lm_dev))
            cb = CALLBACK(_tramp)
            self._cb_keep = cb
        check(lib().nerf_mipnerf_get_gradient(self._h, *[_hp(x) for x in a], n, C.cast(cb, C.c_void_p) if cb else None,
                                              None, None))
        return self.mlp.allGradients

    # ---- additions the metric needs
    @property
    def num_params(self):
        n = C.c_long()
        check(lib().nerf_mipnerf_flat_params(self._h, None, None, C.byref(n)))
        return n.value

    def flat_pointers(self):
        p, g, n = C.c_void_p(), C.c_void_p(), C.c_long()
        check(lib().nerf_mipnerf_flat_params(self._h, C.byref(p), C.byref(g), C.byref(n)))
        return p.value, g.value, n.value

    def set_params(self, flat):
        flat = _host(flat)
        check(lib().nerf_mipnerf_set_params(self._h, _hp(flat), flat.shape[0]))

    def get_params(self):
        out = np.empty(self.num_params, np.float32)
        check(lib().nerf_mipnerf_get_params(self._h, _hp(out), out.shape[0]))
        return out

    def get_gradients(self):
        out = np.empty(self.num_params, np.float32)
        check(lib().nerf_mipnerf_get_gradients(self._h, _hp(out), out.shape[0]))
        return out

    def set_pixels(self, pixels):
        p = _host(pixels)
        check(lib().nerf_mipnerf_set_pixels(self._h, _hp(p), p.shape[0]))

    def set_sampling_uniforms(self, u):
        if u is None:
            check(lib().nerf_mipnerf_set_sampling_uniforms(self._h, None, 0))
            return
        u = _host(u)
        assert u.ndim == 3 and u.shape[0] == self.cfg.n_levels and u.shape[2] == self.cfg.n_samples + 1
        check(lib().nerf_mipnerf_set_sampling_uniforms(self._h, _hp(u), u.shape[1]))

    def set_step(self, step):
        check(lib().nerf_mipnerf_set_step(self._h, int(step)))

    def get_gradient_dev(self, origins, directions, radii, nears, fars, loss_mults, pixels, n_rays):
        check(lib().nerf_mipnerf_get_gradient_dev(self._h, *[_dp(x) for x in (origins, directions, radii, nears, fars,
                                                                               loss_mults, pixels)], n_rays))

    def render(self, origins, directions, radii, nears, fars):
        a = [_host(x) for x in (origins, directions, radii, nears, fars)]
        n = a[0].shape[0]
        rgb, depth, acc = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        check(lib().nerf_mipnerf_render(self._h, *[_hp(x) for x in a], n, _hp(rgb), _hp(depth), _hp(acc)))
        return rgb, depth, acc

    def render_dev(self, origins, directions, radii, nears, fars, n_rays, rgb, depth=None, acc=None):
        check(lib().nerf_mipnerf_render_dev(self._h, *[_dp(x) for x in (origins, directions, radii, nears, fars)], n_rays,
                                            _dp(rgb), _dp(depth), _dp(acc)))

    def render_view(self, c2w, focal, width, height, near=2.0, far=6.0, edge_mode=1, first_pixel=0, n_pixels=None):
        """Render a view from a 3x4 camera-to-world pose; rays are generated on the device (SN/Dataset.cs:111-176).
        Returns (rgb [n,3], depth [n], acc [n]) for the pixel range, row-major."""
        c = _host(np.asarray(c2w, np.float32).reshape(12))
        n = width * height - first_pixel if n_pixels is None else n_pixels
        rgb, depth, acc = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        check(lib().nerf_mipnerf_render_view(self._h, _hp(c), focal, width, height, near, far, edge_mode, first_pixel, n,
                                             _hp(rgb), _hp(depth), _hp(acc), 0))
        return rgb, depth, acc

    def level_outputs(self, level):
        v = [C.c_uint64() for _ in range(5)]
        check(lib().nerf_mipnerf_level_outputs(self._h, level, *[C.byref(x) for x in v]))
        return dict(zip(("comp_rgb", "depth", "acc", "weights", "t_vals"), [x.value for x in v]))

    def get_loss(self):
        per = (C.c_float * self.cfg.n_levels)()
        tot = C.c_float()
        check(lib().nerf_mipnerf_get_loss(self._h, per, C.byref(tot)))
        return list(per), tot.value

    def synchronize(self):
        check(lib().nerf_mipnerf_synchronize(self._h))

    def stream(self):
        """cudaStream_t (as int) every call on this handle enqueues on."""
        s = C.c_uint64()
        check(lib().nerf_mipnerf_stream(self._h, C.byref(s)))
        return s.value

    def set_profiling(self, on=True):
        check(lib().nerf_mipnerf_set_profiling(self._h, int(bool(on))))

    def read_profile(self, reset=True):
        """{kernel family: (total ms, launches)} measured with CUDA events on the handle's stream."""
        n, names = C.c_int(), (C.c_char_p * 32)()
        ms, launches = (C.c_double * 32)(), (C.c_long * 32)()
        check(lib().nerf_mipnerf_read_profile(self._h, 32, C.byref(n), names, ms, launches, int(reset)))
        return {names[i].decode(): (ms[i], launches[i]) for i in range(n.value) if launches[i] > 0 or ms[i] > 0}

    def launch_count(self):
        n = C.c_long()
        check(lib().nerf_mipnerf_launch_count(self._h, C.byref(n)))
        return n.value

    def train_step(self, optimizer, origins, directions, radii, nears, fars, loss_mults, pixels, lr, want_loss=True):
        """Host part of SN/Program.cs:48-62 in one call (HOST arrays; H2D inside)."""
        a = [_host(x) for x in (origins, directions, radii, nears, fars, loss_mults, pixels)]
        loss = C.c_float()
        check(lib().nerf_mipnerf_train_step(self._h, optimizer._h, *[_hp(x) for x in a], a[0].shape[0], lr,
                                            C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def train_step_dev(self, optimizer, origins, directions, radii, nears, fars, loss_mults, pixels, n_rays, lr,
                       want_loss=False):
        loss = C.c_float()
        check(lib().nerf_mipnerf_train_step_dev(self._h, optimizer._h, *[_dp(x) for x in (origins, directions, radii,
                                                nears, fars, loss_mults, pixels)], n_rays, lr,
                                                C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def train_step_dataset(self, optimizer, dataset, n_rays, sampler_seed, lr, want_loss=True):
        """One iteration of Train() (SN/Program.cs:28-45) with the batch drawn and gathered on the device from a resident
        BinDataset: no per-step host->device traffic."""
        loss = C.c_float()
        check(lib().nerf_mipnerf_train_step_dataset(self._h, optimizer._h, dataset._h, n_rays, sampler_seed, lr,
                                                    C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    # ---- checkpoints
    def save_checkpoint(self, path, optimizer=None):
        check(lib().nerf_checkpoint_save(self._h, optimizer._h if optimizer is not None else None, str(path).encode()))

    def load_checkpoint(self, path, optimizer=None):
        check(lib().nerf_checkpoint_load(self._h, optimizer._h if optimizer is not None else None, str(path).encode()))

    # ---- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        check(lib().nerf_mipnerf_comm_init(self._h, buf, rank, world))

    def allreduce_gradients(self):
        check(lib().nerf_mipnerf_allreduce_gradients(self._h))


class BinDataset:
    """SN/BinDataset.cs:10-52 with the records resident in device memory (64-byte records, `scene.pack_records`)."""

    def __init__(self, records_or_path, device=0):
        self._h = C.c_void_p()
        if isinstance(records_or_path, (str, bytes)) or hasattr(records_or_path, "__fspath__"):
            check(lib().nerf_dataset_load(str(records_or_path).encode(), device, C.byref(self._h)))
        else:
            rec = np.ascontiguousarray(records_or_path, np.float32).reshape(-1, 16)
            check(lib().nerf_dataset_create(_hp(rec), rec.shape[0], device, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.nerf_dataset_destroy(self._h)
            self._h = None

    def __len__(self):
        n = C.c_long()
        check(lib().nerf_dataset_size(self._h, C.byref(n)))
        return n.value

    def draw_indices(self, seed, step, n_rays, first_slot=0):
        idx = np.empty(n_rays, np.int64)
        check(lib().nerf_dataset_draw_indices(self._h, seed, step, first_slot, n_rays, _hp(idx)))
        return idx

    def gather(self, n_rays, indices=None, seed=0, step=0, first_slot=0):
        """LoadBatch on the device; returns host copies of the SoA batch (parity hook)."""
        import torch
        bufs = {k: torch.empty((n_rays, w), dtype=torch.float32, device="cuda")
                for k, w in (("origins", 3), ("directions", 3), ("radii", 1), ("nears", 1), ("fars", 1), ("loss_mults", 1), ("pixels", 3))}
        idx = None if indices is None else np.ascontiguousarray(indices, np.int64)
        check(lib().nerf_dataset_gather(self._h, _hp(idx) if idx is not None else None, seed, step, first_slot, n_rays,
                                        *[C.c_void_p(bufs[k].data_ptr()) for k in ("origins", "directions", "radii", "nears", "fars",
                                                                                   "loss_mults", "pixels")]))
        out = {k: v.cpu().numpy() for k, v in bufs.items()}
        for k in ("radii", "nears", "fars", "loss_mults"):
            out[k] = out[k][:, 0]
        return out


def generate_rays(c2w, focal, width, height, near=2.0, far=6.0, edge_mode=1, first_pixel=0, n_pixels=None):
    """Dataset.GenerateRays (SN/Dataset.cs:111-176) on the device; returns host copies of the SoA ray arrays (parity hook)."""
    import torch
    c = _host(np.asarray(c2w, np.float32).reshape(12))
    n = width * height - first_pixel if n_pixels is None else n_pixels
    bufs = {k: torch.empty((n, w), dtype=torch.float32, device="cuda") for k, w in (("origins", 3), ("directions", 3), ("radii", 1), ("nears", 1), ("fars", 1))}
    check(lib().nerf_generate_rays(_hp(c), focal, width, height, near, far, edge_mode, first_pixel, n,
                                   *[C.c_void_p(bufs[k].data_ptr()) for k in ("origins", "directions", "radii", "nears", "fars")]))
    out = {k: v.cpu().numpy() for k, v in bufs.items()}
    for k in ("radii", "nears", "fars"):
        out[k] = out[k][:, 0]
    return out


def image_error(a, b):
    """(mse, psnr) of two images; psnr = MseToPsnr (SN/MipHelpers.cs:672).  Computed on the device."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        raise ValueError("image_error: shape mismatch")
    mse, psnr = C.c_double(), C.c_double()
    check(lib().nerf_image_error(_hp(a), _hp(b), a.size, 0, C.byref(mse), C.byref(psnr)))
    return mse.value, psnr.value


def image_ssim(a, b, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, want_map=False):
    """ComputeSsimAverage (SN/MipHelpers.cs:728-737) of two [H, W, 3] images, computed on the device.
    Returns the mean SSIM, or (mean, map [H, W, 3]) with want_map."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("image_ssim: two [H, W, 3] images expected")
    mean = C.c_double()
    m = np.empty_like(a) if want_map else None
    check(lib().nerf_image_ssim(_hp(a), _hp(b), a.shape[1], a.shape[0], max_val, filter_size, filter_sigma, k1, k2, 0,
                                C.byref(mean), _hp(m) if want_map else None))
    return (mean.value, m) if want_map else mean.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().nerf_comm_get_unique_id(buf))
    return buf.raw


class AcceleratedAdamOptimizer:
    """ANU/AcceleratedAdamOptimizer.h:5-20."""

    def __init__(self, layer_sizes, eps_mode=0, device=0):
        s = (C.c_int * len(layer_sizes))(*layer_sizes)
        h = C.c_void_p()
        check(lib().nerf_adam_create(s, len(layer_sizes), eps_mode, device, C.byref(h)))
        self._h = h
        self._n = len(layer_sizes)

    def step(self, params, grads, learning_rate):  # ANU/AcceleratedAdamOptimizer.cpp:23-41
        p = (C.c_void_p * self._n)(*params)
        g = (C.c_void_p * self._n)(*grads)
        check(lib().nerf_adam_step(self._h, p, g, learning_rate))

    def state(self):
        m, v, n, it = C.c_void_p(), C.c_void_p(), C.c_long(), C.c_int()
        check(lib().nerf_adam_state(self._h, C.byref(m), C.byref(v), C.byref(n), C.byref(it)))
        return m.value, v.value, n.value, it.value

    def set_state(self, m, v, iteration):
        m, v = _host(m), _host(v)
        check(lib().nerf_adam_set_state(self._h, _hp(m), _hp(v), m.shape[0], iteration))

    def close(self):
        if getattr(self, "_h", None):
            lib().nerf_adam_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AcceleratedGradientCalculator:
    """ANU/AcceleratedGradientCalculator.h:8-17."""

    def __init__(self, batch_size, n_levels=2, coarse_loss_mult=0.1, device=0):
        h = C.c_void_p()
        check(lib().nerf_gradcalc_create(batch_size, n_levels, coarse_loss_mult, device, C.byref(h)))
        self._h = h

    def get_output_gradient(self, input_dev, pixels, loss_mults_dev, loss_mult_sum, level):  # .cpp:18-30
        p = _host(pixels)
        out = C.c_uint64()
        check(lib().nerf_gradcalc_get_output_gradient(self._h, int(input_dev), _hp(p), p.shape[0], int(loss_mults_dev),
                                                      loss_mult_sum, level, C.byref(out)))
        return out.value

    def close(self):
        if getattr(self, "_h", None):
            lib().nerf_gradcalc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OutputRetriever:
    """ANU/OutputRetriever.h:7-11."""

    @staticmethod
    def RetrieveOutput(dev_output, size):  # ANU/OutputRetriever.cpp:6-14
        out = np.empty((size, 3), np.float32)
        check(lib().nerf_retrieve_output(int(dev_output), size, _hp(out)))
        return out


def learning_rate_decay(step, lr_init=5e-4, lr_final=5e-6, max_steps=1000000, lr_delay_steps=2500, lr_delay_mult=0.01):
    """SN/MipHelpers.cs:758-773 with the defaults of SN/TrainState.cs:54-60 — the library's C entry point."""
    return float(lib().nerf_learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult))


def _learning_rate_decay_numpy(step, lr_init=5e-4, lr_final=5e-6, max_steps=1000000, lr_delay_steps=2500, lr_delay_mult=0.01):
    """numpy float32 restatement of the same schedule (kept for cross-checks)."""
    f = np.float32
    delay = f(1.0)
    if lr_delay_steps > 0:
        p = np.clip(f(step) / f(lr_delay_steps), f(0), f(1))
        delay = f(lr_delay_mult) + (f(1) - f(lr_delay_mult)) * np.sin(f(0.5) * f(np.pi) * p, dtype=f)
    t = np.clip(f(step) / f(max_steps), f(0), f(1))
    return float(delay * np.exp(np.log(f(lr_init)) * (f(1) - t) + np.log(f(lr_final)) * t, dtype=f))
