"""Helpers for the `-m gpu` parity tests: torch is used only as a device-memory allocator."""
import ctypes as C

import numpy as np
import torch

import nerf_or_nothing_b200 as nb
from oracle import oracle as orc


_KEEP = []  # device tensors stay alive until the end of the test: `ptr(dev(x))` must not dangle


def dev(a, dtype=np.float32):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()
    _KEEP.append(t)
    return t


def release():
    _KEEP.clear()


def empty(*shape):
    return torch.empty(*shape, dtype=torch.float32, device="cuda")


def zeros(*shape):
    return torch.zeros(*shape, dtype=torch.float32, device="cuda")


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _DevView:
    def __init__(self, p, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(p), False), "version": 2}


def view_ptr(p, shape):
    """Zero-copy torch view of a library-owned fp32 device buffer."""
    n = int(np.prod(shape))
    return torch.as_tensor(_DevView(p, n), device="cuda").view(*shape)


def from_ptr(p, shape):
    """Copy a library-owned device buffer to the host."""
    torch.cuda.synchronize()
    return view_ptr(p, shape).cpu().numpy().copy()


def call(name, *args):
    nb.check(getattr(nb.lib(), name)(*args))


def rel_err(a, b):
    """max |a-b| relative to the tensor scale max|b| (the fp32-path tolerance of BASELINE.md is on this)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def configs_pair(**kw):
    """Matching (nerf_config, orc_config) for the same hyper-parameters."""
    ncfg = nb.default_config(**kw)
    okw = {k: v for k, v in kw.items() if k not in ("n_rays", "precision", "device", "chunk_rays", "seed", "engine_flags")}
    ocfg = orc.default_config(**okw)
    return ncfg, ocfg


def batch(R, S, width=100, seed=2024, levels=2, u_seed=99, step=0):
    rays, pix = orc.synthetic_rays(R, width=width, height=width, seed=seed)
    u = np.stack([orc.sampling_uniforms(u_seed, step, lv, 0, R, S + 1) for lv in range(levels)])
    return rays, pix, u
