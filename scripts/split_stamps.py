"""Debug build only (-DNERF_SPLIT_DBG=1): clock64 stamps of one epilogue warp and the MMA warp of the fp32-accurate fused TRAINING
forward for CTA 0's 4th tile.  Prints, per layer and half, cycles relative to the layer's first stamp.
    NERF_NVCC_DEFS=-DNERF_SPLIT_DBG=1 python -m nerf_or_nothing_b200.build --force ; gpurun -- python scripts/split_stamps.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nerf_or_nothing_b200 as nb
from nerf_or_nothing_b200.scene import synthetic_rays

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R, S = 4096, 128
cfg = nb.default_config(n_rays=R, precision="fp32_tc", n_samples=S, engine_flags=flags)
m = nb.AcceleratedMipNeRF(cfg)
opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
rays, pix = synthetic_rays(R, width=800, height=800, seed=1)
args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"], pix)
for _ in range(4):
    m.train_step(opt, *args, 1e-4)
epi = (C.c_ulonglong * (16 * 2 * 8))()
mma = (C.c_ulonglong * (16 * 2 * 4))()
assert nb.lib().nerf_debug_split_stamps(epi, mma) == 0
e = np.array(epi[:], dtype=np.int64).reshape(16, 2, 8)
q = np.array(mma[:], dtype=np.int64).reshape(16, 2, 4)
t0 = q[0, 0, 0]
print("MMA warp: per layer s, half h: [issue start, wait act_ready begin, wait end, commit issued] (cycles since layer 0 start)")
for s in range(9):
    print(s, [[int(x - t0) if x else None for x in q[s, h]] for h in range(2)])
print("epilogue warp 0: per layer s, half h: [wait acc_full begin, acc_full seen, accumulators in registers + signals, chunk0 done / (h=1: chunk0 computed), ship0 done / (h=1: act_ready signalled), chunk1 done / ship0 done, ship1 done]")
for s in range(9):
    print(s, [[int(x - t0) if x else None for x in e[s, h, :7]] for h in range(2)])
