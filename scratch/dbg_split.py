import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nerf_or_nothing_b200 as nb
from nerf_or_nothing_b200.scene import synthetic_rays
R, S = 4096, 128
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32_tc"
m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, n_samples=S, precision=prec))
rays, pix = synthetic_rays(R, width=800, height=800, n_views=100, seed=1)
m.set_pixels(pix)
args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
m.GetGradient(*args)
os.environ["NERF_FUSED_DBG"] = "1"
m.GetGradient(*args)
