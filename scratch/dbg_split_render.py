import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nerf_or_nothing_b200 as nb
from nerf_or_nothing_b200.scene import synthetic_rays
R, S = 4096, 128
m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, n_samples=S, precision="fp32_tc"))
rays, pix = synthetic_rays(R, width=800, height=800, n_views=100, seed=1)
args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
m.render(*args)
os.environ["NERF_FUSED_DBG"] = "1"
m.render(*args)
