import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nerf_or_nothing_b200 as nb
from nerf_or_nothing_b200.scene import synthetic_rays
R, S = 65536, 128
kw = dict(n_samples=S, net_depth=4, net_width=128, net_depth_condition=1, net_width_condition=64, skip_layer=4, deg_point=16, deg_view=4)
m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, precision="bf16", **kw))
rays, pix = synthetic_rays(R, width=800, height=800, n_views=8, seed=1)
m.set_pixels(pix)
args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
for _ in range(2):
    m.GetGradient(*args)
print("done")
