"""CPU simulation: forward error of alternative split products through the 8x256 MLP (vs fp64).
    python scripts/fp8_corrections_precision.py [weight scale]      f8c_gpu = the representation the fused forward kernels use
(fp16 main product + two E4M3 correction products with the kernel's powers of two); bf16x3 = hi*hi + lo*hi + hi*lo in bf16."""
import sys, numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from oracle import oracle as orc
from tests import torch_spec
torch.set_num_threads(8)
R, S = 64, 128
ocfg = orc.default_config(n_samples=S)
rays, pix = orc.synthetic_rays(R, width=800, height=800, seed=2024)
u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, R, S + 1) for lv in range(2)])
params = orc.init_params(ocfg, 7)
o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64", with_backward=False)
shapes = orc.layer_shapes(ocfg)
tp = torch.tensor(params.astype(np.float64))
tr = {k: torch.tensor(np.asarray(v, np.float64)) for k, v in rays.items()}
t = torch.tensor(o64["t_vals"][1].astype(np.float64))
mean, cov = torch_spec.cast_rays(t, tr["origins"], tr["directions"], tr["radii"])
ep = torch_spec.ipe(mean, cov, ocfg.deg_point).reshape(R * S, -1)
ed = torch_spec.dir_enc(tr["directions"], ocfg.deg_view)[:, None, :].expand(-1, S, -1).reshape(R * S, -1)
f32 = lambda x: x.to(torch.float32)
def bf(x): return f32(x).to(torch.bfloat16).to(torch.float64)
def h16(x): return f32(x).to(torch.float16).to(torch.float64)
def e5(x, sc): return (f32(x * sc)).to(torch.float8_e5m2).to(torch.float64) / sc
def e4(x, sc): return (f32(x * sc)).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float64) / sc
def prod(a, W, mode):
    a = f32(a).to(torch.float64)  # activations are fp32 in the epilogue
    W = f32(W).to(torch.float64)
    if mode == "f64": return a @ W.T
    if mode == "bf16x3":
        ah, wh = bf(a), bf(W); al, wl = bf(a - ah), bf(W - wh)
        return ah @ wh.T + al @ wh.T + ah @ wl.T
    if mode == "bf16": return bf(a) @ bf(W).T
    if mode == "f16": return h16(a) @ h16(W).T
    if mode == "f8c_gpu":  # exactly the scales of mlp_fused_split.cu REP = 1 (accumulator = 2^15 x the sum); saturating conversions
        sat16 = lambda x: f32(x).clamp(-65504, 65504).to(torch.float16).to(torch.float64)
        ah, wh = sat16(a * 32) / 32, sat16(W * 1024) / 1024
        al, wl = a - ah, W - wh
        return ah @ wh.T + (e4(al, 2.0 ** 9) @ e4(wh, 2.0 ** 6).T) + (e4(ah, 1.0) @ e4(wl, 2.0 ** 15).T)
    if mode == "f8c":  # constrained scales: the two factors of each correction product multiply to 1
        ah, wh = h16(a), h16(W); al, wl = a - ah, W - wh
        return ah @ wh.T + e5(al, 2.0 ** 4) @ e5(wh, 2.0 ** -4).T + e5(ah, 2.0 ** -8) @ e5(wl, 2.0 ** 8).T
    if mode.startswith("f16+"):
        ah, wh = h16(a), h16(W); al, wl = a - ah, W - wh
        q = {"e5m2": e5, "e4m3": e4}[mode[4:]]
        # residuals scaled into range (exact powers of two), hi copies in fp8
        sa, sw = 2.0 ** 10, 2.0 ** 14
        return ah @ wh.T + q(al, sa) @ q(wh, 2.0 ** 4).T + q(ah, 1.0) @ q(wl, sw).T
    raise ValueError(mode)
def mlp(mode):
    Ws, bs = torch_spec.layer_views(ocfg, shapes, tp)
    D = ocfg.net_depth
    h = ep; zs = []
    for i in range(D):
        x = torch.cat([h, ep], -1) if (ocfg.skip_layer > 0 and i % ocfg.skip_layer == 0 and i > 0) else h
        z = prod(x, Ws[i], mode) + bs[i]; zs.append(z); h = torch.relu(z)
    rd = (h @ Ws[D].T + bs[D])[..., 0]
    c = torch.relu(prod(torch.cat([h, ed], -1), Ws[D + 1], mode) + bs[D + 1])
    rr = c @ Ws[D + 2].T + bs[D + 2]
    dens = torch.nn.functional.softplus(rd + ocfg.density_bias).reshape(R, S)
    rgb = (torch.sigmoid(rr) * (1 + 2 * ocfg.rgb_padding) - ocfg.rgb_padding).reshape(R, S, 3)
    comp, acc, w = torch_spec.render(rgb, dens, t, tr["directions"], bool(ocfg.white_bkgd))
    return comp, w, zs
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
c0, w0, z0 = mlp("f64")
WSCALE = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
if WSCALE != 1.0:
    nW = sum(o * (a + b) for o, a, b in zip(*shapes))
    tp = tp.clone(); tp[:nW] *= WSCALE
c0, w0, z0 = mlp("f64")
print("weight scale", WSCALE, "act max per layer", [round(float(torch.relu(z).max()), 2) for z in z0])
for mode in ("bf16x3", "f8c_gpu", "f8c", "f16"):
    c, w, zs = mlp(mode)
    flips = sum(int(((a > 0) != (b > 0)).sum()) for a, b in zip(zs, z0)); n = sum(a.numel() for a in z0)
    print(f"{mode:10s} comp_rgb {rel(c, c0):.2e} weights {rel(w, w0):.2e} z7 {rel(zs[-1], z0[-1]):.2e} mask flips {flips / n:.2e}")
