"""CPU experiment: whole-step gradient error if wgrad consumes rounded X / dZ (fp64 everything else)."""
import sys, numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from oracle import oracle as orc
from tests import torch_spec
torch.set_num_threads(8)
R, S = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 128
ocfg = orc.default_config(n_samples=S)
rays, pix = orc.synthetic_rays(R, width=800, height=800, seed=2024)
rays["loss_mults"] = np.random.default_rng(1).uniform(0.5, 1.5, R).astype(np.float32)
u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, R, S + 1) for lv in range(2)])
params = orc.init_params(ocfg, 7)
o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64")
t_levels = [torch.tensor(t.astype(np.float64)) for t in o64["t_vals"]]

MODE = {"x": None, "dz": None, "scale": 1.0}
def q(t, kind, scale=1.0):
    if kind is None: return t
    if kind == "f16": return (t * scale).to(torch.float32).to(torch.float16).to(torch.float64) / scale
    if kind == "bf16": return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)
    if kind == "bf16x2":
        f = t.to(torch.float32); hi = f.to(torch.bfloat16); lo = (f - hi.to(torch.float32)).to(torch.bfloat16)
        return hi.to(torch.float64) + lo.to(torch.float64)
    if kind == "f32": return t.to(torch.float32).to(torch.float64)
    raise ValueError(kind)
stats = {}
class QLin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return x @ W.T + b
    @staticmethod
    def backward(ctx, dz):
        x, W = ctx.saved_tensors
        x2, dz2 = x.reshape(-1, x.shape[-1]), dz.reshape(-1, dz.shape[-1])
        stats.setdefault("dzmax", []).append(float(dz2.abs().max())); stats.setdefault("xmax", []).append(float(x2.abs().max()))
        dW = q(dz2, MODE["dz"], MODE["scale"]).T @ q(x2, MODE["x"])
        return dz @ W, dW, dz2.sum(0)
# patch torch_spec.mlp's matmuls by monkeypatching: re-implement mlp with QLin
def mlp(cfg, shapes, params, enc_pos, enc_dir, masks=None, masks_out=None):
    Ws, bs = torch_spec.layer_views(cfg, shapes, params)
    D, Cn = cfg.net_depth, cfg.net_depth_condition
    h = enc_pos
    for i in range(D):
        x = torch.cat([h, enc_pos], -1) if (cfg.skip_layer > 0 and i % cfg.skip_layer == 0 and i > 0) else h
        h = torch.relu(QLin.apply(x, Ws[i], bs[i]))
    rd = QLin.apply(h, Ws[D], bs[D])[..., 0]
    c = torch.cat([h, enc_dir], -1)
    for i in range(Cn):
        c = torch.relu(QLin.apply(c, Ws[D + 1 + i], bs[D + 1 + i]))
    return rd, QLin.apply(c, Ws[D + Cn + 1], bs[D + Cn + 1])
torch_spec.mlp = mlp
def grad():
    tp = torch.tensor(params.astype(np.float64), requires_grad=True)
    tr = {k: torch.tensor(np.asarray(v, np.float64)) for k, v in rays.items()}
    loss, _ = torch_spec.total_loss(ocfg, orc.layer_shapes(ocfg), tp, tr, torch.tensor(pix.astype(np.float64)), t_levels)
    (g,) = torch.autograd.grad(loss, tp)
    return g.numpy()
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
g0 = grad()
print("dz max per layer (backward order)", ["%.2e" % v for v in stats["dzmax"]])
print("x max", ["%.2e" % v for v in stats["xmax"]])
print("vs oracle f64:", rel(g0, o64["grads"]))
sizes = [o * (a + b) for o, a, b in zip(*orc.layer_shapes(ocfg))] + list(orc.layer_shapes(ocfg)[0])
for x, dz, sc in [("f16", "f16", 4096.0), ("bf16x2", "f16", 4096.0), ("f16", "bf16x2", 1.0), ("bf16x2", "bf16x2", 1.0)]:
    MODE.update(x=x, dz=dz, scale=sc)
    g = grad()
    worst, off = 0, 0
    per = []
    for n in sizes:
        e = rel(g[off:off+n], g0[off:off+n]); per.append(e)
        worst = max(worst, e); off += n
    print("  per tensor:", " ".join("%.1e" % e for e in per[:12]))
    print(f"x={x} dz={dz} scale={sc}: whole {rel(g, g0):.2e} worst tensor {worst:.2e}")
