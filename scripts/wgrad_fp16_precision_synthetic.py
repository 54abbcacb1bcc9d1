"""CPU: per-tensor wgrad error with fp16-rounded operands on the synthetic test of tests/test_tc_gpu.py::test_tc_mlp_backward."""
import sys, numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from oracle import oracle as orc
from tests import torch_spec
torch.set_num_threads(4)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ocfg = orc.default_config(n_samples=64)
rng = np.random.default_rng(4)
params = orc.init_params(ocfg, 7)
nb_ = sum(orc.layer_shapes(ocfg)[0]); params[-nb_:] = np.random.default_rng(5).normal(size=nb_).astype(np.float32) * 0.1
P, Dd = 6 * ocfg.deg_point, 3 + 6 * ocfg.deg_view
ep = rng.uniform(-1, 1, (M, P)).astype(np.float32); ed = rng.uniform(-1, 1, (M, Dd)).astype(np.float32)
cg, dg = rng.normal(size=(M, 3)).astype(np.float32), rng.normal(size=M).astype(np.float32)
MODE = {"x": None, "dz": None}
def q(t, kind):
    if kind is None: return t
    if kind == "f16":
        s = 2.0 ** -np.floor(np.log2(float(t.abs().max())))
        return (t * s).to(torch.float32).to(torch.float16).to(torch.float64) / s
    if kind == "bf16x2":
        f = t.to(torch.float32); hi = f.to(torch.bfloat16); lo = (f - hi.to(torch.float32)).to(torch.bfloat16)
        return hi.to(torch.float64) + lo.to(torch.float64)
class QLin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W); return x @ W.T + b
    @staticmethod
    def backward(ctx, dz):
        x, W = ctx.saved_tensors
        return dz @ W, q(dz, MODE["dz"]).T @ q(x, MODE["x"]), dz.sum(0)
def mlp(cfg, shapes, params, enc_pos, enc_dir):
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
def grad():
    tp = torch.tensor(params.astype(np.float64), requires_grad=True)
    rd, rr = mlp(ocfg, orc.layer_shapes(ocfg), tp, torch.tensor(ep.astype(np.float64)), torch.tensor(ed.astype(np.float64)))
    den = torch.nn.functional.softplus(rd + ocfg.density_bias); rgb = torch.sigmoid(rr) * (1 + 2 * ocfg.rgb_padding) - ocfg.rgb_padding
    L = (den * torch.tensor(dg.astype(np.float64))).sum() + (rgb * torch.tensor(cg.astype(np.float64))).sum()
    (g,) = torch.autograd.grad(L, tp); return g.numpy()
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
g0 = grad()
sizes = [o * (a + b) for o, a, b in zip(*orc.layer_shapes(ocfg))]
for x, dz in [("f16", "f16"), ("bf16x2", "f16"), ("f16", "bf16x2")]:
    MODE.update(x=x, dz=dz); g = grad(); off = 0; per = []
    for n in sizes: per.append(rel(g[off:off+n], g0[off:off+n])); off += n
    print(f"M={M} x={x} dz={dz}: per tensor", " ".join("%.1e" % e for e in per))
