"""Independent torch-fp64 statement of SURVEY Appendix B (pin P3): used only to check the oracle's
gradients by autograd.  Not a product path."""
import torch


def layer_views(cfg, shapes, params):
    out, ia, ib = shapes
    L = len(out)
    Ws, bs, off = [], [], 0
    for l in range(L):
        n = out[l] * (ia[l] + ib[l])
        Ws.append(params[off:off + n].view(out[l], ia[l] + ib[l]))
        off += n
    for l in range(L):
        bs.append(params[off:off + out[l]])
        off += out[l]
    return Ws, bs


def mlp(cfg, shapes, params, enc_pos, enc_dir, masks=None, masks_out=None):
    """masks (optional): one bool tensor per hidden layer; the layer then computes z * mask instead of relu(z) — the
    network as a path that took exactly those ReLU branches (what a gradient under imposed masks needs).
    masks_out (optional list): receives z > 0 of every hidden layer."""
    Ws, bs = layer_views(cfg, shapes, params)
    D, Cn = cfg.net_depth, cfg.net_depth_condition

    def act(z, i):
        if masks_out is not None:
            masks_out.append(z.detach() > 0)
        return torch.relu(z) if masks is None else z * masks[i].reshape(z.shape).to(z.dtype)

    h = enc_pos
    for i in range(D):
        x = torch.cat([h, enc_pos], -1) if (cfg.skip_layer > 0 and i % cfg.skip_layer == 0 and i > 0) else h
        h = act(x @ Ws[i].T + bs[i], i)
    raw_density = (h @ Ws[D].T + bs[D])[..., 0]
    c = torch.cat([h, enc_dir], -1)
    for i in range(Cn):
        c = act(c @ Ws[D + 1 + i].T + bs[D + 1 + i], D + i)
    raw_rgb = c @ Ws[D + Cn + 1].T + bs[D + Cn + 1]
    return raw_density, raw_rgb


def cast_rays(t, o, d, radii):
    t0, t1 = t[:, :-1], t[:, 1:]
    mu, hw = (t0 + t1) / 2, (t1 - t0) / 2
    den = 3 * mu**2 + hw**2
    t_mean = mu + 2 * mu * hw**2 / den
    t_var = hw**2 / 3 - (4 / 15) * (hw**4 * (12 * mu**2 - hw**2)) / den**2
    r_var = radii[:, None] ** 2 * (mu**2 / 4 + (5 / 12) * hw**2 - (4 / 15) * hw**4 / den)
    mean = o[:, None, :] + d[:, None, :] * t_mean[..., None]
    dd = d * d
    dmag = torch.clamp(dd.sum(-1, keepdim=True), min=1e-10)
    cov = t_var[..., None] * dd[:, None, :] + r_var[..., None] * (1 - dd / dmag)[:, None, :]
    return mean, cov


def ipe(mean, cov, deg):
    sc = 2.0 ** torch.arange(deg, dtype=mean.dtype, device=mean.device)
    y = mean[..., None, :] * sc[:, None]
    yv = cov[..., None, :] * (sc**2)[:, None]
    e = torch.exp(-0.5 * yv)
    return torch.cat([e * torch.sin(y), e * torch.cos(y)], -1).flatten(-2)


def dir_enc(d, deg):
    outs = [d]
    for j in range(deg):
        outs += [torch.sin(d * 2.0**j), torch.cos(d * 2.0**j)]
    return torch.cat(outs, -1)


def render(rgb, density, t, d, white=True):
    delta = t[:, 1:] - t[:, :-1]
    alpha = 1 - torch.exp(-density * delta * d.norm(dim=-1, keepdim=True))
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha[:, :-1]], 1), 1)
    w = alpha * T
    comp = (w[..., None] * rgb).sum(1)
    acc = w.sum(1)
    if white:
        comp = comp + (1 - acc)[:, None]
    return comp, acc, w


def total_loss(cfg, shapes, params, rays, pixels, t_levels, masks_levels=None, masks_out=None):
    """t_levels: list of [R,S+1] tensors (treated as constants: no gradient through sampling).
    masks_levels / masks_out: per level, the `masks` / `masks_out` of mlp()."""
    o, d = rays["origins"], rays["directions"]
    lm = rays["loss_mults"]
    loss = 0.0
    comps = []
    for lv, t in enumerate(t_levels):
        mean, cov = cast_rays(t, o, d, rays["radii"])
        ep = ipe(mean, cov, cfg.deg_point)
        ed = dir_enc(d, cfg.deg_view)[:, None, :].expand(-1, ep.shape[1], -1)
        mo = [] if masks_out is not None else None
        rd, rr = mlp(cfg, shapes, params, ep, ed, None if masks_levels is None else masks_levels[lv], mo)
        if masks_out is not None:
            masks_out.append(mo)
        density = torch.nn.functional.softplus(rd + cfg.density_bias)
        rgb = torch.sigmoid(rr) * (1 + 2 * cfg.rgb_padding) - cfg.rgb_padding
        comp, acc, w = render(rgb, density, t, d, bool(cfg.white_bkgd))
        mult = cfg.coarse_loss_mult if lv < len(t_levels) - 1 else 1.0
        loss = loss + mult * (lm * ((comp - pixels) ** 2).sum(-1)).sum() / lm.sum()
        comps.append(comp)
    return loss, comps


def relu_margin(cfg, shapes, params, enc_pos, enc_dir):
    """Per-sample min |z| over every hidden (ReLU) unit, in fp64.  A forward perturbation eps flips the ReLU mask of
    the units with |z| < eps; the gradient is discontinuous there, so gradient parity is only well defined on samples
    whose margin exceeds the forward error of the path under test."""
    Ws, bs = layer_views(cfg, shapes, params)
    D, Cn = cfg.net_depth, cfg.net_depth_condition
    h = enc_pos
    margin = torch.full((enc_pos.shape[0],), float("inf"), dtype=enc_pos.dtype)
    for i in range(D):
        x = torch.cat([h, enc_pos], -1) if (cfg.skip_layer > 0 and i % cfg.skip_layer == 0 and i > 0) else h
        z = x @ Ws[i].T + bs[i]
        margin = torch.minimum(margin, z.abs().min(-1).values)
        h = torch.relu(z)
    c = torch.cat([h, enc_dir], -1)
    for i in range(Cn):
        z = c @ Ws[D + 1 + i].T + bs[D + 1 + i]
        margin = torch.minimum(margin, z.abs().min(-1).values)
        c = torch.relu(z)
    return margin
