/*
 * oracle_impl.h — body of the CPU oracle, included twice by oracle.c with
 *   REAL = float  / SUF = _f32   (reference operation order, -ffp-contract=off)
 *   REAL = double / SUF = _f64   (shadow used to arbitrate tolerances)
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  Citations: ANU/ = ScratchNerf/AcceleratedNeRFUtils/,
 * SN/ = ScratchNerf/ScratchNerf/, ".cu" = ANU/accelerated_functions.cu.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ------------------------------------------------------------------ sampling (B.3) */

/* Level 0: stratified samples between bin mid-points.
 * Intent of .cu:222-242 and SN/MipHelpers.cs:611-631 (mip-NeRF sample_along_rays):
 *   s_i = near*(1-i/S) + far*(i/S); mids m_i; lower=[s_0,m..], upper=[m..,s_S];
 *   t_i = lower_i + u_i*(upper_i-lower_i),  i = 0..S   (A-D8, A-D16). */
void FN(orc_sample_t_vals)(const REAL* nears, const REAL* fars, const float* u, int R, int S,
                           int randomized, REAL* t) {
  for (int r = 0; r < R; r++) {
    const REAL nr = nears[r], fr = fars[r];
    REAL* tr = t + (long)r * (S + 1);
    for (int i = 0; i <= S; i++) {
      REAL a = (REAL)i / (REAL)S;             /* SN/MipHelpers.cs:616 */
      tr[i] = nr * ((REAL)1 - a) + fr * a;    /* SN/MipHelpers.cs:622 == .cu:233 */
    }
    if (!randomized) continue;
    /* in-place would destroy s_i; use the definition directly */
    REAL prev_s = tr[0], s0 = tr[0], sS = tr[S];
    REAL prev_mid = 0;
    for (int i = 0; i <= S; i++) {
      REAL cur = tr[i];
      REAL next = i < S ? tr[i + 1] : cur;
      REAL mid_hi = i < S ? (REAL)0.5 * (cur + next) : sS;   /* upper_i */
      REAL lower = i == 0 ? s0 : prev_mid;                    /* lower_i */
      (void)prev_s;
      REAL ui = (REAL)u[(long)r * (S + 1) + i];
      REAL val = lower + (mid_hi - lower) * ui;               /* SN/MipHelpers.cs:629 */
      prev_mid = mid_hi;
      prev_s = cur;
      tr[i] = val; /* safe: tr[i+1] not yet modified, tr[i] no longer needed */
    }
  }
}

/* Level 1: blur-pool the weights, build the CDF, invert it at S+1 stratified u's.
 * SN/MipHelpers.cs:634-666 (ResampleAlongRay) + :774-851 (SortedPiecewiseConstantPDF).
 * The CUDA variant (.cu:246-291) is defective (A-D9); this follows the C#. */
void FN(orc_resample_t_vals)(const REAL* t, const REAL* w, const float* u, int R, int S,
                             REAL padding, int randomized, REAL* t_new) {
  const int nb = S;      /* numBins */
  const int ns = S + 1;  /* numSamples = tVals.Length (SN/MipHelpers.cs:663) */
  REAL* blur = (REAL*)malloc(sizeof(REAL) * (nb + 1) * 2);
  REAL* cdf = blur + nb;
  for (int r = 0; r < R; r++) {
    const REAL* wr = w + (long)r * S;
    const REAL* tr = t + (long)r * (S + 1);
    /* SN/MipHelpers.cs:645-661: pad with edge values, pairwise max, average, + padding */
    for (int i = 0; i < nb; i++) {
      REAL wl = wr[i > 0 ? i - 1 : 0], wc = wr[i], wh = wr[i < nb - 1 ? i + 1 : nb - 1];
      REAL m0 = wl > wc ? wl : wc, m1 = wc > wh ? wc : wh;
      blur[i] = (REAL)0.5 * (m0 + m1) + padding;
    }
    /* SN/MipHelpers.cs:784-796 */
    REAL sum = 0;
    for (int i = 0; i < nb; i++) sum += blur[i];
    REAL pad = (REAL)1e-5 - sum;
    if (pad > 0) {
      REAL per = pad / (REAL)nb;
      for (int i = 0; i < nb; i++) blur[i] += per;
      sum += pad;
    }
    /* SN/MipHelpers.cs:799-812: cdf = [0, min(1, cumsum(pdf[:-1])), 1] */
    cdf[0] = 0;
    REAL cum = 0;
    for (int i = 0; i < nb - 1; i++) {
      cum += blur[i] / sum;
      cdf[i + 1] = cum < (REAL)1 ? cum : (REAL)1;
    }
    cdf[nb] = 1;
    const REAL s1 = (REAL)1 / (REAL)ns;
    for (int s = 0; s < ns; s++) {
      REAL us;
      if (randomized) { /* SN/MipHelpers.cs:819 */
        us = (REAL)s * s1 + (REAL)u[(long)r * ns + s] * (s1 - (REAL)1e-7);
        REAL cap = (REAL)1 - (REAL)1e-7;
        if (us > cap) us = cap;
      } else { /* mip-NeRF: linspace(0, 1-eps, ns); the C# ignores `randomized` */
        us = (REAL)s * (((REAL)1 - (REAL)1.1920929e-7) / (REAL)(ns - 1));
      }
      /* SN/MipHelpers.cs:827-832: largest idx with cdf[idx] <= u, clamped to [0, nb-1] */
      int lo = 0, hi = nb + 1; /* count of entries <= us in cdf[0..nb] */
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] <= us) lo = mid + 1; else hi = mid;
      }
      int idx = lo - 1;
      if (idx < 0) idx = 0;
      if (idx > nb - 1) idx = nb - 1;
      REAL b0 = tr[idx], b1 = tr[idx + 1], c0 = cdf[idx], c1 = cdf[idx + 1];
      REAL den = c1 - c0;
      REAL tt = den > 0 ? (us - c0) / den : (REAL)0; /* SN/MipHelpers.cs:844 */
      if (tt < 0) tt = 0;
      if (tt > 1) tt = 1;
      t_new[(long)r * ns + s] = b0 + tt * (b1 - b0);  /* SN/MipHelpers.cs:847 */
    }
  }
  free(blur);
}

/* ------------------------------------------------------------------ frustum -> Gaussian (B.1) */

/* .cu:292-317, same arithmetic order; == SN/MipHelpers.cs:391-401 + :367-379. */
void FN(orc_cast_rays)(const REAL* t, const REAL* o, const REAL* d, const REAL* radii, int R,
                       int S, REAL* mean, REAL* cov) {
  for (int r = 0; r < R; r++) {
    const REAL dx = d[r * 3], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
    const REAL ox = o[r * 3], oy = o[r * 3 + 1], oz = o[r * 3 + 2];
    const REAL radius = radii[r];
    REAL dmag = dx * dx + dy * dy + dz * dz; /* .cu:311 */
    if (dmag < (REAL)1e-10) dmag = (REAL)1e-10;
    const REAL ddx = dx * dx, ddy = dy * dy, ddz = dz * dz;
    const REAL nx = 1 - ddx / dmag, ny = 1 - ddy / dmag, nz = 1 - ddz / dmag;
    for (int s = 0; s < S; s++) {
      const REAL t0 = t[(long)r * (S + 1) + s], t1 = t[(long)r * (S + 1) + s + 1];
      const REAL mu = (t0 + t1) / 2, hw = (t1 - t0) / 2;
      const REAL mu2 = mu * mu, hw2 = hw * hw;
      const REAL den = 3 * mu2 + hw2;
      const REAL t_mean = mu + 2 * mu * hw2 / den; /* .cu:306 */
      const REAL t_var = hw2 / 3 - (REAL)4 / (REAL)15 * (hw2 * hw2 * (12 * mu2 - hw2)) / (den * den);
      const REAL r_var =
          radius * radius * (mu2 / 4 + (REAL)5 / (REAL)12 * hw2 - (REAL)4 / (REAL)15 * (hw2 * hw2) / den);
      long m = ((long)r * S + s) * 3;
      mean[m] = dx * t_mean + ox;
      mean[m + 1] = dy * t_mean + oy;
      mean[m + 2] = dz * t_mean + oz;
      cov[m] = t_var * ddx + r_var * nx;
      cov[m + 1] = t_var * ddy + r_var * ny;
      cov[m + 2] = t_var * ddz + r_var * nz;
    }
  }
}

/* ------------------------------------------------------------------ encodings (B.2) */

/* .cu:194-204: enc[f*6 + a] = exp(-.5*var_a*4^f) * sin(mean_a*2^f), enc[f*6+3+a] = ... cos. */
void FN(orc_encode_position)(const REAL* mean, const REAL* cov, long M, int deg, REAL* enc) {
  const int P = 6 * deg;
#pragma omp parallel for schedule(static)
  for (long m = 0; m < M; m++) {
    for (int f = 0; f < deg; f++) {
      const REAL scale = (REAL)(1 << f); /* .cu:196 */
      for (int a = 0; a < 3; a++) {
        const REAL x = mean[m * 3 + a] * scale;
        const REAL yv = cov[m * 3 + a] * scale * scale; /* .cu:198 */
        const REAL e = R_EXP((REAL)-0.5 * yv);           /* .cu:185 */
        enc[m * P + f * 6 + a] = e * R_SIN(x);
        enc[m * P + f * 6 + 3 + a] = e * R_COS(x);       /* .cu:186 (true cos, A-D18) */
      }
    }
  }
}

/* SN/MipHelpers.cs:337-356 flattened as SN/MLP.cs:99-102: [d, sin(2^0 d), cos(2^0 d), ...]. */
void FN(orc_encode_direction)(const REAL* d, int R, int deg, REAL* enc) {
  const int Dd = 3 + 6 * deg;
  for (int r = 0; r < R; r++) {
    REAL* e = enc + (long)r * Dd;
    for (int a = 0; a < 3; a++) e[a] = d[r * 3 + a];
    for (int j = 0; j < deg; j++) {
      const REAL scale = (REAL)(1 << j);
      for (int a = 0; a < 3; a++) {
        const REAL xb = d[r * 3 + a] * scale;
        e[3 + j * 6 + a] = R_SIN(xb);
        e[3 + j * 6 + 3 + a] = R_COS(xb);
      }
    }
  }
}

/* ------------------------------------------------------------------ MLP */

typedef struct FN(mlp_ctx) {
  int L, D, C, W, Wc, P, Dd;
  int out[64], in_a[64], in_b[64];
  long w_off[64], b_off[64];
  long n_params;
  int act_stride;
  REAL* wt[64]; /* transposed weights [in][out] for the forward (same per-output k-ascending order) */
} FN(mlp_ctx);

static void FN(mlp_ctx_init)(FN(mlp_ctx) * x, const orc_config* c, const REAL* params) {
  x->D = c->net_depth; x->C = c->net_depth_condition; x->W = c->net_width;
  x->Wc = c->net_width_condition; x->P = 6 * c->deg_point; x->Dd = 3 + 6 * c->deg_view;
  x->L = orc_num_layers(c);
  orc_layer_shapes(c, x->out, x->in_a, x->in_b);
  long off = 0;
  for (int l = 0; l < x->L; l++) { x->w_off[l] = off; off += (long)x->out[l] * (x->in_a[l] + x->in_b[l]); }
  for (int l = 0; l < x->L; l++) { x->b_off[l] = off; off += x->out[l]; }
  x->n_params = off;
  x->act_stride = x->D * x->W + x->C * x->Wc;
  for (int l = 0; l < x->L; l++) {
    x->wt[l] = NULL;
    if (!params) continue;
    int in = x->in_a[l] + x->in_b[l], out = x->out[l];
    x->wt[l] = (REAL*)malloc(sizeof(REAL) * (size_t)in * out);
    for (int j = 0; j < out; j++)
      for (int k = 0; k < in; k++) x->wt[l][(long)k * out + j] = params[x->w_off[l] + (long)j * in + k];
  }
}
static void FN(mlp_ctx_free)(FN(mlp_ctx) * x) {
  for (int l = 0; l < x->L; l++) free(x->wt[l]);
}

/* Z = [xa|xb] W^T + b with each output accumulated k-ascending, bias last
 * (SN/MLP.cs:187-191 == .cu:42-45, .cu:82-87). Vectorised across outputs via W^T. */
static inline void FN(dense)(const REAL* wt, const REAL* b, const REAL* xa, int na, const REAL* xb,
                             int nb, int out, REAL* z) {
  for (int j = 0; j < out; j++) z[j] = 0;
  for (int k = 0; k < na; k++) {
    const REAL xv = xa[k];
    const REAL* wr = wt + (long)k * out;
    for (int j = 0; j < out; j++) z[j] += xv * wr[j];
  }
  for (int k = 0; k < nb; k++) {
    const REAL xv = xb[k];
    const REAL* wr = wt + (long)(na + k) * out;
    for (int j = 0; j < out; j++) z[j] += xv * wr[j];
  }
  for (int j = 0; j < out; j++) z[j] += b[j];
}

/* One sample forward. acts: [D*W + C*Wc] post-ReLU outputs. SN/MLP.cs:112-136. */
static void FN(mlp_fwd_sample)(const FN(mlp_ctx) * x, const REAL* params, const REAL* x0,
                               const REAL* xdir, REAL* acts, REAL* raw_density, REAL* raw_rgb) {
  const int D = x->D, C = x->C, W = x->W, Wc = x->Wc;
  const REAL* h = x0;
  for (int i = 0; i < D; i++) {
    REAL* o = acts + (long)i * W;
    FN(dense)(x->wt[i], params + x->b_off[i], h, x->in_a[i], x0, x->in_b[i], W, o);
    for (int j = 0; j < W; j++) o[j] = o[j] > 0 ? o[j] : 0; /* ReLU, SN/MLP.cs:223 */
    h = o;
  }
  FN(dense)(x->wt[D], params + x->b_off[D], h, W, NULL, 0, 1, raw_density); /* SN/MLP.cs:123 */
  const REAL* ci = h;
  for (int i = 0; i < C; i++) {
    REAL* o = acts + (long)D * W + (long)i * Wc;
    int l = D + 1 + i;
    FN(dense)(x->wt[l], params + x->b_off[l], ci, x->in_a[l], xdir, x->in_b[l], Wc, o);
    for (int j = 0; j < Wc; j++) o[j] = o[j] > 0 ? o[j] : 0;
    ci = o;
  }
  FN(dense)(x->wt[D + C + 1], params + x->b_off[D + C + 1], ci, Wc, NULL, 0, 3, raw_rgb);
}

/* dW += dz (x) [xa|xb]; db += dz; din = W^T dz (first n_keep inputs only).
 * SN/MLP.cs:196-220 with act'(Z) (A-D16) == .cu:97-110. */
static inline void FN(dense_bwd)(const REAL* Wm, const REAL* dz, int out, const REAL* xa, int na,
                                 const REAL* xb, int nb, REAL* gW, REAL* gb, REAL* din, int n_keep) {
  const int in = na + nb;
  for (int k = 0; k < n_keep; k++) din[k] = 0;
  for (int j = 0; j < out; j++) {
    const REAL g = dz[j];
    if (g == 0) continue; /* adds exact zeros: skipping leaves results unchanged */
    gb[j] += g;
    REAL* gw = gW + (long)j * in;
    for (int k = 0; k < na; k++) gw[k] += g * xa[k];
    for (int k = 0; k < nb; k++) gw[na + k] += g * xb[k];
    const REAL* wr = Wm + (long)j * in;
    for (int k = 0; k < n_keep; k++) din[k] += g * wr[k];
  }
}

/* One sample backward, accumulating into grads. SN/MLP.cs:138-175. */
static void FN(mlp_bwd_sample)(const FN(mlp_ctx) * x, const REAL* params, const REAL* x0,
                               const REAL* xdir, const REAL* acts, REAL d_raw_density,
                               const REAL* d_raw_rgb, REAL* grads, REAL* scratch) {
  const int D = x->D, C = x->C, W = x->W, Wc = x->Wc;
  const int mx = (W > Wc ? W : Wc);
  REAL* dz = scratch;            /* [mx] */
  REAL* din = scratch + mx;      /* [mx] */
  REAL* dh = scratch + 2 * mx;   /* [mx] */
  const REAL* trunk_out = acts + (long)(D - 1) * W;
  /* RGB head (identity activation inside the MLP, SN/MLP.cs:143) */
  {
    int l = D + C + 1;
    const REAL* ci = acts + (long)D * W + (long)(C - 1) * Wc;
    FN(dense_bwd)(params + x->w_off[l], d_raw_rgb, 3, ci, Wc, NULL, 0, grads + x->w_off[l],
                  grads + x->b_off[l], din, Wc);
  }
  /* condition layers, reverse (SN/MLP.cs:144-147) */
  for (int i = C - 1; i >= 0; i--) {
    int l = D + 1 + i;
    const REAL* o = acts + (long)D * W + (long)i * Wc;
    for (int j = 0; j < Wc; j++) dz[j] = o[j] > 0 ? din[j] : 0;
    const REAL* ci = i == 0 ? trunk_out : acts + (long)D * W + (long)(i - 1) * Wc;
    int keep = i == 0 ? W : Wc; /* SN/MLP.cs:148: drop the direction part */
    FN(dense_bwd)(params + x->w_off[l], dz, Wc, ci, x->in_a[l], xdir, x->in_b[l],
                  grads + x->w_off[l], grads + x->b_off[l], din, keep);
  }
  for (int k = 0; k < W; k++) dh[k] = din[k];
  /* density head (SN/MLP.cs:149-153) */
  {
    REAL dzs = d_raw_density;
    FN(dense_bwd)(params + x->w_off[D], &dzs, 1, trunk_out, W, NULL, 0, grads + x->w_off[D],
                  grads + x->b_off[D], din, W);
    for (int k = 0; k < W; k++) dh[k] += din[k];
  }
  /* trunk, reverse (SN/MLP.cs:155-159) */
  for (int i = D - 1; i >= 0; i--) {
    const REAL* o = acts + (long)i * W;
    for (int j = 0; j < W; j++) dz[j] = o[j] > 0 ? dh[j] : 0;
    const REAL* hin = i == 0 ? x0 : acts + (long)(i - 1) * W;
    int keep = i == 0 ? 0 : W; /* encodings need no gradient (.cu:154-182) */
    FN(dense_bwd)(params + x->w_off[i], dz, W, hin, x->in_a[i], x0, x->in_b[i], grads + x->w_off[i],
                  grads + x->b_off[i], din, keep);
    for (int k = 0; k < keep; k++) dh[k] = din[k];
  }
}

void FN(orc_mlp_forward)(const orc_config* c, const REAL* params, const REAL* enc_pos,
                         const REAL* enc_dir, long M, REAL* acts, REAL* raw_density, REAL* raw_rgb) {
  FN(mlp_ctx) x;
  FN(mlp_ctx_init)(&x, c, params);
#pragma omp parallel
  {
    REAL* tmp = (REAL*)malloc(sizeof(REAL) * x.act_stride);
#pragma omp for schedule(static)
    for (long m = 0; m < M; m++) {
      REAL* a = acts ? acts + m * x.act_stride : tmp;
      FN(mlp_fwd_sample)(&x, params, enc_pos + m * x.P, enc_dir + m * x.Dd, a, raw_density + m,
                         raw_rgb + m * 3);
    }
    free(tmp);
  }
  FN(mlp_ctx_free)(&x);
}

void FN(orc_mlp_backward)(const orc_config* c, const REAL* params, const REAL* enc_pos,
                          const REAL* enc_dir, const REAL* acts, const REAL* d_raw_density,
                          const REAL* d_raw_rgb, long M, REAL* grads) {
  FN(mlp_ctx) x;
  FN(mlp_ctx_init)(&x, c, NULL);
  const int nt = omp_get_max_threads();
  REAL* tl = (REAL*)calloc((size_t)nt * x.n_params, sizeof(REAL));
#pragma omp parallel num_threads(nt)
  {
    const int tid = omp_get_thread_num();
    REAL* scratch = (REAL*)malloc(sizeof(REAL) * 3 * (x.W > x.Wc ? x.W : x.Wc));
#pragma omp for schedule(static)
    for (long m = 0; m < M; m++)
      FN(mlp_bwd_sample)(&x, params, enc_pos + m * x.P, enc_dir + m * x.Dd, acts + m * x.act_stride,
                         d_raw_density[m], d_raw_rgb + m * 3, tl + (long)tid * x.n_params, scratch);
    free(scratch);
  }
  for (int t = 0; t < nt; t++) /* fixed reduction order */
    for (long i = 0; i < x.n_params; i++) grads[i] += tl[(long)t * x.n_params + i];
  free(tl);
}

/* ------------------------------------------------------------------ output activations */

static inline REAL FN(sigmoid)(REAL v) { return (REAL)1 / ((REAL)1 + R_EXP(-v)); } /* .cu:9 */
static inline REAL FN(softplus)(REAL v) { return R_LOG((REAL)1 + R_EXP(v)); }      /* .cu:14 */

/* density = softplus(raw + bias); rgb = sigmoid(raw)*(1+2pad) - pad. SN/MipNerfModel.cs:81-83. */
void FN(orc_output_activations)(const orc_config* c, const REAL* raw_density, const REAL* raw_rgb,
                                long M, REAL* density, REAL* rgb) {
  const REAL bias = (REAL)c->density_bias, pad = (REAL)c->rgb_padding;
  for (long m = 0; m < M; m++) {
    density[m] = FN(softplus)(raw_density[m] + bias);
    for (int a = 0; a < 3; a++)
      rgb[m * 3 + a] = FN(sigmoid)(raw_rgb[m * 3 + a]) * (1 + 2 * pad) - pad;
  }
}
/* SN/MipNerfModel.cs:184-189; .cu:120 (sigmoid'), .cu:141 (softplus' = sigmoid). */
void FN(orc_output_activations_grad)(const orc_config* c, const REAL* raw_density,
                                     const REAL* raw_rgb, const REAL* d_density, const REAL* d_rgb,
                                     long M, REAL* d_raw_density, REAL* d_raw_rgb) {
  const REAL bias = (REAL)c->density_bias, pad = (REAL)c->rgb_padding;
  for (long m = 0; m < M; m++) {
    d_raw_density[m] = d_density[m] * FN(sigmoid)(raw_density[m] + bias);
    for (int a = 0; a < 3; a++) {
      REAL s = FN(sigmoid)(raw_rgb[m * 3 + a]);
      d_raw_rgb[m * 3 + a] = d_rgb[m * 3 + a] * (s * (1 - s)) * (1 + 2 * pad);
    }
  }
}

/* ------------------------------------------------------------------ compositing (B.4, B.5) */

/* .cu:318-344 over all S samples; depth/acc per SN/MipHelpers.cs:488-490 (A-D11). */
void FN(orc_volumetric_rendering)(const REAL* rgb, const REAL* density, const REAL* t, const REAL* d,
                                  int R, int S, int white_bkgd, REAL* comp_rgb, REAL* depth,
                                  REAL* acc, REAL* weights, REAL* alpha, REAL* transmittance) {
  for (int r = 0; r < R; r++) {
    const REAL dl = R_SQRT(d[r * 3] * d[r * 3] + d[r * 3 + 1] * d[r * 3 + 1] + d[r * 3 + 2] * d[r * 3 + 2]);
    const REAL* tr = t + (long)r * (S + 1);
    REAL cr = 0, cg = 0, cb = 0, a_sum = 0, wd = 0, T = 1, prev_alpha = 0;
    for (int i = 0; i < S; i++) {
      const long ix = (long)r * S + i;
      const REAL al = 1 - R_EXP(-density[ix] * (tr[i + 1] - tr[i]) * dl); /* .cu:330 */
      T = i == 0 ? (REAL)1 : T * (1 - prev_alpha);                          /* .cu:331 */
      const REAL w = al * T;                                                /* .cu:332 */
      cr += w * rgb[ix * 3]; cg += w * rgb[ix * 3 + 1]; cb += w * rgb[ix * 3 + 2];
      a_sum += w;
      wd += w * (tr[i] + tr[i + 1]) / 2; /* SN/MipHelpers.cs:488 */
      if (weights) weights[ix] = w;
      if (alpha) alpha[ix] = al;
      if (transmittance) transmittance[ix] = T;
      prev_alpha = al;
    }
    if (white_bkgd) { cr += 1 - a_sum; cg += 1 - a_sum; cb += 1 - a_sum; } /* .cu:338-340 */
    comp_rgb[r * 3] = cr; comp_rgb[r * 3 + 1] = cg; comp_rgb[r * 3 + 2] = cb;
    if (depth) { /* SN/MipHelpers.cs:490 */
      REAL dv = a_sum > 0 ? wd / a_sum : (REAL)INFINITY;
      if (dv < tr[0]) dv = tr[0];
      if (dv > tr[S]) dv = tr[S];
      depth[r] = dv;
    }
    if (acc) acc[r] = a_sum;
  }
}

/* .cu:347-361, one independent buffer per level (A-D13). */
void FN(orc_output_gradient)(const REAL* comp_rgb, const REAL* pixels, const REAL* loss_mults, int R,
                             REAL loss_mult_sum, REAL level_mult, REAL* g) {
  for (int r = 0; r < R; r++)
    for (int a = 0; a < 3; a++)
      g[r * 3 + a] = 2 * loss_mults[r] / loss_mult_sum * (comp_rgb[r * 3 + a] - pixels[r * 3 + a]) * level_mult;
}

/* Reverse recurrences of .cu:379-401 / SN/MipHelpers.cs:565-596.
 * last_sample_mode 0: true gradient of the S-sample forward (A-D12);
 *                  1: the reference kernel — sample S-1 gets no gradient and its term is dropped. */
void FN(orc_volumetric_rendering_gradient)(const REAL* g, const REAL* rgb, const REAL* density,
                                           const REAL* t, const REAL* d, int R, int S, int white_bkgd,
                                           int last_sample_mode, REAL* d_rgb, REAL* d_density) {
  REAL* al = (REAL*)malloc(sizeof(REAL) * 3 * S);
  REAL* Tr = al + S;
  REAL* wt = al + 2 * S;
  for (int r = 0; r < R; r++) {
    const REAL dl = R_SQRT(d[r * 3] * d[r * 3] + d[r * 3 + 1] * d[r * 3 + 1] + d[r * 3 + 2] * d[r * 3 + 2]);
    const REAL* tr = t + (long)r * (S + 1);
    for (int i = 0; i < S; i++) { /* recompute the forward caches exactly as the forward does */
      const long ix = (long)r * S + i;
      al[i] = 1 - R_EXP(-density[ix] * (tr[i + 1] - tr[i]) * dl);
      Tr[i] = i == 0 ? (REAL)1 : Tr[i - 1] * (1 - al[i - 1]);
      wt[i] = al[i] * Tr[i];
    }
    const REAL gx = g[r * 3], gy = g[r * 3 + 1], gz = g[r * 3 + 2];
    const REAL dLdAcc = white_bkgd ? -(gx + gy + gz) : (REAL)0; /* .cu:370 */
    REAL dLdT_next = 0; /* dLdTransmittance[i+1]; .cu:375 */
    int start = S - 1;
    if (last_sample_mode == 1) {
      const long ix = (long)r * S + S - 1;
      d_rgb[ix * 3] = d_rgb[ix * 3 + 1] = d_rgb[ix * 3 + 2] = 0; d_density[ix] = 0;
      start = S - 2;
    }
    for (int i = start; i >= 0; i--) {
      const long ix = (long)r * S + i;
      const REAL dLdw = gx * rgb[ix * 3] + gy * rgb[ix * 3 + 1] + gz * rgb[ix * 3 + 2] + dLdAcc; /* .cu:385 */
      d_rgb[ix * 3] = gx * wt[i]; d_rgb[ix * 3 + 1] = gy * wt[i]; d_rgb[ix * 3 + 2] = gz * wt[i]; /* .cu:388 */
      const REAL dLdAlpha = dLdw * Tr[i] - dLdT_next * Tr[i];          /* .cu:390 */
      const REAL dLdT = dLdw * al[i] + dLdT_next * (1 - al[i]);        /* .cu:391 */
      const REAL dAlpha = (1 - al[i]) * (tr[i + 1] - tr[i]) * dl;      /* .cu:396-398 */
      d_density[ix] = dLdAlpha * dAlpha;                               /* .cu:400 */
      dLdT_next = dLdT;
    }
  }
  free(al);
}

/* ------------------------------------------------------------------ Adam (B.6) */

/* .cu:403-416 with the host part of ANU/AcceleratedAdamOptimizer.cpp:26-28 (eps_mode 0);
 * SN/TrainState.cs:25-37 (eps_mode 1). */
void FN(orc_adam_step)(REAL* p, const REAL* g, REAL* m, REAL* v, long n, REAL lr, int iteration,
                       int eps_mode) {
  const REAL b1 = (REAL)0.9, b2 = (REAL)0.999;
  const REAL inv1 = 1 / (1 - R_POW(b1, (REAL)iteration));
  const REAL inv2 = 1 / (1 - R_POW(b2, (REAL)iteration));
  for (long i = 0; i < n; i++) {
    const REAL gi = g[i];
    const REAL mi = b1 * m[i] + (1 - b1) * gi;
    const REAL vi = b2 * v[i] + (1 - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const REAL mh = mi * inv1, vh = vi * inv2;
    if (eps_mode == 0) p[i] -= lr * mh * ((REAL)1 / R_SQRT(vh + (REAL)1e-8)); /* .cu:415 */
    else p[i] -= lr * mh / (R_SQRT(vh) + (REAL)1e-8);                          /* SN/TrainState.cs:34 */
  }
}

/* ------------------------------------------------------------------ whole step */

/* SN/MipNerfModel.cs:99-200 (+ loss SN/Program.cs:64-89), one ray at a time: each ray's gradient
 * g = 2*lambda*lm/sum(lm)*(rgb - pix) depends only on that ray and the batch-wide sum(lm), so the
 * per-ray forward and backward can run back to back without the C#'s batch-wide caches. */
double FN(orc_train_gradient)(const orc_config* c, const REAL* params, const REAL* origins,
                              const REAL* dirs, const REAL* radii, const REAL* nears,
                              const REAL* fars, const REAL* loss_mults, const REAL* pixels,
                              const float* u, int R, int with_backward, REAL* grads, REAL* comp_rgb,
                              REAL* depth, REAL* acc, REAL* t_vals, REAL* weights, REAL* loss) {
  FN(mlp_ctx) x;
  FN(mlp_ctx_init)(&x, c, params);
  const int S = c->n_samples, NL = c->n_levels, P = x.P, Dd = x.Dd;
  REAL lm_sum = 0;
  for (int r = 0; r < R; r++) lm_sum += loss_mults[r]; /* float sum (A-D13) */
  const int nt = omp_get_max_threads();
  REAL* tl = with_backward ? (REAL*)calloc((size_t)nt * x.n_params, sizeof(REAL)) : NULL;
  REAL* rgb_all = (REAL*)malloc(sizeof(REAL) * (size_t)NL * R * 3);
#pragma omp parallel num_threads(nt)
  {
    const int tid = omp_get_thread_num();
    const int mx = (x.W > x.Wc ? x.W : x.Wc);
    REAL* buf = (REAL*)malloc(sizeof(REAL) * ((size_t)NL * ((S + 1) + S * (6 + P + x.act_stride + 4 + 4 + 1)) + Dd + 3 * mx + 8 * S));
    REAL* tv = buf;                              /* [NL][S+1] */
    REAL* mean = tv + NL * (S + 1);              /* [S*3] (reused per level) */
    REAL* cov = mean + S * 3;                    /* [S*3] */
    REAL* enc = cov + S * 3;                     /* [NL][S*P] */
    REAL* acts = enc + (size_t)NL * S * P;       /* [NL][S*act_stride] */
    REAL* raw = acts + (size_t)NL * S * x.act_stride; /* [NL][S*4]: raw density [S], raw rgb [S*3] */
    REAL* outv = raw + NL * S * 4;               /* [NL][S*4]: density [S], rgb [S*3] */
    REAL* wts = outv + NL * S * 4;               /* [NL][S] */
    REAL* xdir = wts + NL * S;                   /* [Dd] */
    REAL* scratch = xdir + Dd;                   /* [3*mx] */
    REAL* dtmp = scratch + 3 * mx;               /* [8*S] */
#pragma omp for schedule(static)
    for (int r = 0; r < R; r++) {
      FN(orc_encode_direction)(dirs + r * 3, 1, c->deg_view, xdir); /* raw Direction (A-D10) */
      for (int lv = 0; lv < NL; lv++) {
        REAL* tl_ = tv + lv * (S + 1);
        const float* ul = u + ((size_t)lv * R + r) * (S + 1);
        if (lv == 0) FN(orc_sample_t_vals)(nears + r, fars + r, ul, 1, S, c->randomized, tl_);
        else FN(orc_resample_t_vals)(tv + (lv - 1) * (S + 1), wts + (lv - 1) * S, ul, 1, S,
                                     (REAL)c->resample_padding, c->randomized, tl_);
        FN(orc_cast_rays)(tl_, origins + r * 3, dirs + r * 3, radii + r, 1, S, mean, cov);
        REAL* e = enc + (size_t)lv * S * P;
        for (int s = 0; s < S; s++) { /* serial IPE (the public one is omp-parallel) */
          for (int f = 0; f < c->deg_point; f++) {
            const REAL scale = (REAL)(1 << f);
            for (int a = 0; a < 3; a++) {
              const REAL xx = mean[s * 3 + a] * scale, yv = cov[s * 3 + a] * scale * scale;
              const REAL ee = R_EXP((REAL)-0.5 * yv);
              e[s * P + f * 6 + a] = ee * R_SIN(xx);
              e[s * P + f * 6 + 3 + a] = ee * R_COS(xx);
            }
          }
        }
        REAL* rw = raw + lv * S * 4;
        REAL* ov = outv + lv * S * 4;
        for (int s = 0; s < S; s++)
          FN(mlp_fwd_sample)(&x, params, e + s * P, xdir, acts + ((size_t)lv * S + s) * x.act_stride,
                             rw + s, rw + S + s * 3);
        FN(orc_output_activations)(c, rw, rw + S, S, ov, ov + S);
        REAL crgb[3], dep, ac;
        FN(orc_volumetric_rendering)(ov + S, ov, tl_, dirs + r * 3, 1, S, c->white_bkgd, crgb, &dep,
                                     &ac, wts + lv * S, NULL, NULL);
        for (int a = 0; a < 3; a++) rgb_all[((size_t)lv * R + r) * 3 + a] = crgb[a];
        if (comp_rgb) for (int a = 0; a < 3; a++) comp_rgb[((size_t)lv * R + r) * 3 + a] = crgb[a];
        if (depth) depth[(size_t)lv * R + r] = dep;
        if (acc) acc[(size_t)lv * R + r] = ac;
        if (t_vals) memcpy(t_vals + ((size_t)lv * R + r) * (S + 1), tl_, sizeof(REAL) * (S + 1));
        if (weights) memcpy(weights + ((size_t)lv * R + r) * S, wts + lv * S, sizeof(REAL) * S);
      }
      if (!with_backward) continue;
      for (int lv = NL - 1; lv >= 0; lv--) { /* SN/MipNerfModel.cs:171 */
        REAL g[3];
        const REAL mult = lv < NL - 1 ? (REAL)c->coarse_loss_mult : (REAL)1;
        FN(orc_output_gradient)(rgb_all + ((size_t)lv * R + r) * 3, pixels + r * 3, loss_mults + r, 1,
                                lm_sum, mult, g);
        REAL* ov = outv + lv * S * 4;
        REAL* rw = raw + lv * S * 4;
        REAL* d_rgb = dtmp, *d_den = dtmp + 3 * S, *d_raw_den = dtmp + 4 * S, *d_raw_rgb = dtmp + 5 * S;
        FN(orc_volumetric_rendering_gradient)(g, ov + S, ov, tv + lv * (S + 1), dirs + r * 3, 1, S,
                                              c->white_bkgd, c->last_sample_mode, d_rgb, d_den);
        FN(orc_output_activations_grad)(c, rw, rw + S, d_den, d_rgb, S, d_raw_den, d_raw_rgb);
        const REAL* e = enc + (size_t)lv * S * P;
        for (int s = 0; s < S; s++)
          FN(mlp_bwd_sample)(&x, params, e + s * P, xdir, acts + ((size_t)lv * S + s) * x.act_stride,
                             d_raw_den[s], d_raw_rgb + s * 3, tl + (size_t)tid * x.n_params, scratch);
      }
    }
    free(buf);
  }
  if (with_backward && grads) {
    for (long i = 0; i < x.n_params; i++) grads[i] = 0;
    for (int t = 0; t < nt; t++)
      for (long i = 0; i < x.n_params; i++) grads[i] += tl[(size_t)t * x.n_params + i];
  }
  /* loss — SN/Program.cs:64: sum(lm * |rgb - pix|^2) / sum(lm), per level; total per :81 */
  double total = 0;
  for (int lv = 0; lv < NL; lv++) {
    REAL acc_l = 0;
    for (int r = 0; r < R; r++) {
      REAL e2 = 0;
      for (int a = 0; a < 3; a++) {
        REAL df = rgb_all[((size_t)lv * R + r) * 3 + a] - pixels[r * 3 + a];
        e2 += df * df;
      }
      acc_l += loss_mults[r] * e2;
    }
    REAL l = acc_l / lm_sum;
    if (loss) loss[lv] = l;
    total += (lv < NL - 1 ? c->coarse_loss_mult : 1.0) * (double)l;
  }
  free(tl);
  free(rgb_all);
  FN(mlp_ctx_free)(&x);
  return total;
}

#undef FN
#undef CAT
#undef CAT_
