// Chain / job construction for every MLP entry point of the C ABI on SVS_ENGINE_TC (see mlp_tc.cuh).
#pragma once
#include "mlp_tc_host.cuh"

namespace svs {
namespace tc {

static void sdf_geometry(TcChain* ch, const svs_mlp_desc* d, const float* x, int64_t P, int64_t n_clamped) {
  ch->P = P;
  ch->x = x;
  ch->d_in = d->d_in;
  ch->n_freqs = d->n_freqs;
  ch->radius = d->sphere_radius;
  ch->sph_scale = d->sphere_scale;
  ch->n_clamped = n_clamped;
}

// forward layers 0..L-2 (softplus); saves h_{l+1} when `sv` is given
static void add_sdf_forward_steps(TcChain* ch, const Layout& lo, const WImages& wi, const float* wbuf, const SdfSaved* sv,
                                  bool split) {
  const uint8_t* reg = wimg_region(lo, wbuf);
  ch->split = split ? 1 : 0;
  ch->prologue = PRO_PE;
  ch->pro_kb = kb_of(lo.in[0]);
  if (sv) {
    ch->img[0] = timg(sv->A0);
    ch->pro_save = 0;
  }
  for (int l = 0; l < lo.L - 1; ++l) {
    TcStep s = make_step(reg + wi.fwd[l], wbuf + lo.boff[l], wi.fwd_kb[l], wi.fwd_npad[l], lo.out[l], EP_SOFTPLUS);
    s.next_kb = kb_of(lo.in[l + 1]);
    if (split) {
      s.w_lo = reg + wi.fwd_lo[l];
      s.bias_t = reinterpret_cast<const float*>(reg + wi.fwd_bt[l]);
    }
    if (l + 1 == lo.skip) {
      s.scale = kInvSqrt2;
      s.flags = TC_PEFILL;
    }
    if (sv) {
      ch->img[l + 1] = timg(sv->H[l + 1]);
      s.save = l + 1;
    }
    ch->st[ch->n_steps++] = s;
  }
}

// y (P, ldy) and/or clamped sdf (P) without saving anything: ImplicitNetwork.forward / get_sdf_vals under no_grad
static int sdf_forward(const svs_mlp_desc* d, const Layout& lo, const float* wbuf, const float* x, int64_t P, float* y,
                       float* sdf, cudaStream_t st, bool split) {
  WImages wi;
  SVS_TRY(make_wimages(d, lo, &wi, nullptr, nullptr, split));
  const uint8_t* reg = wimg_region(lo, wbuf);
  const int L = lo.L;
  TcChain ch;
  init_chain(&ch);
  sdf_geometry(&ch, d, x, P, P);
  add_sdf_forward_steps(&ch, lo, wi, wbuf, nullptr, split);
  ch.y = y;
  ch.ldy = lo.ldy;
  ch.sdf = sdf;
  if (sdf) {
    TcStep s0 = make_step(reg + wi.fwd_sdf, wbuf + lo.boff[L - 1], wi.fwd_kb[L - 1], 16, 1, EP_SDF);
    if (split) s0.w_lo = reg + wi.fwd_sdf_lo;
    ch.st[ch.n_steps++] = s0;
  }
  if (y) {
    TcStep s0 = make_step(reg + wi.fwd_sdf, wbuf + lo.boff[L - 1], wi.fwd_kb[L - 1], 16, 1, EP_Y);
    if (split) s0.w_lo = reg + wi.fwd_sdf_lo;
    ch.st[ch.n_steps++] = s0;
    TcStep s1 = make_step(reg + wi.fwd[L - 1], wbuf + lo.boff[L - 1] + 1, wi.fwd_kb[L - 1], wi.fwd_npad[L - 1], lo.out[L - 1] - 1, EP_Y);
    if (split) s1.w_lo = reg + wi.fwd_lo[L - 1];
    s1.y_col = 1;
    ch.st[ch.n_steps++] = s1;
  }
  return launch_chain(ch, "mlp_tc_sdf_fwd", chain_flops(ch), 0.0, st);
}

// get_outputs()/gradient(): y, clamped sdf, d sdf/dx; `saved` keeps A0, h_1..h_{L-1}, U_0..U_{L-2} for the backward
static int sdf_outputs_forward(const svs_mlp_desc* d, const Layout& lo, const float* wbuf, const float* x, int64_t P,
                               int64_t clamp, float* y, float* sdf, float* grad, void* saved, cudaStream_t st, bool split) {
  WImages wi;
  SVS_TRY(make_wimages(d, lo, &wi, nullptr, nullptr, split));
  const uint8_t* reg = wimg_region(lo, wbuf);
  const int L = lo.L;
  SdfSaved sv;
  map_sdf_saved(lo, P, saved, &sv);
  SVS_CHECK_ARG(!grad || lo.pe_w <= kStashLd, "tcgen05 engine: analytic gradient supports PE widths up to %d (got %d)", kStashLd, lo.pe_w);
  {
    TcChain ch;
    init_chain(&ch);
    sdf_geometry(&ch, d, x, P, clamp);
    add_sdf_forward_steps(&ch, lo, wi, wbuf, &sv, split);
    ch.y = y;
    ch.ldy = lo.ldy;
    TcStep s0 = make_step(reg + wi.fwd_sdf, wbuf + lo.boff[L - 1], wi.fwd_kb[L - 1], 16, 1, EP_Y);
    if (split) s0.w_lo = reg + wi.fwd_sdf_lo;
    ch.st[ch.n_steps++] = s0;
    TcStep s1 = make_step(reg + wi.fwd[L - 1], wbuf + lo.boff[L - 1] + 1, wi.fwd_kb[L - 1], wi.fwd_npad[L - 1], lo.out[L - 1] - 1, EP_Y);
    if (split) s1.w_lo = reg + wi.fwd_lo[L - 1];
    s1.y_col = 1;
    ch.st[ch.n_steps++] = s1;
    SVS_TRY(launch_chain(ch, "mlp_tc_sdf_fwd", chain_flops(ch), 0.0, st));
  }
  if (!grad) {
    // no analytic gradient wanted (e.g. the background SDF net): the backward then has no tangent sweep and never
    // reads U, so the reverse sweep is skipped; only the clamped sdf remains to be produced
    if (sdf) {
      sdf_clamp_kernel<<<blocks_for(P), 256, 0, st>>>(x, y, lo.ldy, P, d->d_in, d->sphere_radius, d->sphere_scale, clamp, sdf);
      SVS_LAUNCH_OK();
    }
    return SVS_OK;
  }
  {
    // reverse sweep: p_l = W_l^T (s_l * p_{l+1}), kept as U_l = s_l * p~_{l+1}
    TcChain ch;
    init_chain(&ch);
    sdf_geometry(&ch, d, x, P, clamp);
    ch.yin = y;
    ch.ldy = lo.ldy;
    ch.sdf = sdf;
    ch.grad = grad;
    ch.prologue = PRO_LOAD_ULAST;
    ch.pro_kb = kb_of(lo.in[L - 1]);
    ch.pro_vec = wbuf + lo.woff[L - 1];
    int ni = 0;
    ch.img[ni] = timg(sv.H[L - 1]);
    ch.pro_img = ni++;
    ch.img[ni] = timg(sv.U[L - 2]);
    ch.pro_save = ni++;
    for (int l = L - 2; l >= 1; --l) {
      TcStep s = make_step(reg + wi.bwd[l], nullptr, wi.bwd_kb[l], wi.bwd_npad[l], lo.in[l], EP_REVERSE);
      s.n_split = lo.out[l - 1];
      if (l == lo.skip) {
        s.scale = kInvSqrt2;
        s.hscale = kSqrt2;
      }
      ch.img[ni] = timg(sv.H[l]);
      s.aux1 = ni++;
      ch.img[ni] = timg(sv.U[l - 1]);
      s.save = ni++;
      s.next_kb = kb_of(lo.out[l - 1]);
      ch.st[ch.n_steps++] = s;
    }
    TcStep s = make_step(reg + wi.bwd[0], nullptr, wi.bwd_kb[0], wi.bwd_npad[0], lo.in[0], EP_PEGRAD);
    ch.st[ch.n_steps++] = s;
    SVS_TRY(launch_chain(ch, "mlp_tc_sdf_rev", chain_flops(ch), 0.0, st));
  }
  return SVS_OK;
}

static int sdf_outputs_backward(const svs_mlp_desc* d, const Layout& lo, const float* wbuf, const float* x, int64_t P,
                                int64_t clamp, const void* saved, const float* y, const float* dy, const float* d_sdf,
                                const float* d_grad, float* dwbuf, void* ws, cudaStream_t st) {
  WImages wi;
  SVS_TRY(make_wimages(d, lo, &wi, nullptr, nullptr));
  const uint8_t* reg = wimg_region(lo, wbuf);
  const int L = lo.L;
  SdfSaved sv;
  map_sdf_saved(lo, P, const_cast<void*>(saved), &sv);
  SdfBwdWs bw;
  map_sdf_bwd(lo, P, ws, &bw);
  const bool have_tangent = d_grad != nullptr;
  const bool have_top = dy != nullptr || d_sdf != nullptr;
  if (!have_tangent && !have_top) return SVS_OK;
  // gradients travel in fp16 scaled by a power of two derived from the largest upstream entry (J_PE amplifies the
  // tangent by up to 2^(n_freqs-1); sigma'' = 100 s (1-s) amplifies zeta), undone when dW / db are written
  uint32_t* amax = reinterpret_cast<uint32_t*>(ws);
  const float amax_target = 8.0f;
  SVS_TRY(launch_amax(dy, dy ? P * (int64_t)lo.ldy : 0, 1.f, d_sdf, d_sdf ? P : 0, 1.f, d_grad,
                      d_grad ? P * (int64_t)d->d_in : 0, (float)(1 << (d->n_freqs > 0 ? d->n_freqs - 1 : 0)), amax, st));

  if (have_tangent) {
    // tangent sweep (adjoint of the reverse sweep): q_{l+1} = s_l * (W_l q_l), zeta_l = sigma'' p~ r
    TcChain ch;
    init_chain(&ch);
    sdf_geometry(&ch, d, x, P, clamp);
    ch.yin = y;
    ch.ldy = lo.ldy;
    ch.d_grad = d_grad;
    ch.amax = amax;
    ch.amax_target = amax_target;
    ch.prologue = PRO_PE_JVP;
    ch.pro_kb = kb_of(lo.in[0]);
    int ni = 0;
    ch.img[ni] = timg(bw.Q[0]);
    ch.pro_save = ni++;
    for (int l = 0; l < L - 1; ++l) {
      TcStep s = make_step(reg + wi.fwd[l], nullptr, wi.fwd_kb[l], wi.fwd_npad[l], lo.out[l], EP_TANGENT);
      if (l + 1 == lo.skip) {
        s.scale = kInvSqrt2;
        s.hscale = kSqrt2;
        s.flags = TC_QFILL;
      }
      ch.img[ni] = timg(sv.H[l + 1]);
      s.aux1 = ni++;
      ch.img[ni] = timg(sv.U[l]);
      s.aux2 = ni++;
      ch.img[ni] = timg(bw.ZETA[l]);
      s.out2 = ni++;
      s.next_kb = kb_of(lo.in[l + 1]);
      if (l + 1 <= L - 2) {
        ch.img[ni] = timg(bw.Q[l + 1]);
        s.save = ni++;
      } else {
        // y_0 = W_{L-1}[0,:] a_{L-1} + b: the tangent reaches row 0 of the last weight directly
        s.colsum = 0;
        ch.colsum_out[0] = dwbuf + lo.woff[L - 1];
        ch.colsum_n[0] = lo.in[L - 1];
      }
      ch.st[ch.n_steps++] = s;
    }
    SVS_TRY(launch_chain(ch, "mlp_tc_sdf_tan", chain_flops(ch), 0.0, st));
  }
  {
    // ordinary backward: dz_{l-1} = s(h_l) * (dz_l W_l) + zeta_{l-1}
    TcChain ch;
    init_chain(&ch);
    sdf_geometry(&ch, d, x, P, clamp);
    ch.yin = y;
    ch.ldy = lo.ldy;
    ch.dy = dy;
    ch.d_sdf = d_sdf;
    ch.dy_cols = lo.out[L - 1];
    // fp32 rows staged through the aux ring by bulk copies (needs 16-byte aligned rows of the tile's flat array)
    static const bool no_bulk = getenv("SVS_DY_BULK") != nullptr && getenv("SVS_DY_BULK")[0] == '0';   // A/B runs, tests of the register path
    ch.dy_bulk = (!no_bulk && dy != nullptr && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (int64_t)kTile * lo.ldy < 65536 &&
                  lo.ldy <= 64 * kb_of(lo.out[L - 1])) ? 1 : 0;
    ch.dy_magic = (uint32_t)(0x100000000ull / (uint64_t)lo.ldy) + 1u;
    ch.amax = amax;
    ch.amax_target = amax_target;
    ch.prologue = PRO_DY;
    ch.pro_kb = kb_of(lo.out[L - 1]);
    int ni = 0;
    ch.img[ni] = timg(bw.DY);
    ch.pro_save = ni++;
    // bias gradients = column sums of dy / dz_l: taken by the weight-gradient kernel from the saved images (job_bias)
    for (int l = L - 1; l >= 1; --l) {
      TcStep s = make_step(reg + wi.bwd[l], nullptr, wi.bwd_kb[l], wi.bwd_npad[l], lo.out[l - 1], EP_BACKWARD);
      if (l == lo.skip) {
        s.scale = kInvSqrt2;
        s.hscale = kSqrt2;
      }
      ch.img[ni] = timg(sv.H[l]);
      s.aux1 = ni++;
      if (have_tangent) {
        ch.img[ni] = timg(bw.ZETA[l - 1]);
        s.aux2 = ni++;
      }
      ch.img[ni] = timg(bw.DZ[l - 1]);
      s.save = ni++;
      s.next_kb = kb_of(lo.out[l - 1]);
      ch.st[ch.n_steps++] = s;
    }
    SVS_TRY(launch_chain(ch, "mlp_tc_sdf_bwd", chain_flops(ch), 0.0, st));
  }
  {
    DwParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.amax = amax;
    prm.amax_target = amax_target;
    for (int l = 0; l < L - 1; ++l) {
      DwJob j = make_job(dwbuf + lo.woff[l], lo.ldi[l], lo.out[l], lo.in[l]);
      if (have_tangent) job_pair(&j, sv.U[l], bw.Q[l]);
      job_pair(&j, bw.DZ[l], l ? sv.H[l] : sv.A0);
      job_bias(&j, dwbuf + lo.boff[l]);
      j.x_blk0 = 0;
      j.n_mblk = (int)cdiv(lo.out[l], 128);
      j.y_blk0 = 0;
      j.n_yblk = j.y_kb;
      prm.job[prm.n_jobs++] = j;
    }
    {
      const int l = L - 1;
      const int n_main = lo.out[l] < 256 ? lo.out[l] : 256;
      DwJob j = make_job(dwbuf + lo.woff[l], lo.ldi[l], n_main, lo.in[l]);
      job_pair(&j, bw.DY, sv.H[l]);
      job_bias(&j, dwbuf + lo.boff[l]);
      j.n_mblk = (int)cdiv(n_main, 128);
      j.n_yblk = j.y_kb;
      prm.job[prm.n_jobs++] = j;
      if (lo.out[l] > 256) {
        DwJob k = make_job(dwbuf + lo.woff[l] + (size_t)256 * lo.ldi[l], lo.ldi[l], lo.out[l] - 256, lo.in[l]);
        job_pair(&k, bw.DY, sv.H[l]);
        job_bias(&k, dwbuf + lo.boff[l] + 256);
        k.x_blk0 = 4;
        k.n_mblk = 1;
        k.n_yblk = k.y_kb;
        prm.job[prm.n_jobs++] = k;
      }
    }
    SVS_TRY(launch_dw(prm, P, st));
  }
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// rendering network
// ---------------------------------------------------------------------------------------------------------------
static int render_forward(const svs_mlp_desc* d, const Layout& lo, const float* wbuf, const float* points,
                          const float* view_dirs, const float* normals, const float* feat, int ld_feat, int64_t P,
                          float* rgb, void* saved, cudaStream_t st, bool split) {
  WImages wi;
  SVS_TRY(make_wimages(d, lo, &wi, nullptr, nullptr, split));
  const uint8_t* reg = wimg_region(lo, wbuf);
  const int L = lo.L;
  RenderSaved sv;
  map_render_saved(lo, wi, P, saved, &sv);
  TcChain ch;
  init_chain(&ch);
  ch.P = P;
  ch.points = points; ch.view = view_dirs; ch.normals = normals; ch.feat = feat; ch.ld_feat = ld_feat;
  ch.view_freqs = d->n_freqs;
  ch.idr = d->render_mode == SVS_RENDER_IDR;
  ch.F = wi.F;
  ch.rgb = rgb;
  ch.split = split ? 1 : 0;
  ch.prologue = PRO_RENDER_IN;
  ch.pro_kb = wi.F / 64 + 1;
  int ni = 0;
  ch.img[ni] = timg(sv.RIN);
  ch.pro_save = ni++;
  for (int l = 0; l < L - 1; ++l) {
    TcStep s = make_step(reg + wi.fwd[l], wbuf + lo.boff[l], wi.fwd_kb[l], wi.fwd_npad[l], lo.out[l], EP_RELU);
    s.next_kb = kb_of(lo.in[l + 1]);
    if (split) s.w_lo = reg + wi.fwd_lo[l];
    ch.img[ni] = timg(sv.H[l + 1]);
    s.save = ni++;
    ch.st[ch.n_steps++] = s;
  }
  {
    TcStep s = make_step(reg + wi.fwd[L - 1], wbuf + lo.boff[L - 1], wi.fwd_kb[L - 1], wi.fwd_npad[L - 1], lo.out[L - 1], EP_RGB);
    if (split) s.w_lo = reg + wi.fwd_lo[L - 1];
    ch.st[ch.n_steps++] = s;
  }
  return launch_chain(ch, "mlp_tc_render_fwd", chain_flops(ch), 0.0, st);
}

static int render_backward(const svs_mlp_desc* d, const Layout& lo, const float* wbuf, int64_t P, const void* saved,
                           const float* rgb, const float* d_rgb, float* d_normals, float* d_feat, int ld_dfeat,
                           float* dwbuf, void* ws, cudaStream_t st) {
  WImages wi;
  SVS_TRY(make_wimages(d, lo, &wi, nullptr, nullptr));
  const uint8_t* reg = wimg_region(lo, wbuf);
  const int L = lo.L;
  const bool idr = d->render_mode == SVS_RENDER_IDR;
  RenderSaved sv;
  map_render_saved(lo, wi, P, const_cast<void*>(saved), &sv);
  RenderWs rw;
  map_render_ws(lo, P, ws, &rw);
  uint32_t* amax = reinterpret_cast<uint32_t*>(ws);
  const float amax_target = 64.0f;
  SVS_TRY(launch_amax(d_rgb, P * (int64_t)lo.out[L - 1], 1.f, nullptr, 0, 0.f, nullptr, 0, 0.f, amax, st));
  {
    TcChain ch;
    init_chain(&ch);
    ch.P = P;
    ch.amax = amax;
    ch.amax_target = amax_target;
    ch.rgb_in = rgb;
    ch.d_rgb = d_rgb;
    ch.n_rgb = lo.out[L - 1];
    ch.d_normals = idr ? d_normals : nullptr;
    ch.d_feat = d_feat;
    ch.ld_dfeat = ld_dfeat;
    ch.prologue = PRO_SIGMOID_BWD;
    ch.pro_kb = 1;
    int ni = 0;
    ch.img[ni] = timg(rw.DZ[L - 1]);
    ch.pro_save = ni++;
    // bias gradients: the last layer's (n_rgb sums of the prologue's fp32 values, heavily cancelling under an L1 loss) stay
    // here in fp32; the hidden layers' are column sums of the saved dz images, taken by the weight-gradient kernel
    ch.pro_colsum = 0;
    ch.colsum_out[0] = dwbuf + lo.boff[L - 1];
    ch.colsum_n[0] = lo.out[L - 1];
    for (int l = L - 1; l >= 1; --l) {
      TcStep s = make_step(reg + wi.bwd[l], nullptr, wi.bwd_kb[l], wi.bwd_npad[l], lo.in[l], EP_RELU_BWD);
      ch.img[ni] = timg(sv.H[l]);
      s.aux1 = ni++;
      ch.img[ni] = timg(rw.DZ[l - 1]);
      s.save = ni++;
      s.next_kb = kb_of(lo.in[l]);
      ch.st[ch.n_steps++] = s;
    }
    if (d_feat) ch.st[ch.n_steps++] = make_step(reg + wi.bwd[0], nullptr, wi.bwd_kb[0], wi.bwd_npad[0], wi.F, EP_DFEAT);
    if (idr && d_normals) {
      TcStep s = make_step(reg + wi.bwd_small, nullptr, wi.bwd_kb[0], wi.small_npad, wi.n_small, EP_DSMALL);
      s.y_col = 3 + 3 * (1 + 2 * d->n_freqs);  // [points, PE(view), normals]
      ch.st[ch.n_steps++] = s;
    }
    SVS_TRY(launch_chain(ch, "mlp_tc_render_bwd", chain_flops(ch), 0.0, st));
  }
  {
    DwParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.amax = amax;
    prm.amax_target = amax_target;
    {
      // layer 0: the saved input is [features | other columns]; W columns are [other | features]
      DwJob a = make_job(dwbuf + lo.woff[0], lo.ldi[0], lo.out[0], wi.F);
      job_pair(&a, rw.DZ[0], sv.RIN);
      job_bias(&a, dwbuf + lo.boff[0]);
      a.n_mblk = (int)cdiv(lo.out[0], 128);
      a.y_blk0 = 0;
      a.n_yblk = wi.F / 64;
      a.col_shift = wi.n_small;
      prm.job[prm.n_jobs++] = a;
      DwJob b = make_job(dwbuf + lo.woff[0], lo.ldi[0], lo.out[0], wi.n_small);
      job_pair(&b, rw.DZ[0], sv.RIN);
      b.n_mblk = (int)cdiv(lo.out[0], 128);
      b.y_blk0 = wi.F / 64;
      b.n_yblk = 1;
      prm.job[prm.n_jobs++] = b;
    }
    for (int l = 1; l < L; ++l) {
      DwJob j = make_job(dwbuf + lo.woff[l], lo.ldi[l], lo.out[l], lo.in[l]);
      job_pair(&j, rw.DZ[l], sv.H[l]);
      if (l < L - 1) job_bias(&j, dwbuf + lo.boff[l]);
      j.n_mblk = (int)cdiv(lo.out[l], 128);
      j.n_yblk = j.y_kb;
      prm.job[prm.n_jobs++] = j;
    }
    SVS_TRY(launch_dw(prm, P, st));
  }
  return SVS_OK;
}

}  // namespace tc
}  // namespace svs
