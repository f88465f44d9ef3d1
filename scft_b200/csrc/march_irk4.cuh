// march_irk4.cuh — contour march with the reference's production time stepper: the 2-stage
// Gauss-Legendre implicit Runge-Kutta scheme ("IRK4") of DEALII_SCFT (scft.cc:671-693 block matrix,
// drivescft.cc:130-146 step), on the same one-CTA-per-problem substructured solver as the
// implicit-Euler kernel (march1d.cuh).
//
// The reference solves, every step, the 2n x 2n block system
//     [A + dt/4 D      c01 D  ] [k1]   [-D q]          c01 = (1/4 - sqrt3/6) dt
//     [  c10 D      A + dt/4 D] [k2] = [-D q]          c10 = (1/4 + sqrt3/6) dt,   q+ = q + dt/2 (k1+k2)
// with UMFPACK.  For the linear problem M q' = -D q this is q+ = R(-dt M^-1 D) q with the (2,2) Pade
// approximant R(z) = (1 + z/2 + z^2/12)/(1 - z/2 + z^2/12), whose partial fractions give
//     q+ = q - 2 Re[ alpha (z1 A + dt D)^-1 A q ],   z1 = 3 + i sqrt3,  alpha = 6 - 6 i sqrt3:
// ONE complex symmetric tridiagonal solve per step instead of a real block solve of twice the
// size.  The result equals the block-LU march to rounding (5e-13 over 2048 steps, tests).
#pragma once
#include "march1d.cuh"

namespace scftb {

struct cx { double re, im; };
__device__ __forceinline__ cx mk(double re, double im = 0.0) { cx r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ cx operator+(cx a, cx b) { return mk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cx operator-(cx a, cx b) { return mk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cx operator-(cx a) { return mk(-a.re, -a.im); }
__device__ __forceinline__ cx operator*(cx a, cx b) { return mk(fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)); }
__device__ __forceinline__ cx operator*(cx a, double b) { return mk(a.re * b, a.im * b); }
__device__ __forceinline__ cx cinv(cx a) { double d = 1.0 / fma(a.re, a.re, a.im * a.im); return mk(a.re * d, -a.im * d); }
// c - a*b
__device__ __forceinline__ cx nfma(cx a, cx b, cx c) {
  return mk(fma(-a.re, b.re, fma(a.im, b.im, c.re)), fma(-a.re, b.im, fma(-a.im, b.re, c.im)));
}
// c + a*b
__device__ __forceinline__ cx pfma(cx a, cx b, cx c) {
  return mk(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im)));
}
// c + a*x with real x
__device__ __forceinline__ cx rfma(cx a, double x, cx c) { return mk(fma(a.re, x, c.re), fma(a.im, x, c.im)); }
__device__ __forceinline__ cx shfl_up_c(cx v, int d) { return mk(shfl_up_d(v.re, d), shfl_up_d(v.im, d)); }
__device__ __forceinline__ cx shfl_dn_c(cx v, int d) { return mk(shfl_dn_d(v.re, d), shfl_dn_d(v.im, d)); }

constexpr int PUBC = 18;  // doubles per warp: [0] qf, [2,3] zf, [4,5] Z0 (lane 0); [6,7] Z30 (lane 30); [8,9] rsep (lane 31); 144-byte rows keep the
                          // level-3 loads of up to eight rows in disjoint banks (96-byte rows conflicted 2-way)

template <int C, int T, bool UNI>
__global__ void __launch_bounds__(T) march_irk4_kernel(MarchParams P) {
  constexpr int CI = C - 1, CA = CI > 0 ? CI : 1, NW = T / 32, SL = T * C;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  __shared__ cx s_ex[2][T];
  __shared__ cx s_l3[NW][9];       // P, D, Nx, GL0, GR0, GL30, GR30, cAu(real), csu
  __shared__ cx s_minv[NW][NW];
  __shared__ __align__(16) double s_pub[2][NW][PUBC];
  __shared__ double s_red[NW];
  const int n = P.nsteps;
  const double dt = 1.0 / n;
  const cx z1 = mk(3.0, 1.7320508075688772);
  const double a_re = 12.0, a_im = 12.0 * 1.7320508075688772;   // 2*Re[alpha y] = 12 y_re + 12 sqrt3 y_im

  for (int p = blockIdx.x; p < P.nprob; p += gridDim.x) {
    if (P.skip && P.skip[p]) continue;
    const int pp = P.pshare ? 0 : p;      // parameter slot (MarchParams::pshare)
    const double L = P.L[pp];
    auto wrow = [&](const Row &r, cx &wl, cx &wd, cx &wu) {   // W = z1 A + dt D
      wl = mk(fma(dt, r.Dl, z1.re * r.Al), z1.im * r.Al);
      wd = mk(fma(dt, r.Dd, z1.re * r.Ad), z1.im * r.Ad);
      wu = mk(fma(dt, r.Du, z1.re * r.Au), z1.im * r.Au);
    };
    // ---------------------------------------------------------------- assembly + level 1 (LU order)
    cx ca[CA], cd[CA], cu[CA], al[CA], be[CA], gl[CA], gr[CA];
    double sAl, sAd, sAu;
    cx sl, sd, su;
    {
      Row rs = assemble_row(P, p, t * C + CI, L, dt);
      sAl = rs.Al; sAd = rs.Ad; sAu = rs.Au;
      if (UNI) sAd = (rs.Al != 0.0) ? rs.Al : rs.Au;   // UNI: A's row is A_off (1,4,1); one coefficient is kept (march1d.cuh)
      if (t * C + CI >= P.ni) { sl = mk(0.0); sd = mk(1.0); su = mk(0.0); }   // padding: identity row
      else wrow(rs, sl, sd, su);
    }
    if constexpr (CI > 0) {
      cx Tl0 = mk(0.0), TuL = mk(0.0), pinv_prev = mk(0.0), Wu_prev = mk(0.0);
      cx alo[CA];   // pinv_k * Wl_k (setup only)
#pragma unroll
      for (int k = 0; k < CI; k++) {
        Row r = assemble_row(P, p, t * C + k, L, dt);
        cx wl, wd, wu;
        if (t * C + k >= P.ni) { wl = mk(0.0); wd = mk(1.0); wu = mk(0.0); }
        else wrow(r, wl, wd, wu);
        cx piv = (k == 0) ? wd : wd - (wl * pinv_prev) * Wu_prev;
        cx pinv = cinv(piv);
        ca[k] = pinv * ((UNI && r.Al == 0.0) ? r.Au : r.Al); cd[k] = pinv * r.Ad; cu[k] = pinv * r.Au;
        alo[k] = (k == 0) ? mk(0.0) : pinv * wl;
        // UNI: the forward sweep runs on u = y / A_off (real right-hand side) with the plain multiplier Wl_k / piv_{k-1}
        al[k] = (k == 0) ? mk(0.0) : (UNI ? wl * pinv_prev : alo[k]);
        be[k] = (k == CI - 1) ? mk(0.0) : pinv * wu;
        if (k == 0) Tl0 = pinv * wl;
        if (k == CI - 1) TuL = pinv * wu;
        pinv_prev = pinv; Wu_prev = wu;
      }
      cx y[CA];
      y[0] = Tl0;
#pragma unroll
      for (int k = 1; k < CI; k++) y[k] = -(alo[k] * y[k - 1]);
      gl[CI - 1] = y[CI - 1];
#pragma unroll
      for (int k = CI - 2; k >= 0; k--) gl[k] = nfma(be[k], gl[k + 1], y[k]);
      gr[CI - 1] = TuL;
#pragma unroll
      for (int k = CI - 2; k >= 0; k--) gr[k] = -(be[k] * gr[k + 1]);
    }
    // ---------------------------------------------------------------- Schur rows on the separators
    cx a, b, c;
    __syncthreads();
    if constexpr (CI > 0) {
      s_ex[0][t] = gl[0]; s_ex[1][t] = gr[0];
      __syncthreads();
      cx gl0n = (t + 1 < T) ? s_ex[0][t + 1] : mk(0.0), gr0n = (t + 1 < T) ? s_ex[1][t + 1] : mk(0.0);
      a = -(sl * gl[CI - 1]);
      b = nfma(su, gl0n, nfma(sl, gr[CI - 1], sd));
      c = -(su * gr0n);
    } else { a = sl; b = sd; c = su; }
    // ---------------------------------------------------------------- level 2: cyclic reduction
    const cx l3P = a, l3D = b, l3N = c;
    const cx A0 = (lane == 0) ? a : mk(0.0), C30 = (lane == 30) ? c : mk(0.0);
    if (lane == 31) { a = mk(0.0); b = mk(1.0); c = mk(0.0); }
    if (lane == 0) a = mk(0.0);
    if (lane == 30) c = mk(0.0);
    cx pa_[5], pg_[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int d = 1 << s;
      cx am = shfl_up_c(a, d), bm = shfl_up_c(b, d), cm = shfl_up_c(c, d);
      cx ap = shfl_dn_c(a, d), bp = shfl_dn_c(b, d), cp = shfl_dn_c(c, d);
      cx alpha = (lane >= d) ? -(a * cinv(bm)) : mk(0.0);
      cx gamma = (lane + d <= 31) ? -(c * cinv(bp)) : mk(0.0);
      if (lane < d) { am = mk(0.0); cm = mk(0.0); }
      if (lane + d > 31) { ap = mk(0.0); cp = mk(0.0); }
      b = pfma(gamma, ap, pfma(alpha, cm, b));
      a = alpha * am;
      c = gamma * cp;
      pa_[s] = alpha; pg_[s] = gamma;
    }
    const cx binv = cinv(b);
    auto pcr = [&](cx r) {
#pragma unroll
      for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        cx rm = shfl_up_c(r, d), rp = shfl_dn_c(r, d);
        r = pfma(pa_[s], rm, pfma(pg_[s], rp, r));
      }
      return r * binv;
    };
    const cx GL = pcr(A0), GR = pcr(C30);
    // ---------------------------------------------------------------- level 3 setup
    if (lane == 31) { s_l3[wid][0] = l3P; s_l3[wid][1] = l3D; s_l3[wid][2] = l3N;
                      s_l3[wid][7] = mk((wid + 1 < NW) ? (UNI ? sAd : sAu) : 0.0); s_l3[wid][8] = (CI > 0) ? su : mk(0.0); }
    if (lane == 0) { s_l3[wid][3] = GL; s_l3[wid][4] = GR; }
    if (lane == 30) { s_l3[wid][5] = GL; s_l3[wid][6] = GR; }
    __syncthreads();
    if (t < NW) {
      cx cc[NW], dd[NW];
      cx cprev = mk(0.0), dprev = mk(0.0);
#pragma unroll
      for (int w = 0; w < NW; w++) {
        cx Pw = s_l3[w][0], Dw = s_l3[w][1], Nw = s_l3[w][2];
        cx lo = (w > 0) ? -(Pw * s_l3[w][5]) : mk(0.0);
        cx di = nfma(Pw, s_l3[w][6], Dw);
        if (w + 1 < NW) di = nfma(Nw, s_l3[(w + 1) % NW][3], di);
        cx up = (w + 1 < NW) ? -(Nw * s_l3[(w + 1) % NW][4]) : mk(0.0);
        cx rhs = (w == t) ? mk(1.0) : mk(0.0);
        cx deninv = cinv(nfma(lo, cprev, di));
        cc[w] = up * deninv;
        dd[w] = nfma(lo, dprev, rhs) * deninv;
        cprev = cc[w]; dprev = dd[w];
      }
      cx xn = mk(0.0);
#pragma unroll
      for (int w = NW - 1; w >= 0; w--) { xn = nfma(cc[w], xn, dd[w]); s_minv[w][t] = xn; }
    }
    __syncthreads();

    // the lane's share of the level-3 solve (march1d.cuh): lane v3 forms R_v3, butterfly sum over NW lanes
    const int v3 = lane & (NW - 1);
    const cx c3P = s_l3[v3][0], c3su = s_l3[v3][8], c3Nx = s_l3[v3][2];
    const double c3Au = s_l3[v3][7].re;
    const cx mW = s_minv[wid][v3], mM = (wid > 0) ? s_minv[(wid + NW - 1) % NW][v3] : mk(0.0);
    // ---------------------------------------------------------------- initial condition
    double q[C], phi[C];
#pragma unroll
    for (int k = 0; k < C; k++) { q[k] = (t * C + k < P.ni) ? 1.0 : 0.0; phi[k] = 0.0; }
    double XL = (t > 0 && t * C - 1 < P.ni) ? 1.0 : 0.0;
    double qn = ((t + 1) * C < P.ni) ? 1.0 : 0.0;
    double *hb = P.hist + (size_t)(P.store_full ? p : blockIdx.x) * P.hist_stride;
    auto hidx = [&](int k) { return hist_index(C, T, t, k); };
#pragma unroll
    for (int k = 0; k < C; k++) hb[hidx(k)] = q[k];
    const bool full = P.store_full != 0;

    // ---------------------------------------------------------------- the contour march
    for (int j = 1; j <= n; j++) {
      const bool pairing = (2 * j > n);
      double qo[C];
      if (pairing) {
        const double *hs = hb + (size_t)(n - j) * SL;
#pragma unroll
        for (int k = 0; k < C; k++) qo[k] = hs[hidx(k)];
      }
      // rhs A q (real); chunk solve with zero separators
      cx z[CA];
      cx zlast = mk(0.0), z0 = mk(0.0);
      if constexpr (CI > 0) {
        if constexpr (UNI) {
#pragma unroll
          for (int k = 0; k < CI; k++) {
            double qm = (k == 0) ? XL : q[k - 1], qp = q[k + 1];
            const double tk = fma(4.0, q[k], qm + qp);
            z[k] = (k == 0) ? mk(tk) : nfma(al[k], z[k - 1], mk(tk));
          }
          z[CI - 1] = ca[CI - 1] * z[CI - 1];
#pragma unroll
          for (int k = CI - 2; k >= 0; k--) z[k] = nfma(be[k], z[k + 1], ca[k] * z[k]);
        } else {
#pragma unroll
          for (int k = 0; k < CI; k++) {
            double qm = (k == 0) ? XL : q[k - 1], qp = q[k + 1];
            cx bk = rfma(ca[k], qm, rfma(cu[k], qp, cd[k] * q[k]));
            z[k] = (k == 0) ? bk : nfma(al[k], z[k - 1], bk);
          }
#pragma unroll
          for (int k = CI - 2; k >= 0; k--) z[k] = nfma(be[k], z[k + 1], z[k]);
        }
        zlast = z[CI - 1]; z0 = z[0];
      }
      const double qprev = (CI > 0) ? q[CI > 0 ? CI - 1 : 0] : XL;
      cx r = mk(UNI ? sAd * fma(4.0, q[C - 1], qprev) : fma(sAl, qprev, sAd * q[C - 1]));
      if constexpr (CI > 0) r = nfma(sl, zlast, r);
      const cx rsep = r;
      {
        r.re = fma(UNI ? sAd : sAu, qn, r.re);
        if constexpr (CI > 0) { cx zfn = shfl_dn_c(z0, 1); r = nfma(su, zfn, r); }
      }
      if (lane == 31) r = mk(0.0);
#pragma unroll
      for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        cx rm = shfl_up_c(r, d), rp = shfl_dn_c(r, d);
        r = pfma(pa_[s], rm, pfma(pg_[s], rp, r));
      }
      const cx Z = r * binv;
      double *pb = s_pub[j & 1][wid];
      if (lane == 0) { pb[0] = q[0]; pb[2] = z0.re; pb[3] = z0.im; pb[4] = Z.re; pb[5] = Z.im; }
      if (lane == 30) { pb[6] = Z.re; pb[7] = Z.im; }
      if (lane == 31) { pb[8] = rsep.re; pb[9] = rsep.im; }
      if constexpr (NW > 1) __syncthreads(); else __syncwarp();
      cx Wm, Ww;
      {
        const double *pv = s_pub[j & 1][v3], *pn = s_pub[j & 1][(v3 + 1) % NW];
        cx R = nfma(c3P, mk(pv[6], pv[7]), mk(pv[8], pv[9]));
        cx R2 = nfma(c3su, mk(pn[2], pn[3]), mk(c3Au * pn[0]));
        R2 = nfma(c3Nx, mk(pn[4], pn[5]), R2);
        R = R + R2;
        Ww = mW * R; Wm = mM * R;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
          Ww.re += __shfl_xor_sync(0xffffffffu, Ww.re, d); Ww.im += __shfl_xor_sync(0xffffffffu, Ww.im, d);
          Wm.re += __shfl_xor_sync(0xffffffffu, Wm.re, d); Wm.im += __shfl_xor_sync(0xffffffffu, Wm.im, d);
        }
      }
      const cx Y = (lane == 31) ? Ww : nfma(GL, Wm, nfma(GR, Ww, Z));   // solution at the own separator
      cx YL = shfl_up_c(Y, 1);
      if (lane == 0) YL = Wm;                                            // previous warp's separator
      // q+ = q - 2 Re[alpha y]
      if constexpr (CI > 0) {
#pragma unroll
        for (int k = 0; k < CI; k++) {
          cx yk = nfma(gl[k], YL, nfma(gr[k], Y, z[k]));
          q[k] = fma(-a_re, yk.re, fma(-a_im, yk.im, q[k]));
        }
      }
      q[C - 1] = fma(-a_re, Y.re, fma(-a_im, Y.im, q[C - 1]));
      XL = fma(-a_re, YL.re, fma(-a_im, YL.im, XL));
      qn = shfl_dn_d(q[0], 1);
      if (lane == 31) qn = 0.0;
      if (full || 2 * j < n) {
        double *hs = hb + (size_t)j * SL;
#pragma unroll
        for (int k = 0; k < C; k++) hs[hidx(k)] = q[k];
      }
      if (2 * j >= n) {
        const double wj = __ldg(P.w + j);
#pragma unroll
        for (int k = 0; k < C; k++) phi[k] = fma(wj * q[k], pairing ? qo[k] : q[k], phi[k]);
      }
    }

    // ---------------------------------------------------------------- residual, phi, Q
    double qsum = 0.0;
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int g = t * C + k;
      if (g < P.ni) {
        const int i = g + 1;
        const double f0 = P.f0[(size_t)pp * P.N + i];
        P.out[(size_t)p * P.out_stride + g] = P.sign * (f0 - phi[k]);
        if (!P.pshare) P.phi[(size_t)p * P.N + i] = phi[k];
        double hw2;
        if (P.uniform) { double h = L / (P.N - 1); hw2 = 0.5 * (h + h); }
        else { const double *x = P.x + (size_t)pp * P.N; hw2 = 0.5 * ((x[i] - x[i - 1]) + (x[i + 1] - x[i])); }
        qsum += hw2 * q[k];
        if (P.eta_full && !P.pshare) P.eta_full[(size_t)p * P.N + i] = P.eta_mid[(size_t)p * P.eta_stride + g];
      }
    }
    if (t == 0 && !P.pshare) {
      P.phi[(size_t)p * P.N] = 0.0; P.phi[(size_t)p * P.N + P.N - 1] = 0.0;
      if (P.eta_full) {
        P.eta_full[(size_t)p * P.N] = eta_node(P, p, 0, L);
        P.eta_full[(size_t)p * P.N + P.N - 1] = eta_node(P, p, P.N - 1, L);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, d);
    if (lane == 0) s_red[wid] = qsum;
    __syncthreads();
    if (t == 0 && !P.pshare) {
      double s = 0.0;
      for (int w = 0; w < NW; w++) s += s_red[w];
      double len = P.uniform ? L : (P.x[(size_t)pp * P.N + P.N - 1] - P.x[(size_t)pp * P.N]);
      P.Q[p] = s / len;
    }
  }
}

}  // namespace scftb
