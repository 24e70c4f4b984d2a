// Fused Gram-matrix kernels.
//
// Forward: K[i,j] = program(prim_0(x_i,x'_j), ..., prim_{P-1}(x_i,x'_j)) -- the scaled squared
// distance / dot products, the exp / Matern / periodic bodies and the Sum / Product / Neural-
// Kernel-Network composition are evaluated per element in registers; neither the distance
// matrix nor any per-primitive Gram is ever written to memory (the reference materialises
// k + #layers full N x M temporaries, Periodic even an N x M x D one: kernels.py:806-819,
// neural_kernel_network.py:41-47).
//
// Every primitive is reduced to inner products of per-point FEATURES computed once per call
// (O(N D) work):
//   stationary (kernels.py:408-421): f = [x/l (nd), |x/l|^2]; d2 = s_i + s_j - 2 <f_i,f_j>,
//                                    clipped at 0 exactly like the reference;
//   linear     (kernels.py:499-505): f = x;  k = sum_d (x_d v_d) x'_d;
//   periodic   (kernels.py:806-819): f = [cos a, sin a, a], a = 2 pi x / p, using
//        sum_d sin^2(pi (x_d-x'_d)/p) = (nd - sum_d cos(a_d - a'_d)) / 2   (exact identity),
//        so the N x M x D sin() evaluations of the reference become N x D.
//
// Backward: dtheta[t] = sum_ij W_ij dK_ij/dtheta_t (and optionally the gradient w.r.t. the rows
// of X), recomputing K tile by tile; W is either a dense matrix or, for the fused GPR
// objective, formed on the fly as 1/2 (R K^-1 - beta beta^T) from the lower triangle of K^-1.
// Reductions are deterministic (per-CTA partials + a fixed-order second pass).
#include <algorithm>
#include <math.h>
#include <string.h>

#include "internal.cuh"

namespace {

constexpr int TILE = 64;
constexpr int GRAM_THREADS = 256;
constexpr int MAX_ACC = 192;     // per-thread gradient accumulators (n_theta + 1)
constexpr int MAX_FEAT = 160;    // features per point
constexpr int MAX_R = 16;        // output columns in the fused GPR weight

struct PrimC {
  int16_t type, ard, ndims, theta_off, feat_off, feat_cnt;
};
struct OpC {
  int16_t op, dst, a, b, c, d, n, pad;
};
struct Plan {
  int n_prims, n_ops, n_theta, out_slot, FT, S;   // S = padded (odd) feature row stride
  PrimC prims[GPS_MAX_PRIMS];
  OpC ops[GPS_MAX_OPS];
};
struct PlanDims {
  uint8_t dims[GPS_MAX_PRIMS][GPS_MAX_DIMS];
};

__host__ __device__ inline bool is_stationary(int t) { return t <= GPS_MATERN52; }

int build_plan(gps_handle* h, const gps_kernel_desc* d, int64_t xcols, Plan* pl, PlanDims* pd) {
  if (!d) return gps_fail(h, -2, "null kernel descriptor");
  if (d->n_prims < 1 || d->n_prims > GPS_MAX_PRIMS) return gps_fail(h, -2, "desc: n_prims out of range");
  if (d->n_ops < 0 || d->n_ops > GPS_MAX_OPS) return gps_fail(h, -2, "desc: n_ops out of range");
  if (d->n_theta < 1 || d->n_theta + 1 > MAX_ACC) return gps_fail(h, -2, "desc: n_theta out of range (max %d)", MAX_ACC - 1);
  memset(pl, 0, sizeof(*pl));
  memset(pd, 0, sizeof(*pd));
  pl->n_prims = d->n_prims; pl->n_ops = d->n_ops; pl->n_theta = d->n_theta; pl->out_slot = d->out_slot;
  int ft = 0;
  for (int p = 0; p < d->n_prims; ++p) {
    const gps_prim& q = d->prims[p];
    if (q.type < GPS_RBF || q.type > GPS_PERIODIC) return gps_fail(h, -2, "desc: primitive %d has unknown type %d", p, q.type);
    if (q.ndims < 1 || q.ndims > GPS_MAX_DIMS) return gps_fail(h, -2, "desc: primitive %d ndims out of range", p);
    int np = is_stationary(q.type) ? 1 + (q.ard ? q.ndims : 1)
             : q.type == GPS_LINEAR ? (q.ard ? q.ndims : 1) : 3;
    if (q.theta_off < 0 || q.theta_off + np > d->n_theta) return gps_fail(h, -2, "desc: primitive %d theta range", p);
    for (int k = 0; k < q.ndims; ++k) {
      if (q.dims[k] < 0 || q.dims[k] >= xcols || q.dims[k] > 255)
        return gps_fail(h, -2, "desc: primitive %d uses column %d but X has %lld columns", p, q.dims[k], (long long)xcols);
      pd->dims[p][k] = (uint8_t)q.dims[k];
    }
    PrimC& c = pl->prims[p];
    c.type = (int16_t)q.type; c.ard = (int16_t)(q.ard ? 1 : 0); c.ndims = (int16_t)q.ndims;
    c.theta_off = (int16_t)q.theta_off; c.feat_off = (int16_t)ft;
    c.feat_cnt = (int16_t)(is_stationary(q.type) ? q.ndims + 1 : q.type == GPS_LINEAR ? q.ndims : 3 * q.ndims);
    ft += c.feat_cnt;
  }
  if (ft > MAX_FEAT) return gps_fail(h, -2, "desc: %d features per point exceed the limit %d", ft, MAX_FEAT);
  pl->FT = ft;
  pl->S = ft | 1;
  int nslots = d->n_prims;
  for (int i = 0; i < d->n_ops; ++i) {
    const gps_op& o = d->ops[i];
    OpC& c = pl->ops[i];
    c.op = (int16_t)o.op; c.dst = (int16_t)o.dst; c.a = (int16_t)o.a; c.b = (int16_t)o.b;
    c.c = (int16_t)o.c; c.d = (int16_t)o.d; c.n = (int16_t)o.n; c.pad = 0;
    int nout = 1;
    bool ok = o.dst >= 0;
    switch (o.op) {
      case GPS_OP_CONST: ok = ok && o.a >= 0 && o.a < d->n_theta; break;
      case GPS_OP_ADD: case GPS_OP_MUL: ok = ok && o.a >= 0 && o.a < nslots && o.b >= 0 && o.b < nslots; break;
      case GPS_OP_COPY: ok = ok && o.a >= 0 && o.a < nslots; break;
      case GPS_OP_LINEAR:
        nout = o.n;
        ok = ok && o.b >= 1 && o.n >= 1 && o.a >= 0 && o.a + o.b <= nslots && o.c >= 0 &&
             o.c + o.n * o.b <= d->n_theta && o.d >= 0 && o.d + o.n <= d->n_theta;
        break;
      case GPS_OP_PRODUCT:
        nout = o.n;
        ok = ok && o.b >= 1 && o.n >= 1 && o.a >= 0 && o.a + o.b * o.n <= nslots;
        break;
      default: ok = false;
    }
    // ops append slots in order: dst must be the next free slot
    if (!ok || o.dst != nslots || o.dst + nout > GPS_MAX_SLOTS)
      return gps_fail(h, -2, "desc: op %d is malformed (op=%d dst=%d, next free slot %d)", i, o.op, o.dst, nslots);
    nslots += nout;
  }
  if (d->out_slot < 0 || d->out_slot >= nslots) return gps_fail(h, -2, "desc: out_slot out of range");
  return 0;
}

// --------------------------------------------------------------------------- features
__global__ void feature_kernel(const Plan pl, const PlanDims pd, const double* __restrict__ theta,
                               const double* __restrict__ X, int64_t ldx, int64_t N,
                               double* __restrict__ F) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* x = X + i * ldx;
  double* f = F + i * pl.FT;
  for (int p = 0; p < pl.n_prims; ++p) {
    const PrimC P = pl.prims[p];
    double* fp = f + P.feat_off;
    const double* th = theta + P.theta_off;
    if (is_stationary(P.type)) {
      double s = 0.0;
      for (int k = 0; k < P.ndims; ++k) {
        double v = x[pd.dims[p][k]] / th[1 + (P.ard ? k : 0)];   // X / lengthscales (kernels.py:409)
        fp[k] = v;
        s += v * v;                                              // reduce_sum(square(X)) (:410)
      }
      fp[P.ndims] = s;
    } else if (P.type == GPS_LINEAR) {
      for (int k = 0; k < P.ndims; ++k) fp[k] = x[pd.dims[p][k]];
    } else {
      const double per = th[2];
      for (int k = 0; k < P.ndims; ++k) {
        double a = 2.0 * M_PI * x[pd.dims[p][k]] / per;
        double sn, cs;
        sincos(a, &sn, &cs);
        fp[k] = cs;
        fp[P.ndims + k] = sn;
        fp[2 * P.ndims + k] = a;
      }
    }
  }
}

// --------------------------------------------------------------------------- primitive bodies
struct PrimEval {
  double k;     // value
  double dk;    // stationary: dk/d(d2) (0 where clipped); periodic: r = sum sin^2/l^2
  double d2;    // stationary: unclipped-side squared distance (after clip)
};

// value and d/d(d2) of a stationary kernel at squared distance d2 (already clipped at 0;
// `live` is false where the clip was active and the reference's gradient is zero)
__device__ __forceinline__ void stat_body(int type, double var, double d2, bool live, double& k,
                                          double& dk) {
  if (type == GPS_RBF) {
    double ex = exp(-0.5 * d2);
    k = var * ex;
    dk = live ? -0.5 * var * ex : 0.0;
  } else {
    double r = sqrt(d2 + 1e-12);                               // kernels.py:424-426
    double drd2 = live ? 0.5 / r : 0.0;
    if (type == GPS_EXPONENTIAL) {
      double ex = exp(-0.5 * r);
      k = var * ex;
      dk = -0.5 * var * ex * drd2;
    } else if (type == GPS_MATERN12) {
      double ex = exp(-r);
      k = var * ex;
      dk = -var * ex * drd2;
    } else if (type == GPS_MATERN32) {
      const double s3 = 1.7320508075688772;
      double ex = exp(-s3 * r);
      k = var * (1.0 + s3 * r) * ex;
      dk = var * (-3.0 * r) * ex * drd2;                       // d/dr[(1+s3 r)e^{-s3 r}] = -3 r e^{-s3 r}
    } else {
      const double s5 = 2.23606797749979;
      double ex = exp(-s5 * r);
      k = var * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * ex;
      dk = var * (-(5.0 / 3.0) * r * (1.0 + s5 * r)) * ex * drd2;
    }
  }
}

__device__ __forceinline__ PrimEval prim_eval(const PrimC& P, const double* __restrict__ th,
                                              const double* __restrict__ fi,
                                              const double* __restrict__ fj) {
  PrimEval e;
  e.k = 0; e.dk = 0; e.d2 = 0;
  const double* t = th + P.theta_off;
  if (is_stationary(P.type)) {
    double dot = 0.0;
    for (int k = 0; k < P.ndims; ++k) dot = fma(fi[k], fj[k], dot);
    double raw = -2.0 * dot + (fi[P.ndims] + fj[P.ndims]);      // kernels.py:413-414 / 419-420
    bool live = raw >= 0.0;                                      // clip_by_value(dist, 0, inf)
    double d2 = live ? raw : 0.0;
    e.d2 = d2;
    stat_body(P.type, t[0], d2, live, e.k, e.dk);
  } else if (P.type == GPS_LINEAR) {
    double dot = 0.0;
    for (int k = 0; k < P.ndims; ++k) dot = fma(fi[k] * t[P.ard ? k : 0], fj[k], dot);
    e.k = dot;
  } else {
    const int nd = P.ndims;
    double cs = 0.0;
    for (int k = 0; k < nd; ++k) cs += fi[k] * fj[k] + fi[nd + k] * fj[nd + k];
    double ls = t[1];
    double r = 0.5 * ((double)nd - cs) / (ls * ls);
    e.k = t[0] * exp(-0.5 * r);
    e.dk = r;
  }
  return e;
}

__device__ __forceinline__ void program_fwd(const Plan& pl, const double* __restrict__ th, double* v) {
  for (int q = 0; q < pl.n_ops; ++q) {
    const OpC o = pl.ops[q];
    switch (o.op) {
      case GPS_OP_CONST: v[o.dst] = th[o.a]; break;
      case GPS_OP_ADD: v[o.dst] = v[o.a] + v[o.b]; break;
      case GPS_OP_MUL: v[o.dst] = v[o.a] * v[o.b]; break;
      case GPS_OP_COPY: v[o.dst] = v[o.a]; break;
      case GPS_OP_LINEAR:
        for (int r = 0; r < o.n; ++r) {
          double s = 0.0;
          for (int c = 0; c < o.b; ++c) s = fma(v[o.a + c], th[o.c + r * o.b + c], s);
          v[o.dst + r] = s + th[o.d + r];
        }
        break;
      case GPS_OP_PRODUCT:
        for (int g = 0; g < o.n; ++g) {
          double s = v[o.a + g * o.b];
          for (int c = 1; c < o.b; ++c) s *= v[o.a + g * o.b + c];
          v[o.dst + g] = s;
        }
        break;
    }
  }
}

// reverse sweep; vb = adjoints of the slots, acc = gradient accumulators over theta
__device__ __forceinline__ void program_bwd(const Plan& pl, const double* __restrict__ th,
                                            const double* v, double* vb, double* acc) {
  for (int q = pl.n_ops - 1; q >= 0; --q) {
    const OpC o = pl.ops[q];
    switch (o.op) {
      case GPS_OP_CONST: acc[o.a] += vb[o.dst]; break;
      case GPS_OP_ADD: vb[o.a] += vb[o.dst]; vb[o.b] += vb[o.dst]; break;
      case GPS_OP_MUL: {
        double g = vb[o.dst];
        double va = v[o.a], vbv = v[o.b];
        vb[o.a] += g * vbv;
        vb[o.b] += g * va;
      } break;
      case GPS_OP_COPY: vb[o.a] += vb[o.dst]; break;
      case GPS_OP_LINEAR:
        for (int r = 0; r < o.n; ++r) {
          double g = vb[o.dst + r];
          acc[o.d + r] += g;
          for (int c = 0; c < o.b; ++c) {
            acc[o.c + r * o.b + c] = fma(g, v[o.a + c], acc[o.c + r * o.b + c]);
            vb[o.a + c] = fma(g, th[o.c + r * o.b + c], vb[o.a + c]);
          }
        }
        break;
      case GPS_OP_PRODUCT:
        for (int g = 0; g < o.n; ++g) {
          double gg = vb[o.dst + g];
          for (int c = 0; c < o.b; ++c) {
            double s = gg;
            for (int c2 = 0; c2 < o.b; ++c2)
              if (c2 != c) s *= v[o.a + g * o.b + c2];
            vb[o.a + g * o.b + c] += s;
          }
        }
        break;
    }
  }
}

// --------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(GRAM_THREADS)
gram_fwd_kernel(const Plan pl, const double* __restrict__ theta, const double* __restrict__ FL,
                const double* __restrict__ FR, int64_t N, int64_t M, double diag_add, int sym,
                int uplo, double* __restrict__ K, int64_t ldk) {
  extern __shared__ double sm[];
  double* th = sm;
  double* sl = th + pl.n_theta;
  double* sr = sl + TILE * pl.S;
  const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
  if (sym && uplo && j0 > i0 + TILE - 1) return;
  const int tid = threadIdx.x;
  for (int t = tid; t < pl.n_theta; t += GRAM_THREADS) th[t] = theta[t];
  for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
    sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double v[GPS_MAX_SLOTS];
  for (int a = 0; a < 4; ++a) {
    const int il = ty + 16 * a;
    const int64_t gi = i0 + il;
    if (gi >= N) continue;
    for (int b = 0; b < 4; ++b) {
      const int jl = tx + 16 * b;
      const int64_t gj = j0 + jl;
      if (gj >= M) continue;
      if (sym && uplo && gj > gi) continue;
      for (int p = 0; p < pl.n_prims; ++p) {
        const PrimC P = pl.prims[p];
        v[p] = prim_eval(P, th, sl + il * pl.S + P.feat_off, sr + jl * pl.S + P.feat_off).k;
      }
      program_fwd(pl, th, v);
      double out = v[pl.out_slot];
      if (sym && gi == gj) out += diag_add;
      K[gi * ldk + gj] = out;
    }
  }
}

// --------------------------------------------------------------------------- backward
struct BwdArgs {
  int mode;               // W_DENSE / W_GPR
  const double* W;        // dense weights or K^-1 (lower)
  int64_t ldw;
  const double* beta;     // [R][N]
  int R;
  int sym_lower;          // iterate lower tiles only, off-diagonal elements weigh double
  int want_dx;
  int xcols;
  double dx_scale;        // 2 for the symmetric problem (row + column role of X)
  int njc;                // column chunks
};

__global__ void __launch_bounds__(GRAM_THREADS)
gram_bwd_kernel(const Plan pl, const PlanDims pd, const double* __restrict__ theta,
                const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                const BwdArgs w, double* __restrict__ part_theta, double* __restrict__ part_dx) {
  extern __shared__ double sm[];
  const int nacc = pl.n_theta + 1;
  double* th = sm;                                 // [n_theta]
  double* tinv = th + pl.n_theta;                  // [n_theta]  1 / theta: the per-element gradient terms
                                                   // multiply by these instead of dividing (an FP64 division
                                                   // is ~25 instructions; the NKN config had ~40 per element)
  double* sl = tinv + pl.n_theta;                  // [TILE][S]
  double* sr = sl + TILE * pl.S;                   // [TILE][S]
  double* bi = sr + TILE * pl.S;                   // [MAX_R][TILE]   beta rows (i side)
  double* bj = bi + (w.mode == W_GPR ? w.R * TILE : 0);
  double* red = bj + (w.mode == W_GPR ? w.R * TILE : 0);   // [8][nacc]
  double* dxs = red + 8 * nacc;                    // [TILE][xcols]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * TILE;
  const int64_t jtiles = (M + TILE - 1) / TILE;

  double acc[MAX_ACC];
  for (int t = 0; t < nacc; ++t) acc[t] = 0.0;
  for (int t = tid; t < pl.n_theta; t += GRAM_THREADS) {
    th[t] = theta[t];
    tinv[t] = 1.0 / theta[t];
  }
  for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
  }
  if (w.mode == W_GPR)
    for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
      int r = idx / TILE, c = idx - r * TILE;
      bi[idx] = (i0 + c < N) ? w.beta[(int64_t)r * N + i0 + c] : 0.0;
    }
  if (w.want_dx)
    for (int idx = tid; idx < TILE * w.xcols; idx += GRAM_THREADS) dxs[idx] = 0.0;

  double v[GPS_MAX_SLOTS], vb[GPS_MAX_SLOTS];
  double dxr[GPS_MAX_DIMS];

  for (int64_t jt = blockIdx.x; jt < jtiles; jt += w.njc) {
    const int64_t j0 = jt * TILE;
    if (w.sym_lower && j0 > i0 + TILE - 1) break;
    __syncthreads();   // previous tile fully consumed
    for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
      int r = idx / pl.FT, c = idx - r * pl.FT;
      sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
    }
    if (w.mode == W_GPR)
      for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
        int r = idx / TILE, c = idx - r * TILE;
        bj[idx] = (j0 + c < M) ? w.beta[(int64_t)r * N + j0 + c] : 0.0;
      }
    __syncthreads();
    for (int a = 0; a < 4; ++a) {
      const int il = ty + 16 * a;
      const int64_t gi = i0 + il;
      const bool rowok = gi < N;
      if (w.want_dx)
        for (int d = 0; d < w.xcols; ++d) dxr[d] = 0.0;
      for (int b = 0; b < 4; ++b) {
        const int jl = tx + 16 * b;
        const int64_t gj = j0 + jl;
        if (!rowok || gj >= M) continue;
        if (w.sym_lower && gj > gi) continue;
        double wij;
        if (w.mode == W_GPR) {
          double bb = 0.0;
          for (int r = 0; r < w.R; ++r) bb = fma(bi[r * TILE + il], bj[r * TILE + jl], bb);
          wij = 0.5 * ((double)w.R * w.W[gi * w.ldw + gj] - bb);
          if (gi == gj) acc[pl.n_theta] += wij;          // tr W = d nlml / d noise
        } else {
          wij = w.W[gi * w.ldw + gj];
        }
        if (w.sym_lower && gj != gi) wij *= 2.0;
        const double* fi = sl + il * pl.S;
        const double* fj = sr + jl * pl.S;
        PrimEval ev[GPS_MAX_PRIMS];
        for (int p = 0; p < pl.n_prims; ++p) {
          const PrimC P = pl.prims[p];
          ev[p] = prim_eval(P, th, fi + P.feat_off, fj + P.feat_off);
          v[p] = ev[p].k;
        }
        int nslots_used = pl.n_prims;
        if (pl.n_ops) {
          program_fwd(pl, th, v);
          const OpC last = pl.ops[pl.n_ops - 1];
          nslots_used = last.dst + ((last.op == GPS_OP_LINEAR || last.op == GPS_OP_PRODUCT) ? last.n : 1);
        }
        for (int s = 0; s < nslots_used; ++s) vb[s] = 0.0;
        vb[pl.out_slot] = wij;
        program_bwd(pl, th, v, vb, acc);
        for (int p = 0; p < pl.n_prims; ++p) {
          const PrimC P = pl.prims[p];
          const double g = vb[p];
          const double* t = th + P.theta_off;
          const double* ti = tinv + P.theta_off;
          const double* fip = fi + P.feat_off;
          const double* fjp = fj + P.feat_off;
          if (is_stationary(P.type)) {
            acc[P.theta_off] += g * ev[p].k * ti[0];
            const double G = g * ev[p].dk;               // dObj / d(d2)
            if (P.ard) {
              for (int k = 0; k < P.ndims; ++k) {
                double df = fip[k] - fjp[k];
                acc[P.theta_off + 1 + k] += G * (-2.0) * df * df * ti[1 + k];
                if (w.want_dx) dxr[pd.dims[p][k]] += 2.0 * G * df * ti[1 + k];
              }
            } else {
              acc[P.theta_off + 1] += G * (-2.0) * ev[p].d2 * ti[1];
              if (w.want_dx)
                for (int k = 0; k < P.ndims; ++k)
                  dxr[pd.dims[p][k]] += 2.0 * G * (fip[k] - fjp[k]) * ti[1];
            }
          } else if (P.type == GPS_LINEAR) {
            for (int k = 0; k < P.ndims; ++k) {
              acc[P.theta_off + (P.ard ? k : 0)] += g * fip[k] * fjp[k];
              if (w.want_dx) dxr[pd.dims[p][k]] += g * t[P.ard ? k : 0] * fjp[k];
            }
          } else {
            const int nd = P.ndims;
            const double kk = ev[p].k, r = ev[p].dk, ils = ti[1], iper = ti[2];
            acc[P.theta_off] += g * kk * ti[0];
            acc[P.theta_off + 1] += g * kk * r * ils;
            double dcs = 0.0;   // per * d(sum cos)/dp
            const double gx = -g * kk * (0.5 * M_PI) * iper * ils * ils;
            for (int k = 0; k < nd; ++k) {
              double sind = fip[nd + k] * fjp[k] - fip[k] * fjp[nd + k];   // sin(a_i - a_j)
              dcs += sind * (fip[2 * nd + k] - fjp[2 * nd + k]);
              if (w.want_dx) dxr[pd.dims[p][k]] += gx * sind;
            }
            acc[P.theta_off + 2] += g * kk * dcs * iper * (0.25 * ils * ils);
          }
        }
      }
      if (w.want_dx) {
        // the 16 tx-lanes of a half warp share row il
        for (int d = 0; d < w.xcols; ++d) {
          double s = dxr[d];
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (tx == 0) dxs[il * w.xcols + d] += s;
        }
      }
    }
  }
  __syncthreads();
  // deterministic CTA reduction of the theta accumulators
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = 0; t < nacc; ++t) {
    double s = warp_sum(acc[t]);
    if (lane == 0) red[warp * nacc + t] = s;
  }
  __syncthreads();
  const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int t = tid; t < nacc; t += GRAM_THREADS) {
    double s = 0.0;
    for (int wq = 0; wq < GRAM_THREADS / 32; ++wq) s += red[wq * nacc + t];
    part_theta[cta * nacc + t] = s;
  }
  if (w.want_dx) {
    for (int idx = tid; idx < TILE * w.xcols; idx += GRAM_THREADS) {
      int r = idx / w.xcols;
      if (i0 + r < N)
        part_dx[((int64_t)blockIdx.x * N + i0 + r) * w.xcols + (idx - r * w.xcols)] = dxs[idx];
    }
  }
}


// --------------------------------------------------------------------------- backward, v2
// EXPERIMENTAL (gps_set_option("gram_impl", 2); not the default until measured on a B200).
// Same arithmetic as gram_bwd_kernel, different storage: the interpreter's per-thread arrays
// (theta-gradient accumulators, slot values and slot adjoints) are indexed by RUNTIME program
// data, so in gram_bwd_kernel they live in local memory -- ~120 read-modify-writes of local
// memory per matrix element for the NKN config, which is what makes that config spend more
// time in the Gram backward than in the factorisation.  Here they live in SHARED memory in a
// [index][thread] layout: consecutive threads touch consecutive 8-byte words, so every access
// is a conflict-free LDS/STS, whatever the (warp-uniform) index is.  To make room the CTA is
// 128 threads on a 32 x 32 tile (8 elements per thread).
constexpr int T2 = 32;          // tile edge
constexpr int NT2 = 128;        // threads per CTA

struct SmemSlots {
  double* base;                 // [n][nthreads]
  int tid, nthreads;
  __device__ __forceinline__ double& operator[](int i) const { return base[i * nthreads + tid]; }
};

__device__ __forceinline__ void program_fwd_s(const Plan& pl, const double* __restrict__ th, const SmemSlots& v) {
  for (int q = 0; q < pl.n_ops; ++q) {
    const OpC o = pl.ops[q];
    switch (o.op) {
      case GPS_OP_CONST: v[o.dst] = th[o.a]; break;
      case GPS_OP_ADD: v[o.dst] = v[o.a] + v[o.b]; break;
      case GPS_OP_MUL: v[o.dst] = v[o.a] * v[o.b]; break;
      case GPS_OP_COPY: v[o.dst] = v[o.a]; break;
      case GPS_OP_LINEAR:
        for (int r = 0; r < o.n; ++r) {
          double s = 0.0;
          for (int c = 0; c < o.b; ++c) s = fma(v[o.a + c], th[o.c + r * o.b + c], s);
          v[o.dst + r] = s + th[o.d + r];
        }
        break;
      case GPS_OP_PRODUCT:
        for (int g = 0; g < o.n; ++g) {
          double s = v[o.a + g * o.b];
          for (int c = 1; c < o.b; ++c) s *= v[o.a + g * o.b + c];
          v[o.dst + g] = s;
        }
        break;
    }
  }
}

__device__ __forceinline__ void program_bwd_s(const Plan& pl, const double* __restrict__ th, const SmemSlots& v,
                                              const SmemSlots& vb, const SmemSlots& acc) {
  for (int q = pl.n_ops - 1; q >= 0; --q) {
    const OpC o = pl.ops[q];
    switch (o.op) {
      case GPS_OP_CONST: acc[o.a] += vb[o.dst]; break;
      case GPS_OP_ADD: { double g = vb[o.dst]; vb[o.a] += g; vb[o.b] += g; } break;
      case GPS_OP_MUL: {
        double g = vb[o.dst];
        double va = v[o.a], vbv = v[o.b];
        vb[o.a] += g * vbv;
        vb[o.b] += g * va;
      } break;
      case GPS_OP_COPY: vb[o.a] += vb[o.dst]; break;
      case GPS_OP_LINEAR:
        for (int r = 0; r < o.n; ++r) {
          double g = vb[o.dst + r];
          acc[o.d + r] += g;
          for (int c = 0; c < o.b; ++c) {
            acc[o.c + r * o.b + c] = fma(g, v[o.a + c], acc[o.c + r * o.b + c]);
            vb[o.a + c] = fma(g, th[o.c + r * o.b + c], vb[o.a + c]);
          }
        }
        break;
      case GPS_OP_PRODUCT:
        for (int g = 0; g < o.n; ++g) {
          double gg = vb[o.dst + g];
          for (int c = 0; c < o.b; ++c) {
            double s = gg;
            for (int c2 = 0; c2 < o.b; ++c2)
              if (c2 != c) s *= v[o.a + g * o.b + c2];
            vb[o.a + g * o.b + c] += s;
          }
        }
        break;
    }
  }
}

// forward twin: gram_fwd_kernel with the slot values in shared memory ([slot][thread], 256 threads)
__global__ void __launch_bounds__(GRAM_THREADS)
gram_fwd_smem_kernel(const Plan pl, const int nslots, const double* __restrict__ theta,
                     const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                     double diag_add, int sym, int uplo, double* __restrict__ K, int64_t ldk) {
  extern __shared__ double sm[];
  double* th = sm;
  double* sl = th + pl.n_theta;
  double* sr = sl + TILE * pl.S;
  double* vbase = sr + TILE * pl.S;                // [nslots][GRAM_THREADS]
  const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
  if (sym && uplo && j0 > i0 + TILE - 1) return;
  const int tid = threadIdx.x;
  const SmemSlots v{vbase, tid, GRAM_THREADS};
  for (int t = tid; t < pl.n_theta; t += GRAM_THREADS) th[t] = theta[t];
  for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
    sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  for (int a = 0; a < 4; ++a) {
    const int il = ty + 16 * a;
    const int64_t gi = i0 + il;
    if (gi >= N) continue;
    for (int b = 0; b < 4; ++b) {
      const int jl = tx + 16 * b;
      const int64_t gj = j0 + jl;
      if (gj >= M) continue;
      if (sym && uplo && gj > gi) continue;
      for (int p = 0; p < pl.n_prims; ++p) {
        const PrimC P = pl.prims[p];
        v[p] = prim_eval(P, th, sl + il * pl.S + P.feat_off, sr + jl * pl.S + P.feat_off).k;
      }
      program_fwd_s(pl, th, v);
      double out = v[pl.out_slot];
      if (sym && gi == gj) out += diag_add;
      K[gi * ldk + gj] = out;
    }
  }
  (void)nslots;
}

// number of slots a program touches (primitives + op results)
int plan_nslots(const Plan& pl) {
  int n = pl.n_prims;
  if (pl.n_ops) {
    const OpC last = pl.ops[pl.n_ops - 1];
    n = last.dst + ((last.op == GPS_OP_LINEAR || last.op == GPS_OP_PRODUCT) ? last.n : 1);
  }
  return n;
}

// doubles of dynamic shared memory gram_bwd_smem_kernel needs
size_t bwd_smem_doubles(const Plan& pl, int nslots, int mode, int R, int want_dx, int xcols) {
  const size_t nacc = pl.n_theta + 1;
  return (size_t)pl.n_theta + 2 * (size_t)T2 * pl.S + (mode == W_GPR ? 2 * (size_t)R * T2 : 0) +
         (want_dx ? (size_t)T2 * xcols : 0) + nacc * NT2 + 2 * (size_t)nslots * NT2;
}

__global__ void __launch_bounds__(NT2)
gram_bwd_smem_kernel(const Plan pl, const PlanDims pd, const int nslots, const double* __restrict__ theta,
                     const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                     const BwdArgs w, double* __restrict__ part_theta, double* __restrict__ part_dx) {
  extern __shared__ double sm[];
  const int nacc = pl.n_theta + 1;
  const int tid = threadIdx.x;
  double* th = sm;                                 // [n_theta]
  double* sl = th + pl.n_theta;                    // [T2][S]
  double* sr = sl + T2 * pl.S;                     // [T2][S]
  double* bi = sr + T2 * pl.S;                     // [R][T2]   beta rows (i side)
  double* bj = bi + (w.mode == W_GPR ? w.R * T2 : 0);
  double* dxs = bj + (w.mode == W_GPR ? w.R * T2 : 0);   // [T2][xcols]
  double* accb = dxs + (w.want_dx ? T2 * w.xcols : 0);   // [nacc][NT2]
  double* vbase = accb + nacc * NT2;                     // [nslots][NT2]
  double* vbbase = vbase + nslots * NT2;                 // [nslots][NT2]
  const SmemSlots acc{accb, tid, NT2}, v{vbase, tid, NT2}, vb{vbbase, tid, NT2};
  const int tx = tid & 15, ty = tid >> 4;                // ty in 0..7
  // lower-triangular problems: row tile i visits i / njc column tiles, so the heavy row tiles
  // are scheduled FIRST (CTAs are dispatched in blockIdx order) and the light ones fill the tail
  const int64_t itile = w.sym_lower ? (int64_t)gridDim.y - 1 - blockIdx.y : (int64_t)blockIdx.y;
  const int64_t i0 = itile * T2;
  const int64_t jtiles = (M + T2 - 1) / T2;

  for (int t = 0; t < nacc; ++t) acc[t] = 0.0;
  for (int t = tid; t < pl.n_theta; t += NT2) th[t] = theta[t];
  for (int idx = tid; idx < T2 * pl.FT; idx += NT2) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
  }
  if (w.mode == W_GPR)
    for (int idx = tid; idx < w.R * T2; idx += NT2) {
      int r = idx / T2, c = idx - r * T2;
      bi[idx] = (i0 + c < N) ? w.beta[(int64_t)r * N + i0 + c] : 0.0;
    }
  if (w.want_dx)
    for (int idx = tid; idx < T2 * w.xcols; idx += NT2) dxs[idx] = 0.0;

  double dxr[GPS_MAX_DIMS];
  PrimEval ev[GPS_MAX_PRIMS];

  for (int64_t jt = blockIdx.x; jt < jtiles; jt += w.njc) {
    const int64_t j0 = jt * T2;
    if (w.sym_lower && j0 > i0 + T2 - 1) break;
    __syncthreads();   // previous tile fully consumed (and the prologue's smem writes visible)
    for (int idx = tid; idx < T2 * pl.FT; idx += NT2) {
      int r = idx / pl.FT, c = idx - r * pl.FT;
      sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
    }
    if (w.mode == W_GPR)
      for (int idx = tid; idx < w.R * T2; idx += NT2) {
        int r = idx / T2, c = idx - r * T2;
        bj[idx] = (j0 + c < M) ? w.beta[(int64_t)r * N + j0 + c] : 0.0;
      }
    __syncthreads();
    for (int a = 0; a < 4; ++a) {
      const int il = ty + 8 * a;
      const int64_t gi = i0 + il;
      const bool rowok = gi < N;
      if (w.want_dx)
        for (int d = 0; d < w.xcols; ++d) dxr[d] = 0.0;
      for (int b = 0; b < 2; ++b) {
        const int jl = tx + 16 * b;
        const int64_t gj = j0 + jl;
        if (!rowok || gj >= M) continue;
        if (w.sym_lower && gj > gi) continue;
        double wij;
        if (w.mode == W_GPR) {
          double bb = 0.0;
          for (int r = 0; r < w.R; ++r) bb = fma(bi[r * T2 + il], bj[r * T2 + jl], bb);
          wij = 0.5 * ((double)w.R * w.W[gi * w.ldw + gj] - bb);
          if (gi == gj) acc[pl.n_theta] += wij;          // tr W = d nlml / d noise
        } else {
          wij = w.W[gi * w.ldw + gj];
        }
        if (w.sym_lower && gj != gi) wij *= 2.0;
        const double* fi = sl + il * pl.S;
        const double* fj = sr + jl * pl.S;
        for (int p = 0; p < pl.n_prims; ++p) {
          const PrimC P = pl.prims[p];
          ev[p] = prim_eval(P, th, fi + P.feat_off, fj + P.feat_off);
          v[p] = ev[p].k;
        }
        if (pl.n_ops) program_fwd_s(pl, th, v);
        for (int s = 0; s < nslots; ++s) vb[s] = 0.0;
        vb[pl.out_slot] = wij;
        program_bwd_s(pl, th, v, vb, acc);
        for (int p = 0; p < pl.n_prims; ++p) {
          const PrimC P = pl.prims[p];
          const double g = vb[p];
          const double* t = th + P.theta_off;
          const double* fip = fi + P.feat_off;
          const double* fjp = fj + P.feat_off;
          if (is_stationary(P.type)) {
            acc[P.theta_off] += g * ev[p].k / t[0];
            const double G = g * ev[p].dk;               // dObj / d(d2)
            if (P.ard) {
              for (int k = 0; k < P.ndims; ++k) {
                double df = fip[k] - fjp[k];
                acc[P.theta_off + 1 + k] += G * (-2.0) * df * df / t[1 + k];
                if (w.want_dx) dxr[pd.dims[p][k]] += 2.0 * G * df / t[1 + k];
              }
            } else {
              acc[P.theta_off + 1] += G * (-2.0) * ev[p].d2 / t[1];
              if (w.want_dx)
                for (int k = 0; k < P.ndims; ++k)
                  dxr[pd.dims[p][k]] += 2.0 * G * (fip[k] - fjp[k]) / t[1];
            }
          } else if (P.type == GPS_LINEAR) {
            for (int k = 0; k < P.ndims; ++k) {
              acc[P.theta_off + (P.ard ? k : 0)] += g * fip[k] * fjp[k];
              if (w.want_dx) dxr[pd.dims[p][k]] += g * t[P.ard ? k : 0] * fjp[k];
            }
          } else {
            const int nd = P.ndims;
            const double kk = ev[p].k, r = ev[p].dk, ls = t[1], per = t[2];
            acc[P.theta_off] += g * kk / t[0];
            acc[P.theta_off + 1] += g * kk * r / ls;
            double dcs = 0.0;   // d(sum cos)/dp
            for (int k = 0; k < nd; ++k) {
              double sind = fip[nd + k] * fjp[k] - fip[k] * fjp[nd + k];   // sin(a_i - a_j)
              dcs += sind * (fip[2 * nd + k] - fjp[2 * nd + k]) / per;
              if (w.want_dx) dxr[pd.dims[p][k]] += -g * kk * sind * M_PI / (2.0 * per * ls * ls);
            }
            acc[P.theta_off + 2] += g * kk * dcs / (4.0 * ls * ls);
          }
        }
      }
      if (w.want_dx) {
        // the 16 tx-lanes of a half warp share row il
        for (int d = 0; d < w.xcols; ++d) {
          double s = dxr[d];
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (tx == 0) dxs[il * w.xcols + d] += s;
        }
      }
    }
  }
  __syncthreads();
  // deterministic CTA reduction: thread t sums the NT2 per-thread partials of parameter t,
  // starting at its own column so that the reads of a warp fall into different banks
  const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int t = tid; t < nacc; t += NT2) {
    double s = 0.0;
    for (int k = 0; k < NT2; ++k) s += accb[t * NT2 + ((k + tid) & (NT2 - 1))];
    part_theta[cta * nacc + t] = s;
  }
  if (w.want_dx) {
    for (int idx = tid; idx < T2 * w.xcols; idx += NT2) {
      int r = idx / w.xcols;
      if (i0 + r < N)
        part_dx[((int64_t)blockIdx.x * N + i0 + r) * w.xcols + (idx - r * w.xcols)] = dxs[idx];
    }
  }
}

// --------------------------------------------------------------------------- fast path
// One stationary primitive and no composition program (RBF / Matern / Exponential, ARD or not,
// <= 16 active dimensions) -- the covariance of every GPR / SVGP config in BASELINE.json except
// the NKN one.  Same arithmetic as the interpreter above (same feature vectors, same summation
// order over the dimensions) but with compile-time bounds, so the 4 x 4 micro-tile, the dot
// products and the gradient accumulators live in registers instead of local memory.
//   CTA: 64 x 64 tile, 256 threads; thread (ty, tx) owns rows {2ty, 2ty+1, 32+2ty, 33+2ty} and
//   columns {2tx, 2tx+1, 32+2tx, 33+2tx}: feature reads are 16-byte shared loads (k-major smem),
//   K / W accesses are 16-byte global accesses, 256 contiguous bytes per half-warp.
constexpr int SLD = TILE + 2;   // smem row stride of a k-major feature tile (16-byte aligned rows)

template <int DP>
__device__ __forceinline__ void stat_load_tile(const double* __restrict__ F, int FT, int nd, int64_t r0,
                                               int64_t nrows, double (*sf)[SLD], double* ss, int tid) {
  for (int idx = tid; idx < TILE * (DP + 1); idx += GRAM_THREADS) {
    const int r = idx / (DP + 1), c = idx - r * (DP + 1);
    const bool ok = r0 + r < nrows;
    if (c < DP) sf[c][r] = (ok && c < nd) ? F[(r0 + r) * FT + c] : 0.0;
    else ss[r] = ok ? F[(r0 + r) * FT + nd] : 0.0;
  }
}

template <int DP>
__device__ __forceinline__ void stat_dots(const double (*sl)[SLD], const double (*sr)[SLD], int ty, int tx,
                                          double (&dot)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dot[a][b] = 0.0;
#pragma unroll
  for (int k = 0; k < DP; ++k) {
    const double2 a01 = *reinterpret_cast<const double2*>(&sl[k][2 * ty]);
    const double2 a23 = *reinterpret_cast<const double2*>(&sl[k][32 + 2 * ty]);
    const double2 b01 = *reinterpret_cast<const double2*>(&sr[k][2 * tx]);
    const double2 b23 = *reinterpret_cast<const double2*>(&sr[k][32 + 2 * tx]);
    const double fa[4] = {a01.x, a01.y, a23.x, a23.y};
    const double fb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dot[a][b] = fma(fa[a], fb[b], dot[a][b]);
  }
}

template <int DP>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_fwd_stat_kernel(int type, int nd, int FT, const double* __restrict__ theta,
                     const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                     double diag_add, int sym, int uplo, double* __restrict__ K, int64_t ldk) {
  __shared__ __align__(16) double sl[DP][SLD];
  __shared__ __align__(16) double sr[DP][SLD];
  __shared__ double ssl[TILE], ssr[TILE];
  const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
  if (sym && uplo && j0 > i0 + TILE - 1) return;
  const int tid = threadIdx.x;
  stat_load_tile<DP>(FL, FT, nd, i0, N, sl, ssl, tid);
  stat_load_tile<DP>(FR, FT, nd, j0, M, sr, ssr, tid);
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const double var = theta[0];
  double dot[4][4];
  stat_dots<DP>(sl, sr, ty, tx, dot);
  const bool vec = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int il = (a >> 1) * 32 + 2 * ty + (a & 1);
    const int64_t gi = i0 + il;
    if (gi >= N) continue;
#pragma unroll
    for (int bp = 0; bp < 2; ++bp) {
      const int jl = bp * 32 + 2 * tx;
      const int64_t gj = j0 + jl;
      double kv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double raw = -2.0 * dot[a][2 * bp + q] + (ssl[il] + ssr[jl + q]);   // kernels.py:413-414
        const bool live = raw >= 0.0;                                            // clip_by_value
        double kk, dk;
        stat_body(type, var, live ? raw : 0.0, live, kk, dk);
        if (sym && gi == gj + q) kk += diag_add;
        kv[q] = kk;
      }
      const bool ok0 = gj < M && !(sym && uplo && gj > gi);
      const bool ok1 = gj + 1 < M && !(sym && uplo && gj + 1 > gi);
      double* kp = K + gi * ldk + gj;
      if (vec && ok0 && ok1) {
        *reinterpret_cast<double2*>(kp) = make_double2(kv[0], kv[1]);
      } else {
        if (ok0) kp[0] = kv[0];
        if (ok1) kp[1] = kv[1];
      }
    }
  }
}

// Backward of the same: acc layout = [d/d variance, d/d lengthscale(s), trace W]; per-CTA partials
// in the layout of gram_bwd_kernel, so the fixed-order second pass is shared.
template <int DP>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_bwd_stat_kernel(int type, int ard, int nd, int FT, int n_theta, const double* __restrict__ theta,
                     const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                     const BwdArgs w, double* __restrict__ part_theta, double* __restrict__ Gout, int64_t ldg) {
  __shared__ __align__(16) double sl[DP][SLD];
  __shared__ __align__(16) double sr[DP][SLD];
  __shared__ double ssl[TILE], ssr[TILE];
  __shared__ double bi[MAX_R][TILE], bj[MAX_R][TILE];
  __shared__ double red[GRAM_THREADS / 32][DP + 2];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * TILE;
  const int64_t jtiles = (M + TILE - 1) / TILE;
  const double var = theta[0];
  const bool vecw = ((w.ldw & 1) == 0) && ((reinterpret_cast<uintptr_t>(w.W) & 15) == 0);

  double acc_v = 0.0, acc_tr = 0.0, acc_l[DP];
#pragma unroll
  for (int k = 0; k < DP; ++k) acc_l[k] = 0.0;

  stat_load_tile<DP>(FL, FT, nd, i0, N, sl, ssl, tid);
  if (w.mode == W_GPR)
    for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
      const int r = idx / TILE, c = idx - r * TILE;
      bi[r][c] = (i0 + c < N) ? w.beta[(int64_t)r * N + i0 + c] : 0.0;
    }

  for (int64_t jt = blockIdx.x; jt < jtiles; jt += w.njc) {
    const int64_t j0 = jt * TILE;
    if (w.sym_lower && j0 > i0 + TILE - 1) break;
    __syncthreads();   // previous tile fully consumed
    stat_load_tile<DP>(FR, FT, nd, j0, M, sr, ssr, tid);
    if (w.mode == W_GPR)
      for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
        const int r = idx / TILE, c = idx - r * TILE;
        bj[r][c] = (j0 + c < M) ? w.beta[(int64_t)r * N + j0 + c] : 0.0;
      }
    __syncthreads();
    double dot[4][4];
    stat_dots<DP>(sl, sr, ty, tx, dot);
    double G[4][4];      // dObj / d(d2) per element (0 where masked)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = (a >> 1) * 32 + 2 * ty + (a & 1);
      const int64_t gi = i0 + il;
#pragma unroll
      for (int bp = 0; bp < 2; ++bp) {
        const int jl = bp * 32 + 2 * tx;
        const int64_t gj = j0 + jl;
        const bool ok0 = gi < N && gj < M && !(w.sym_lower && gj > gi);
        const bool ok1 = gi < N && gj + 1 < M && !(w.sym_lower && gj + 1 > gi);
        double wv[2] = {0.0, 0.0};
        const double* wp = w.W + gi * w.ldw + gj;
        if (vecw && ok0 && ok1) {
          const double2 t2 = *reinterpret_cast<const double2*>(wp);
          wv[0] = t2.x; wv[1] = t2.y;
        } else {
          if (ok0) wv[0] = wp[0];
          if (ok1) wv[1] = wp[1];
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const bool ok = q ? ok1 : ok0;
          double wij = wv[q];
          if (w.mode == W_GPR) {
            double bb = 0.0;
            for (int r = 0; r < w.R; ++r) bb = fma(bi[r][il], bj[r][jl + q], bb);
            wij = 0.5 * ((double)w.R * wij - bb);
            if (ok && gi == gj + q) acc_tr += wij;            // tr W = d nlml / d noise
          }
          if (w.sym_lower && gj + q != gi) wij *= 2.0;
          const double raw = -2.0 * dot[a][2 * bp + q] + (ssl[il] + ssr[jl + q]);
          const bool live = raw >= 0.0;
          const double d2 = live ? raw : 0.0;
          double kk, dk;
          stat_body(type, var, d2, live, kk, dk);
          const double g = ok ? wij : 0.0;
          acc_v = fma(g, kk, acc_v);
          const double Gq = g * dk;
          G[a][2 * bp + q] = Gq;
          if (!ard) acc_l[0] = fma(Gq, d2, acc_l[0]);
        }
        if (Gout) {      // dObj / d(d2), kept for the input gradient (one skinny product afterwards)
          double* gp = Gout + gi * ldg + gj;
          if (ok0) gp[0] = G[a][2 * bp];
          if (ok1) gp[1] = G[a][2 * bp + 1];
        }
      }
    }
    if (ard) {
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(&sl[k][2 * ty]);
        const double2 a23 = *reinterpret_cast<const double2*>(&sl[k][32 + 2 * ty]);
        const double2 b01 = *reinterpret_cast<const double2*>(&sr[k][2 * tx]);
        const double2 b23 = *reinterpret_cast<const double2*>(&sr[k][32 + 2 * tx]);
        const double fa[4] = {a01.x, a01.y, a23.x, a23.y};
        const double fb[4] = {b01.x, b01.y, b23.x, b23.y};
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double df = fa[a] - fb[b];
            s = fma(G[a][b] * df, df, s);
          }
        acc_l[k] += s;
      }
    }
  }
  // constant factors: d k / d variance = k / variance; d d2 / d l_k = -2 (f_ik - f_jk)^2 / l_k
  // (features are x / l); non-ARD: d d2 / d l = -2 d2 / l
  acc_v /= var;
#pragma unroll
  for (int k = 0; k < DP; ++k) {
    const int nl = ard ? nd : 1;
    acc_l[k] = (k < nl) ? acc_l[k] * (-2.0) / theta[1 + k] : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  {
    double sv = warp_sum(acc_v), st = warp_sum(acc_tr);
    if (lane == 0) { red[warp][0] = sv; red[warp][DP + 1] = st; }
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      double sk = warp_sum(acc_l[k]);
      if (lane == 0) red[warp][1 + k] = sk;
    }
  }
  __syncthreads();
  const int nacc = n_theta + 1;
  const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int t = tid; t < nacc; t += GRAM_THREADS) {
    const int src = (t == n_theta) ? DP + 1 : t;          // [variance, l_0.., trace]
    double sacc = 0.0;
    for (int wq = 0; wq < GRAM_THREADS / 32; ++wq) sacc += red[wq][src];
    part_theta[cta * nacc + t] = sacc;
  }
}

// Input gradient of the fast path from G = dObj / d(d2) (N x M) and P = G [F_R | 1]  (N x (nd + 1)):
//   d d2_ij / d x_id = 2 (f_id - f_jd) / l_d   =>   dX[i][d] = scale * 2 / l_d * (f_id * sum_j G_ij - sum_j G_ij f_jd)
__global__ void stat_dx_finish_kernel(const double* __restrict__ P, int64_t ldp, const double* __restrict__ FL,
                                      int FT, int nd, int ard, const double* __restrict__ theta, PlanDims pd,
                                      int64_t N, int xcols, double scale, double* __restrict__ dX, int64_t lddx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * xcols) return;
  const int64_t i = idx / xcols;
  const int c = (int)(idx - i * xcols);
  double v = 0.0;
  for (int k = 0; k < nd; ++k)
    if (pd.dims[0][k] == c) {
      const double l = theta[1 + (ard ? k : 0)];
      v += scale * 2.0 / l * (FL[i * FT + k] * P[i * ldp + nd] - P[i * ldp + k]);
    }
  dX[i * lddx + c] = v;
}

// B[k][j] = F[j][k] for k < nd, B[nd][j] = 1   ((nd + 1) x M, the K-contiguous operand of P = G B^T)
__global__ void stat_dx_operand_kernel(const double* __restrict__ F, int FT, int nd, int64_t M,
                                       double* __restrict__ B, int64_t ldb) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  for (int k = 0; k < nd; ++k) B[(int64_t)k * ldb + j] = F[j * FT + k];
  B[(int64_t)nd * ldb + j] = 1.0;
}

// The product form of the input gradient, f_id sum_j G_ij - sum_j G_ij f_jd, cancels the j == i (or
// coincident-point) term only to rounding.  That is harmless where dk/d(d2) is bounded at zero
// distance (RBF, Matern 3/2, 5/2) and not for Matern 1/2 / Exponential, whose derivative is
// ~ 1 / sqrt(1e-12) there: those keep the interpreter, which multiplies by the exact difference.
bool stat_dx_ok(int type) { return type == GPS_RBF || type == GPS_MATERN32 || type == GPS_MATERN52; }

// single stationary primitive with the identity program and <= 16 active dimensions?
bool stat_fast(const gps_handle* h, const Plan& pl) {
  return h->gram_impl == 0 && pl.n_prims == 1 && pl.n_ops == 0 && pl.out_slot == 0 &&
         is_stationary(pl.prims[0].type) && pl.prims[0].ndims <= 16 && pl.prims[0].theta_off == 0;
}

// out[t] = sum_c part[c][t] (fixed order)
__global__ void reduce_cols_kernel(const double* __restrict__ part, int64_t nrows, int64_t ncols,
                                   double scale, double* __restrict__ out, int64_t out_stride_skip,
                                   int accumulate) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncols) return;
  double s = 0.0;
  for (int64_t c = 0; c < nrows; ++c) s += part[c * ncols + t];
  (void)out_stride_skip;
  out[t] = (accumulate ? out[t] : 0.0) + scale * s;
}

// dX[i][d] = scale * sum_jc part[jc][i][d]
__global__ void reduce_dx_kernel(const double* __restrict__ part, int njc, int64_t N, int xcols,
                                 double scale, double* __restrict__ dX, int64_t lddx) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * xcols) return;
  double s = 0.0;
  for (int c = 0; c < njc; ++c) s += part[(int64_t)c * N * xcols + idx];
  int64_t i = idx / xcols;
  dX[i * lddx + (idx - i * xcols)] = scale * s;
}

// --------------------------------------------------------------------------- NKN fast path
// Neural-kernel-network programs of the shape  Linear -> Product(2) -> Linear -> Product(2) -> Linear(->1)
// over <= 7 primitives (the reference's NKN wrapper, neural_kernel_network.py / wrapper.py:100-104, BASELINE
// config C3).  The interpreter kernels above keep slot values, slot adjoints and ~100 theta-gradient
// accumulators per THREAD in runtime-indexed arrays (local memory): ~120 read-modify-writes per matrix
// element.  Here a warp works on 8 matrix elements at a time ("octet"), 4 lanes per element:
//   * lane (lr, lc) = (lane >> 2, lane & 3) belongs to element lr and evaluates primitives lc and 4 + lc;
//   * the Linear layers are 8x8x4 FP64 tensor-core products (DMMA): rows = the 8 elements, columns =
//     layer outputs, reduction = layer inputs; lane (lr, lc) receives outputs 2 lc, 2 lc + 1 of element
//     lr, i.e. exactly one pair of the following Product layer;
//   * backward: adjoint mat-vecs with the transposed weights are DMMAs too, and the WEIGHT gradients are
//     accumulated as 8x8 DMMA tiles  dW[o][i] += sum_e dout[e][o] in[e][i]  (reduction over the 8 elements
//     of the octet, both operands staged through 1 KB of shared memory per warp; column `n_in` of `in` is
//     the constant 1, which makes that column of the tile the bias gradient);
//   * primitive parameter gradients stay in <= 18 registers per lane (two primitives x (1 + 8)).
// Everything is register resident; arithmetic per element is the interpreter's in another summation order.
constexpr int NKN_MAXD = 8;      // active dimensions per primitive handled in registers

struct NknPlan {
  int P, n1, n2, i2, i3;         // primitives; outputs of Linear 1 / 2; inputs of Linear 2 / 3 (n1 / 2, n2 / 2)
  int w1, b1, w2, b2, w3, b3;    // theta offsets of weights and biases
};

bool nkn_match(const gps_handle* h, const Plan& pl, NknPlan* nk) {
  if (h->gram_impl != 0 || pl.n_ops != 5 || pl.n_prims > 7) return false;
  const OpC* o = pl.ops;
  if (o[0].op != GPS_OP_LINEAR || o[1].op != GPS_OP_PRODUCT || o[2].op != GPS_OP_LINEAR ||
      o[3].op != GPS_OP_PRODUCT || o[4].op != GPS_OP_LINEAR)
    return false;
  const int P = pl.n_prims;
  if (o[0].a != 0 || o[0].b != P || o[0].n < 2 || o[0].n > 8 || (o[0].n & 1)) return false;
  if (o[1].a != o[0].dst || o[1].b != 2 || o[1].n * 2 != o[0].n) return false;
  if (o[2].a != o[1].dst || o[2].b != o[1].n || o[2].n < 2 || o[2].n > 8 || (o[2].n & 1)) return false;
  if (o[3].a != o[2].dst || o[3].b != 2 || o[3].n * 2 != o[2].n) return false;
  if (o[4].a != o[3].dst || o[4].b != o[3].n || o[4].n != 1 || pl.out_slot != o[4].dst) return false;
  for (int p = 0; p < P; ++p)
    if (pl.prims[p].ndims > NKN_MAXD) return false;
  nk->P = P; nk->n1 = o[0].n; nk->i2 = o[1].n; nk->n2 = o[2].n; nk->i3 = o[3].n;
  nk->w1 = o[0].c; nk->b1 = o[0].d; nk->w2 = o[2].c; nk->b2 = o[2].d; nk->w3 = o[4].c; nk->b3 = o[4].d;
  return true;
}

// Shared-memory row stride of the feature tiles for the NKN kernels: the 8 elements of an octet read 8
// consecutive rows at the same column, and the lanes of two primitives of the same type run together,
// so rows must land in distinct 8-byte banks of a half warp: stride = +-1 (mod 16).  (The first version
// used the interpreter's odd stride, 83 for the C3 network: ncu counted 1.1e8 bank conflicts at N=4096,
// profiles/r02_nkn_bwd_first_ncu_full.json.)
int nkn_row_stride(int ft) {
  int s = ft;
  while (s % 16 != 1 && s % 16 != 15) ++s;
  return s;
}

// per-lane DMMA operand fragments of the three Linear layers, zero padded to 8 x 8
struct NknFrag {
  double w1f[2], w2f, w3;        // B fragments (weights[out = lr][in = lc + 4 s]) of Linear 1 (two k-steps), Linear 2;
                                 // Linear 3 weight of input lc
  double c1[2], c2[2], b3;       // biases of outputs 2 lc, 2 lc + 1 of Linear 1 / 2; bias of Linear 3
  double w1t[2], w2t[2];         // B fragments of the transposes (weights[out = lc + 4 s][in = lr])
};

__device__ __forceinline__ void nkn_load_frag(const NknPlan& nk, const double* __restrict__ th, int lr, int lc,
                                              NknFrag& f) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int c = lc + 4 * s;
    f.w1f[s] = (lr < nk.n1 && c < nk.P) ? th[nk.w1 + lr * nk.P + c] : 0.0;
    f.w1t[s] = (c < nk.n1 && lr < nk.P) ? th[nk.w1 + c * nk.P + lr] : 0.0;
    f.w2t[s] = (c < nk.n2 && lr < nk.i2) ? th[nk.w2 + c * nk.i2 + lr] : 0.0;
    const int o = 2 * lc + s;
    f.c1[s] = o < nk.n1 ? th[nk.b1 + o] : 0.0;
    f.c2[s] = o < nk.n2 ? th[nk.b2 + o] : 0.0;
  }
  f.w2f = (lr < nk.n2 && lc < nk.i2) ? th[nk.w2 + lr * nk.i2 + lc] : 0.0;
  f.w3 = lc < nk.i3 ? th[nk.w3 + lc] : 0.0;
  f.b3 = th[nk.b3];
}

// A primitive's descriptor as plain ints.  PrimC packs int16 fields in constant memory, fetched by a per-lane
// index: in the first version every use inside the octet loop was a constant load + a sign extension (6.6 %
// of its instructions).  MEASURED AFTERWARDS (profiles/r02_nkn_bwd_source_hotspots.txt): under the
// 128-register cap of two CTAs per SM the compiler still re-fetches the fields inside the loop (13 % of the
// tuned version's instructions sit on nkn_prim) -- the next step is one packed 32-bit word per primitive,
// or the descriptors of the CTA in shared memory.
struct PrimR {
  int type, ard, ndims, theta_off, feat_off;
};
__device__ __forceinline__ PrimR nkn_prim(const Plan& pl, int p) {
  const PrimC c = pl.prims[p];
  PrimR r;
  r.type = c.type; r.ard = c.ard; r.ndims = c.ndims; r.theta_off = c.theta_off; r.feat_off = c.feat_off;
  return r;
}

// prim_eval for the NKN kernels: the loops over the (<= NKN_MAXD) active dimensions are unrolled under a
// guard (the generic loops spend half of their instructions on counters, selects and branches) and the
// periodic kernel multiplies by 1 / lengthscale^2 where prim_eval divides.
__device__ __forceinline__ PrimEval nkn_prim_eval(const PrimR& P, const double* __restrict__ th,
                                                  const double* __restrict__ tinv,
                                                  const double* __restrict__ fi, const double* __restrict__ fj) {
  PrimEval e;
  e.k = 0; e.dk = 0; e.d2 = 0;
  const double* t = th + P.theta_off;
  const int nd = P.ndims;
  if (is_stationary(P.type)) {
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < NKN_MAXD; ++k)
      if (k < nd) dot = fma(fi[k], fj[k], dot);
    double raw = -2.0 * dot + (fi[nd] + fj[nd]);
    bool live = raw >= 0.0;
    double d2 = live ? raw : 0.0;
    e.d2 = d2;
    stat_body(P.type, t[0], d2, live, e.k, e.dk);
  } else if (P.type == GPS_LINEAR) {
    double dot = 0.0;
    if (P.ard) {
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < nd) dot = fma(fi[k] * t[k], fj[k], dot);
    } else {
      const double v = t[0];
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < nd) dot = fma(fi[k] * v, fj[k], dot);
    }
    e.k = dot;
  } else {
    double cs = 0.0;
#pragma unroll
    for (int k = 0; k < NKN_MAXD; ++k)
      if (k < nd) cs += fi[k] * fj[k] + fi[nd + k] * fj[nd + k];
    const double ils = tinv[P.theta_off + 1];
    double r = 0.5 * ((double)nd - cs) * (ils * ils);
    e.k = t[0] * exp(-0.5 * r);
    e.dk = r;
  }
  return e;
}

// forward pass of one octet.  k0 / k1: this lane's two primitive values of its element (0 where the lane
// has no primitive).  Returns the network output of element lr (on all four lanes of the quad) and leaves
// the layer outputs this lane owns in o1[2] (Linear 1), o2[2] (Linear 2), h1, h2 (the Product layers).
__device__ __forceinline__ double nkn_forward(const NknFrag& f, double k0, double k1, double* o1, double* o2,
                                              double& h1, double& h2) {
  o1[0] = f.c1[0]; o1[1] = f.c1[1];
  dmma884(o1[0], o1[1], k0, f.w1f[0]);
  dmma884(o1[0], o1[1], k1, f.w1f[1]);
  h1 = o1[0] * o1[1];
  o2[0] = f.c2[0]; o2[1] = f.c2[1];
  dmma884(o2[0], o2[1], h1, f.w2f);
  h2 = o2[0] * o2[1];
  double part = f.w3 * h2;
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  return part + f.b3;
}

__global__ void __launch_bounds__(GRAM_THREADS)
gram_fwd_nkn_kernel(const Plan pl, const NknPlan nk, const double* __restrict__ theta,
                    const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                    double diag_add, int sym, int uplo, double* __restrict__ K, int64_t ldk) {
  extern __shared__ double sm[];
  double* th = sm;
  double* tinv = th + pl.n_theta;
  double* sl = tinv + pl.n_theta;
  double* sr = sl + TILE * pl.S;
  const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
  if (sym && uplo && j0 > i0 + TILE - 1) return;
  const int tid = threadIdx.x;
  for (int t = tid; t < pl.n_theta; t += GRAM_THREADS) {
    th[t] = theta[t];
    tinv[t] = 1.0 / theta[t];
  }
  for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
    sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  NknFrag f;
  nkn_load_frag(nk, th, lr, lc, f);
  const bool has0 = lc < nk.P, has1 = 4 + lc < nk.P;
  const PrimR P0 = nkn_prim(pl, has0 ? lc : 0), P1 = nkn_prim(pl, has1 ? 4 + lc : 0);
  for (int q = 0; q < TILE; ++q) {
    const int il = warp * 8 + (q >> 3), jb = (q & 7) * 8, jl = jb + lr;
    const int64_t gi = i0 + il, gj = j0 + jl;
    if (gi >= N || j0 + jb >= M) continue;                        // warp uniform
    if (sym && uplo && j0 + jb > gi) continue;
    const double* fi = sl + il * pl.S;
    const double* fj = sr + jl * pl.S;
    double k0 = 0.0, k1 = 0.0;
    if (has0) k0 = nkn_prim_eval(P0, th, tinv, fi + P0.feat_off, fj + P0.feat_off).k;
    if (has1) k1 = nkn_prim_eval(P1, th, tinv, fi + P1.feat_off, fj + P1.feat_off).k;
    double o1[2], o2[2], h1, h2;
    double out = nkn_forward(f, k0, k1, o1, o2, h1, h2);
    if (lc == 0 && gj < M && !(sym && uplo && gj > gi)) {
      if (sym && gi == gj) out += diag_add;
      K[gi * ldk + gj] = out;
    }
  }
}

// parameter gradient of one primitive at one element: acc[q] += g * d k / d theta[theta_off + q]
__device__ __forceinline__ void nkn_prim_grad(const PrimR& P, const double* __restrict__ fi,
                                              const double* __restrict__ fj, const PrimEval& ev, double g,
                                              double* acc) {
  // the factors that depend on theta only (1 / variance, 1 / lengthscale ...) are applied once per lane
  // after the loop over the matrix elements: nkn_prim_scale
  const int nd = P.ndims;
  if (is_stationary(P.type)) {
    acc[0] += g * ev.k;
    const double G = -2.0 * g * ev.dk;
    if (P.ard) {
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < nd) {
          double df = fi[k] - fj[k];
          acc[1 + k] += G * df * df;
        }
    } else {
      acc[1] += G * ev.d2;
    }
  } else if (P.type == GPS_LINEAR) {
    if (P.ard) {
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < nd) acc[k] += g * fi[k] * fj[k];
    } else {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < nd) s = fma(fi[k], fj[k], s);
      acc[0] += g * s;
    }
  } else {
    const double gk = g * ev.k;
    acc[0] += gk;
    acc[1] += gk * ev.dk;
    double dcs = 0.0;
#pragma unroll
    for (int k = 0; k < NKN_MAXD; ++k)
      if (k < nd) {
        double sind = fi[nd + k] * fj[k] - fi[k] * fj[nd + k];    // sin(a_i - a_j)
        dcs += sind * (fi[2 * nd + k] - fj[2 * nd + k]);
      }
    acc[2] += gk * dcs;
  }
}

// theta-only factors of the sums nkn_prim_grad accumulated (same formulas as gram_bwd_kernel)
__device__ __forceinline__ void nkn_prim_scale(const PrimR& P, const double* __restrict__ tinv, double* acc) {
  const double* ti = tinv + P.theta_off;
  if (is_stationary(P.type)) {
    acc[0] *= ti[0];
    if (P.ard) {
#pragma unroll
      for (int k = 0; k < NKN_MAXD; ++k)
        if (k < P.ndims) acc[1 + k] *= ti[1 + k];
    } else {
      acc[1] *= ti[1];
    }
  } else if (P.type == GPS_PERIODIC) {
    const double ils = ti[1], iper = ti[2];
    acc[0] *= ti[0];
    acc[1] *= ils;
    acc[2] *= iper * (0.25 * ils * ils);
  }
}

__device__ __forceinline__ int nkn_prim_nparams(const PrimR& P) {
  return is_stationary(P.type) ? 1 + (P.ard ? P.ndims : 1) : P.type == GPS_LINEAR ? (P.ard ? P.ndims : 1) : 3;
}

__global__ void __launch_bounds__(GRAM_THREADS, 2)
gram_bwd_nkn_kernel(const Plan pl, const NknPlan nk, const double* __restrict__ theta,
                    const double* __restrict__ FL, const double* __restrict__ FR, int64_t N, int64_t M,
                    const BwdArgs w, double* __restrict__ part_theta) {
  extern __shared__ double sm[];
  const int nacc = pl.n_theta + 1;
  double* th = sm;                                 // [n_theta]
  double* tinv = th + pl.n_theta;                  // [n_theta]
  double* bi = tinv + pl.n_theta;                  // [R][TILE]
  double* bj = bi + (w.mode == W_GPR ? w.R * TILE : 0);
  double* stage = bj + (w.mode == W_GPR ? w.R * TILE : 0);   // [8 warps][2][8][8]
  double* sl = stage + 8 * 128;                    // [TILE][S]
  double* sr = sl + TILE * pl.S;                   // [TILE][S]
  double* red = sl;                                // [8][nacc], after the main loop (the host sizes the
                                                   // region as max(2 TILE S, 8 nacc))
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3, qb = lane & ~3;
  double* Ts = stage + warp * 128;                 // adjoints of a layer's outputs   [element][output]
  double* Hs = Ts + 64;                            // the layer's inputs and a 1      [element][input]
  // row tiles in descending order: with lower-triangular weights the last row tile has the most
  // column tiles, and the hardware hands out CTAs in grid order -- heavy ones first, light ones fill the tail
  const int64_t i0 = (int64_t)(gridDim.y - 1 - blockIdx.y) * TILE;
  const int64_t jtiles = (M + TILE - 1) / TILE;

  for (int t = tid; t < pl.n_theta; t += GRAM_THREADS) {
    th[t] = theta[t];
    tinv[t] = 1.0 / theta[t];
  }
  for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
    int r = idx / pl.FT, c = idx - r * pl.FT;
    sl[r * pl.S + c] = (i0 + r < N) ? FL[(i0 + r) * pl.FT + c] : 0.0;
  }
  if (w.mode == W_GPR)
    for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
      int r = idx / TILE, c = idx - r * TILE;
      bi[idx] = (i0 + c < N) ? w.beta[(int64_t)r * N + i0 + c] : 0.0;
    }
  __syncthreads();
  NknFrag f;
  nkn_load_frag(nk, th, lr, lc, f);
  const bool has0 = lc < nk.P, has1 = 4 + lc < nk.P;
  const PrimR P0 = nkn_prim(pl, has0 ? lc : 0), P1 = nkn_prim(pl, has1 ? 4 + lc : 0);
  // constant-1 columns of the staged layer inputs (bias gradients)
  const double one1a = lc == nk.P ? 1.0 : 0.0, one1b = 4 + lc == nk.P ? 1.0 : 0.0;
  const double one2a = lc == nk.i2 ? 1.0 : 0.0, one2b = 4 + lc == nk.i2 ? 1.0 : 0.0;

  double acc0[1 + NKN_MAXD], acc1[1 + NKN_MAXD];
#pragma unroll
  for (int k = 0; k <= NKN_MAXD; ++k) acc0[k] = acc1[k] = 0.0;
  double gw1[2] = {0.0, 0.0}, gw2[2] = {0.0, 0.0};   // DMMA tiles dW1[o = lr][i = 2 lc + q], dW2 likewise
  double gw3 = 0.0, gb3 = 0.0, gtr = 0.0;

  for (int64_t jt = blockIdx.x; jt < jtiles; jt += w.njc) {
    const int64_t j0 = jt * TILE;
    if (w.sym_lower && j0 > i0 + TILE - 1) break;
    __syncthreads();   // previous tile fully consumed
    for (int idx = tid; idx < TILE * pl.FT; idx += GRAM_THREADS) {
      int r = idx / pl.FT, c = idx - r * pl.FT;
      sr[r * pl.S + c] = (j0 + r < M) ? FR[(j0 + r) * pl.FT + c] : 0.0;
    }
    if (w.mode == W_GPR)
      for (int idx = tid; idx < w.R * TILE; idx += GRAM_THREADS) {
        int r = idx / TILE, c = idx - r * TILE;
        bj[idx] = (j0 + c < M) ? w.beta[(int64_t)r * N + j0 + c] : 0.0;
      }
    __syncthreads();
    // the weight of octet q + 1 is requested while octet q is computed (the first version loaded it
    // where it was used: 8 % of all stall samples sat on the instruction after that load)
    auto w_load = [&](int q) -> double {
      const int64_t gi = i0 + warp * 8 + (q >> 3), gj = j0 + (q & 7) * 8 + lr;
      if (gi >= N || gj >= M || (w.sym_lower && gj > gi)) return 0.0;
      return w.W[gi * w.ldw + gj];
    };
    double w_next = w_load(0);
    for (int q = 0; q < TILE; ++q) {
      const double w_raw = w_next;
      w_next = q + 1 < TILE ? w_load(q + 1) : 0.0;
      const int il = warp * 8 + (q >> 3), jb = (q & 7) * 8, jl = jb + lr;
      const int64_t gi = i0 + il, gj = j0 + jl;
      if (gi >= N || j0 + jb >= M) continue;                      // warp uniform
      if (w.sym_lower && j0 + jb > gi) continue;
      const bool live = gj < M && !(w.sym_lower && gj > gi);
      double wij = 0.0;
      if (live) {
        if (w.mode == W_GPR) {
          double bb = 0.0;
          for (int r = 0; r < w.R; ++r) bb = fma(bi[r * TILE + il], bj[r * TILE + jl], bb);
          wij = 0.5 * ((double)w.R * w_raw - bb);
          if (gi == gj && lc == 0) gtr += wij;                    // tr W = d nlml / d noise
        } else {
          wij = w_raw;
        }
        if (w.sym_lower && gj != gi) wij *= 2.0;
      }
      const double* fi = sl + il * pl.S;
      const double* fj = sr + jl * pl.S;
      PrimEval e0, e1;
      e0.k = e0.dk = e0.d2 = 0.0;
      e1 = e0;
      if (has0) e0 = nkn_prim_eval(P0, th, tinv, fi + P0.feat_off, fj + P0.feat_off);
      if (has1) e1 = nkn_prim_eval(P1, th, tinv, fi + P1.feat_off, fj + P1.feat_off);
      double o1[2], o2[2], h1, h2;
      nkn_forward(f, e0.k, e1.k, o1, o2, h1, h2);

      // ---- Linear 3 (-> 1) and Product 2
      gw3 += wij * h2;
      if (lc == 0) gb3 += wij;
      const double dh2 = wij * f.w3;
      const double d2a = dh2 * o2[1], d2b = dh2 * o2[0];          // adjoints of Linear-2 outputs 2 lc, 2 lc + 1
      // ---- Linear 2: weight-gradient tile and adjoint of its inputs
      Ts[lr * 8 + 2 * lc] = d2a;
      Ts[lr * 8 + 2 * lc + 1] = d2b;
      Hs[lr * 8 + lc] = lc < nk.i2 ? h1 : one2a;
      Hs[lr * 8 + 4 + lc] = one2b;
      __syncwarp();
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        dmma884(gw2[0], gw2[1], Ts[(4 * s + lc) * 8 + lr], Hs[(4 * s + lc) * 8 + lr]);
        dmma884(q0, q1, Ts[lr * 8 + 4 * s + lc], f.w2t[s]);       // dh1[e][2 lc], [2 lc + 1]
      }
      __syncwarp();
      const int src = qb + (lc >> 1);
      double t0 = __shfl_sync(0xffffffffu, q0, src), t1 = __shfl_sync(0xffffffffu, q1, src);
      const double dh1 = (lc & 1) ? t1 : t0;                      // adjoint of Product-1 output lc
      const double d1a = dh1 * o1[1], d1b = dh1 * o1[0];          // adjoints of Linear-1 outputs 2 lc, 2 lc + 1
      // ---- Linear 1
      Ts[lr * 8 + 2 * lc] = d1a;
      Ts[lr * 8 + 2 * lc + 1] = d1b;
      Hs[lr * 8 + lc] = has0 ? e0.k : one1a;
      Hs[lr * 8 + 4 + lc] = has1 ? e1.k : one1b;
      __syncwarp();
      double r0 = 0.0, r1 = 0.0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        dmma884(gw1[0], gw1[1], Ts[(4 * s + lc) * 8 + lr], Hs[(4 * s + lc) * 8 + lr]);
        dmma884(r0, r1, Ts[lr * 8 + 4 * s + lc], f.w1t[s]);       // dk[e][2 lc], [2 lc + 1]
      }
      __syncwarp();
      t0 = __shfl_sync(0xffffffffu, r0, src);
      t1 = __shfl_sync(0xffffffffu, r1, src);
      const double g0 = (lc & 1) ? t1 : t0;                       // adjoint of primitive lc
      t0 = __shfl_sync(0xffffffffu, r0, src + 2);
      t1 = __shfl_sync(0xffffffffu, r1, src + 2);
      const double g1 = (lc & 1) ? t1 : t0;                       // adjoint of primitive 4 + lc
      if (has0) nkn_prim_grad(P0, fi + P0.feat_off, fj + P0.feat_off, e0, g0, acc0);
      if (has1) nkn_prim_grad(P1, fi + P1.feat_off, fj + P1.feat_off, e1, g1, acc1);
    }
  }
  if (has0) nkn_prim_scale(P0, tinv, acc0);
  if (has1) nkn_prim_scale(P1, tinv, acc1);
  // ---- CTA reduction (fixed order): per-warp rows of `red` (on top of the feature tiles), then over the 8 warps
  __syncthreads();
  for (int t = tid; t < 8 * nacc; t += GRAM_THREADS) red[t] = 0.0;
  __syncthreads();
  double* rw = red + warp * nacc;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = 2 * lc + q;
    if (lr < nk.n1) {
      if (i < nk.P) rw[nk.w1 + lr * nk.P + i] = gw1[q];
      else if (i == nk.P) rw[nk.b1 + lr] = gw1[q];
    }
    if (lr < nk.n2) {
      if (i < nk.i2) rw[nk.w2 + lr * nk.i2 + i] = gw2[q];
      else if (i == nk.i2) rw[nk.b2 + lr] = gw2[q];
    }
  }
  // sums over the 8 elements (lr) of what each lane column accumulated
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    gw3 += __shfl_xor_sync(0xffffffffu, gw3, o);
    gb3 += __shfl_xor_sync(0xffffffffu, gb3, o);
    gtr += __shfl_xor_sync(0xffffffffu, gtr, o);
#pragma unroll
    for (int k = 0; k <= NKN_MAXD; ++k) {
      acc0[k] += __shfl_xor_sync(0xffffffffu, acc0[k], o);
      acc1[k] += __shfl_xor_sync(0xffffffffu, acc1[k], o);
    }
  }
  if (lr == 0) {
    if (lc < nk.i3) rw[nk.w3 + lc] = gw3;
    if (lc == 0) {
      rw[nk.b3] = gb3;
      rw[pl.n_theta] = gtr;
    }
    const int np0 = has0 ? nkn_prim_nparams(P0) : 0, np1 = has1 ? nkn_prim_nparams(P1) : 0;
#pragma unroll
    for (int k = 0; k <= NKN_MAXD; ++k) {
      if (k < np0) rw[P0.theta_off + k] = acc0[k];
      if (k < np1) rw[P1.theta_off + k] = acc1[k];
    }
  }
  __syncthreads();
  const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int t = tid; t < nacc; t += GRAM_THREADS) {
    double s = 0.0;
    for (int wq = 0; wq < GRAM_THREADS / 32; ++wq) s += red[wq * nacc + t];
    part_theta[cta * nacc + t] = s;
  }
}

// --------------------------------------------------------------------------- Kdiag
__global__ void kdiag_kernel(const Plan pl, const PlanDims pd, const double* __restrict__ theta,
                             const double* __restrict__ X, int64_t ldx, int64_t N,
                             double* __restrict__ out, const double* __restrict__ wvec,
                             double* __restrict__ part_theta, double* __restrict__ dX, int64_t lddx,
                             int xcols) {
  // forward (wvec == nullptr): out[i] = Kdiag(x_i).  backward: accumulates sum_i w_i dKdiag_i/dtheta
  // into per-thread partials and writes dX rows.
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool bwd = wvec != nullptr;
  double acc[MAX_ACC];
  const int nacc = pl.n_theta;
  if (bwd)
    for (int t = 0; t < nacc; ++t) acc[t] = 0.0;
  if (i < N) {
    const double* x = X + i * ldx;
    double v[GPS_MAX_SLOTS], vb[GPS_MAX_SLOTS];
    for (int p = 0; p < pl.n_prims; ++p) {
      const PrimC P = pl.prims[p];
      const double* t = theta + P.theta_off;
      if (P.type == GPS_LINEAR) {                    // kernels.py:507-510
        double s = 0.0;
        for (int k = 0; k < P.ndims; ++k) {
          double xv = x[pd.dims[p][k]];
          s += xv * xv * t[P.ard ? k : 0];
        }
        v[p] = s;
      } else {
        v[p] = t[0];                                 // kernels.py:428-429, :803-804
      }
    }
    program_fwd(pl, theta, v);
    if (!bwd) {
      out[i] = v[pl.out_slot];
    } else {
      int nslots_used = pl.n_prims;
      if (pl.n_ops) {
        const OpC last = pl.ops[pl.n_ops - 1];
        nslots_used = last.dst + ((last.op == GPS_OP_LINEAR || last.op == GPS_OP_PRODUCT) ? last.n : 1);
      }
      for (int s = 0; s < nslots_used; ++s) vb[s] = 0.0;
      vb[pl.out_slot] = wvec[i];
      program_bwd(pl, theta, v, vb, acc);
      if (dX)
        for (int d = 0; d < xcols; ++d) dX[i * lddx + d] = 0.0;
      for (int p = 0; p < pl.n_prims; ++p) {
        const PrimC P = pl.prims[p];
        const double* t = theta + P.theta_off;
        if (P.type == GPS_LINEAR) {
          for (int k = 0; k < P.ndims; ++k) {
            double xv = x[pd.dims[p][k]];
            acc[P.theta_off + (P.ard ? k : 0)] += vb[p] * xv * xv;
            if (dX) dX[i * lddx + pd.dims[p][k]] += vb[p] * 2.0 * xv * t[P.ard ? k : 0];
          }
        } else {
          acc[P.theta_off] += vb[p];
        }
      }
    }
  }
  if (bwd) {
    __shared__ double red[GRAM_THREADS / 32][MAX_ACC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = 0; t < nacc; ++t) {
      double s = warp_sum(acc[t]);
      if (lane == 0) red[warp][t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nacc; t += blockDim.x) {
      double s = 0.0;
      for (int wq = 0; wq < GRAM_THREADS / 32; ++wq) s += red[wq][t];
      part_theta[(int64_t)blockIdx.x * nacc + t] = s;
    }
  }
}

int features(gps_handle* h, const Plan& pl, const PlanDims& pd, const double* theta, Mat X, int slot,
             double** out) {
  double* F = (double*)gps_ws(h, slot, (size_t)X.rows * pl.FT * sizeof(double));
  if (!F) return -102;
  if (X.rows > 0) {
    feature_kernel<<<(unsigned)((X.rows + 127) / 128), 128, 0, h->stream>>>(pl, pd, theta, X.p, X.ld,
                                                                           X.rows, F);
    GPS_LAUNCH_CHECK(h);
  }
  *out = F;
  return 0;
}

bool g_gram_attr = false;
void gram_attrs() {
  if (g_gram_attr) return;
  cudaFuncSetAttribute(gram_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(gram_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(gram_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gram_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gram_fwd_nkn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(gram_bwd_nkn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  g_gram_attr = true;
}

}  // namespace

int gps_gram_fwd_mat(gps_handle* h, const gps_kernel_desc* desc, const double* theta, Mat X,
                     const Mat* X2, double diag_add, int uplo, Mat K) {
  Plan pl;
  PlanDims pd;
  int rc;
  if ((rc = build_plan(h, desc, X.cols, &pl, &pd))) return rc;
  if (X2 && X2->cols != X.cols) return gps_fail(h, -5, "gram: X2 has %lld columns, X has %lld", (long long)X2->cols, (long long)X.cols);
  const int64_t N = X.rows, M = X2 ? X2->rows : X.rows;
  if (K.rows != N || K.cols != M) return gps_fail(h, -8, "gram: K_out must be %lld x %lld", (long long)N, (long long)M);
  if (N == 0 || M == 0) return 0;
  double *FL, *FR;
  if ((rc = features(h, pl, pd, theta, X, WS_FEAT_L, &FL))) return rc;
  FR = FL;
  if (X2 && (rc = features(h, pl, pd, theta, *X2, WS_FEAT_R, &FR))) return rc;
  gram_attrs();
  size_t smem = (size_t)(pl.n_theta + 2 * TILE * pl.S) * sizeof(double);
  dim3 grid((unsigned)((M + TILE - 1) / TILE), (unsigned)((N + TILE - 1) / TILE));
  if (grid.y > 65535) return gps_fail(h, -4, "gram: too many rows");
  if (stat_fast(h, pl)) {
    const PrimC P = pl.prims[0];
    const int sym = X2 ? 0 : 1, up = X2 ? 0 : uplo;
    if (P.ndims <= 4)
      gram_fwd_stat_kernel<4><<<grid, GRAM_THREADS, 0, h->stream>>>(P.type, P.ndims, pl.FT, theta, FL, FR, N, M, diag_add, sym, up, K.p, K.ld);
    else if (P.ndims <= 8)
      gram_fwd_stat_kernel<8><<<grid, GRAM_THREADS, 0, h->stream>>>(P.type, P.ndims, pl.FT, theta, FL, FR, N, M, diag_add, sym, up, K.p, K.ld);
    else
      gram_fwd_stat_kernel<16><<<grid, GRAM_THREADS, 0, h->stream>>>(P.type, P.ndims, pl.FT, theta, FL, FR, N, M, diag_add, sym, up, K.p, K.ld);
    GPS_LAUNCH_CHECK(h);
    return 0;
  }
  NknPlan nk;
  if (nkn_match(h, pl, &nk)) {
    // Linear / Product(2) networks: the layers on the FP64 tensor cores (see gram_bwd_nkn_kernel)
    Plan pn = pl;
    pn.S = nkn_row_stride(pl.FT);
    size_t smem_n = (size_t)(2 * pl.n_theta + 2 * TILE * pn.S) * sizeof(double);
    if (smem_n > 200 * 1024) { pn.S = pl.S; smem_n = smem + (size_t)pl.n_theta * sizeof(double); }
    gram_fwd_nkn_kernel<<<grid, GRAM_THREADS, smem_n, h->stream>>>(pn, nk, theta, FL, FR, N, M, diag_add,
                                                                   X2 ? 0 : 1, X2 ? 0 : uplo, K.p, K.ld);
    GPS_LAUNCH_CHECK(h);
    return 0;
  }
  const int nslots = plan_nslots(pl);
  const size_t smem2 = smem + (size_t)nslots * GRAM_THREADS * sizeof(double);
  if (h->gram_impl == 2 && smem2 <= 227 * 1024) {
    // experimental: slot values in shared memory instead of local memory (see gram_bwd_smem_kernel)
    gram_fwd_smem_kernel<<<grid, GRAM_THREADS, smem2, h->stream>>>(pl, nslots, theta, FL, FR, N, M, diag_add,
                                                                   X2 ? 0 : 1, X2 ? 0 : uplo, K.p, K.ld);
    GPS_LAUNCH_CHECK(h);
    return 0;
  }
  gram_fwd_kernel<<<grid, GRAM_THREADS, smem, h->stream>>>(pl, theta, FL, FR, N, M, diag_add,
                                                           X2 ? 0 : 1, X2 ? 0 : uplo, K.p, K.ld);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

int gps_gram_bwd_mat(gps_handle* h, const gps_kernel_desc* desc, const double* theta, Mat X,
                     const Mat* X2, GramW w, double* dtheta_out, Mat* dX, double* trace_out) {
  Plan pl;
  PlanDims pd;
  int rc;
  if ((rc = build_plan(h, desc, X.cols, &pl, &pd))) return rc;
  const int64_t N = X.rows, M = X2 ? X2->rows : X.rows;
  if (X2 && X2->cols != X.cols) return gps_fail(h, -5, "gram_bwd: X2 column mismatch");
  if (w.W.rows != N || w.W.cols != M) return gps_fail(h, -6, "gram_bwd: W must be %lld x %lld", (long long)N, (long long)M);
  if (w.mode == W_GPR && (w.R < 1 || w.R > MAX_R)) return gps_fail(h, -6, "gram_bwd: R out of range (max %d)", MAX_R);
  if (dX && w.sym_lower) return gps_fail(h, -8, "gram_bwd: dX needs the full weight matrix");
  if (dX && (dX->rows != N || dX->cols != X.cols)) return gps_fail(h, -8, "gram_bwd: dX shape mismatch");
  if (X.cols > GPS_MAX_DIMS && dX) return gps_fail(h, -4, "gram_bwd: dX supports at most %d columns", GPS_MAX_DIMS);
  const int nacc = pl.n_theta + 1;
  if (N == 0 || M == 0) {
    GPS_CUDA(h, cudaMemsetAsync(dtheta_out, 0, pl.n_theta * sizeof(double), h->stream));
    if (trace_out) GPS_CUDA(h, cudaMemsetAsync(trace_out, 0, sizeof(double), h->stream));
    return 0;
  }
  double *FL, *FR;
  if ((rc = features(h, pl, pd, theta, X, WS_FEAT_L, &FL))) return rc;
  FR = FL;
  if (X2 && (rc = features(h, pl, pd, theta, *X2, WS_FEAT_R, &FR))) return rc;
  gram_attrs();
  // experimental shared-memory-accumulator interpreter (gram_impl == 2), when its arrays fit
  const int nslots = plan_nslots(pl);
  const size_t smem2 = bwd_smem_doubles(pl, nslots, w.mode, w.R, dX ? 1 : 0, (int)X.cols) * sizeof(double);
  const bool use_smem_acc = h->gram_impl == 2 && smem2 <= 227 * 1024;
  const int tile = use_smem_acc ? T2 : TILE;
  const int64_t itiles = (N + tile - 1) / tile, jtiles = (M + tile - 1) / tile;
  int64_t njc = (4 * h->sm_count + itiles - 1) / itiles;
  if (njc < 1) njc = 1;
  if (njc > jtiles) njc = jtiles;
  if (itiles > 65535) return gps_fail(h, -4, "gram_bwd: too many rows");
  BwdArgs a;
  a.mode = w.mode; a.W = w.W.p; a.ldw = w.W.ld; a.beta = w.beta; a.R = w.R;
  a.sym_lower = w.sym_lower; a.want_dx = dX ? 1 : 0; a.xcols = (int)X.cols;
  a.dx_scale = X2 ? 1.0 : 2.0; a.njc = (int)njc;
  NknPlan nk;
  Plan pn = pl;
  pn.S = nkn_row_stride(pl.FT);
  const size_t nkn_fixed = (size_t)(2 * pl.n_theta + (w.mode == W_GPR ? 2 * w.R * TILE : 0) + 8 * 128);
  auto nkn_bytes = [&](int S) { return (nkn_fixed + std::max<size_t>(2 * TILE * S, 8 * nacc)) * sizeof(double); };
  size_t smem_nkn = nkn_bytes(pn.S);
  if (smem_nkn > 113 * 1024 && nkn_bytes(pl.S) <= 113 * 1024) {   // two CTAs per SM (228 KB, 1 KB reserved each) first
    pn.S = pl.S;
    smem_nkn = nkn_bytes(pn.S);
  }
  const bool nkn = !use_smem_acc && !dX && smem_nkn <= 220 * 1024 && nkn_match(h, pl, &nk);
  if (nkn) {
    // two CTAs per SM and uneven work per CTA (lower tiles only): many more CTAs than slots
    njc = (16 * h->sm_count + itiles - 1) / itiles;
    if (njc > jtiles) njc = jtiles;
    a.njc = (int)njc;
  }
  const int64_t nctas = itiles * njc;
  double* part = (double*)gps_ws(h, WS_PARTIAL, (size_t)nctas * nacc * sizeof(double));
  if (!part) return -102;
  // register-tiled fast path (with the input gradient where its product form is safe)
  const bool fast = !use_smem_acc && stat_fast(h, pl) && (!dX || stat_dx_ok(pl.prims[0].type));
  double* pdx = nullptr;
  if (dX && !fast) {
    pdx = (double*)gps_ws(h, WS_PARTIAL2, (size_t)njc * N * X.cols * sizeof(double));
    if (!pdx) return -102;
  }
  size_t smem = (size_t)(2 * pl.n_theta + 2 * TILE * pl.S + (w.mode == W_GPR ? 2 * w.R * TILE : 0) +
                         8 * nacc + (dX ? TILE * X.cols : 0)) * sizeof(double);
  if (!use_smem_acc && smem > 220 * 1024) return gps_fail(h, -2, "gram_bwd: kernel too large for shared memory");
  if (use_smem_acc) {
    gram_bwd_smem_kernel<<<dim3((unsigned)njc, (unsigned)itiles), NT2, smem2, h->stream>>>(
        pl, pd, nslots, theta, FL, FR, N, M, a, part, pdx);
  } else if (nkn) {
    gram_bwd_nkn_kernel<<<dim3((unsigned)njc, (unsigned)itiles), GRAM_THREADS, smem_nkn, h->stream>>>(
        pn, nk, theta, FL, FR, N, M, a, part);
  } else if (fast) {
    const PrimC P = pl.prims[0];
    const dim3 g2((unsigned)njc, (unsigned)itiles);
    double* G = nullptr;
    const int64_t ldg = (M + 15) / 16 * 16;
    if (dX) {
      G = (double*)gps_ws(h, WS_GRAM_G, (size_t)N * ldg * sizeof(double));
      if (!G) return -102;
    }
    if (P.ndims <= 4)
      gram_bwd_stat_kernel<4><<<g2, GRAM_THREADS, 0, h->stream>>>(P.type, P.ard, P.ndims, pl.FT, pl.n_theta, theta, FL, FR, N, M, a, part, G, ldg);
    else if (P.ndims <= 8)
      gram_bwd_stat_kernel<8><<<g2, GRAM_THREADS, 0, h->stream>>>(P.type, P.ard, P.ndims, pl.FT, pl.n_theta, theta, FL, FR, N, M, a, part, G, ldg);
    else
      gram_bwd_stat_kernel<16><<<g2, GRAM_THREADS, 0, h->stream>>>(P.type, P.ard, P.ndims, pl.FT, pl.n_theta, theta, FL, FR, N, M, a, part, G, ldg);
    if (dX) {
      GPS_LAUNCH_CHECK(h);
      const int nd = P.ndims;
      const int64_t ldpm = (nd + 1 + 15) / 16 * 16;
      double* B = (double*)gps_ws(h, WS_GRAM_B, (size_t)(nd + 1) * ldg * sizeof(double));
      double* Pm = (double*)gps_ws(h, WS_GRAM_P, (size_t)N * ldpm * sizeof(double));
      if (!B || !Pm) return -102;
      stat_dx_operand_kernel<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(FR, pl.FT, nd, M, B, ldg);
      GPS_LAUNCH_CHECK(h);
      if ((rc = gps_gemm_nt_launch(h, 1.0, Mat(G, N, M, ldg), Mat(B, nd + 1, M, ldg), 0.0, Mat(Pm, N, nd + 1, ldpm),
                                   TRI_NONE, TRI_NONE, C_ALL)))
        return rc;
      const int64_t tot = N * X.cols;
      stat_dx_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(
          Pm, ldpm, FL, pl.FT, nd, P.ard, theta, pd, N, (int)X.cols, a.dx_scale, dX->p, dX->ld);
    }
  } else {
    gram_bwd_kernel<<<dim3((unsigned)njc, (unsigned)itiles), GRAM_THREADS, smem, h->stream>>>(
        pl, pd, theta, FL, FR, N, M, a, part, pdx);
  }
  GPS_LAUNCH_CHECK(h);
  // theta gradient
  reduce_cols_kernel<<<(nacc + 127) / 128, 128, 0, h->stream>>>(part, nctas, nacc, 1.0, part, 0, 0);
  GPS_LAUNCH_CHECK(h);
  // (the reduction wrote row 0 of `part` in place: column t only reads rows of column t)
  GPS_CUDA(h, cudaMemcpyAsync(dtheta_out, part, pl.n_theta * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (trace_out)
    GPS_CUDA(h, cudaMemcpyAsync(trace_out, part + pl.n_theta, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (dX && !fast) {
    int64_t tot = N * X.cols;
    reduce_dx_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(pdx, (int)njc, N, (int)X.cols,
                                                                          a.dx_scale, dX->p, dX->ld);
    GPS_LAUNCH_CHECK(h);
  }
  return 0;
}

int gps_kdiag_fwd_vec(gps_handle* h, const gps_kernel_desc* desc, const double* theta, Mat X,
                      double* out) {
  Plan pl;
  PlanDims pd;
  int rc;
  if ((rc = build_plan(h, desc, X.cols, &pl, &pd))) return rc;
  if (X.rows == 0) return 0;
  kdiag_kernel<<<(unsigned)((X.rows + GRAM_THREADS - 1) / GRAM_THREADS), GRAM_THREADS, 0, h->stream>>>(
      pl, pd, theta, X.p, X.ld, X.rows, out, nullptr, nullptr, nullptr, 0, (int)X.cols);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

// ------------------------------------------------------------------------------- C ABI
namespace {
int theta_ptr(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta, int argidx,
              const double** out) {
  Mat t;
  int rc;
  if ((rc = gps_as_mat(h, theta, argidx, "theta", &t))) return rc;
  if (!desc) return gps_fail(h, -2, "null kernel descriptor");
  if (t.rows * t.cols < desc->n_theta)
    return gps_fail(h, -argidx, "theta has %lld entries, descriptor needs %d", (long long)(t.rows * t.cols), desc->n_theta);
  *out = t.p;
  return 0;
}
}  // namespace

extern "C" {

int gps_gram_fwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta, const DLTensor* X,
                 const DLTensor* X2, double diag_add, int uplo, DLTensor* K_out) {
  if (!h) return -1;
  const double* th;
  Mat x, x2, k;
  int rc;
  if ((rc = theta_ptr(h, desc, theta, 3, &th))) return rc;
  if ((rc = gps_as_mat(h, X, 4, "X", &x, false))) return rc;
  if (X2 && (rc = gps_as_mat(h, X2, 5, "X2", &x2, false))) return rc;
  if ((rc = gps_as_mat(h, K_out, 8, "K_out", &k, false))) return rc;
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_gram_fwd_mat(h, desc, th, x, X2 ? &x2 : nullptr, diag_add, uplo, k);
}

int gps_gram_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta, const DLTensor* X,
                 const DLTensor* X2, const DLTensor* W, DLTensor* dtheta_out, DLTensor* dX_out) {
  if (!h) return -1;
  const double* th;
  Mat x, x2, w, dth, dx;
  int rc;
  if ((rc = theta_ptr(h, desc, theta, 3, &th))) return rc;
  if ((rc = gps_as_mat(h, X, 4, "X", &x, false))) return rc;
  if (X2 && (rc = gps_as_mat(h, X2, 5, "X2", &x2, false))) return rc;
  if ((rc = gps_as_mat(h, W, 6, "W", &w, false))) return rc;
  if ((rc = gps_as_mat(h, dtheta_out, 7, "dtheta_out", &dth))) return rc;
  if (dth.rows * dth.cols < desc->n_theta) return gps_fail(h, -7, "dtheta_out too small");
  if (dX_out && (rc = gps_as_mat(h, dX_out, 8, "dX_out", &dx, false))) return rc;
  GPS_CUDA(h, cudaSetDevice(h->device));
  GramW gw;
  gw.mode = W_DENSE; gw.W = w; gw.beta = nullptr; gw.R = 0; gw.sym_lower = 0;
  return gps_gram_bwd_mat(h, desc, th, x, X2 ? &x2 : nullptr, gw, dth.p, dX_out ? &dx : nullptr, nullptr);
}

int gps_kdiag_fwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta, const DLTensor* X,
                  DLTensor* out) {
  if (!h) return -1;
  const double* th;
  Mat x, o;
  int rc;
  if ((rc = theta_ptr(h, desc, theta, 3, &th))) return rc;
  if ((rc = gps_as_mat(h, X, 4, "X", &x, false))) return rc;
  if ((rc = gps_as_mat(h, out, 5, "out", &o))) return rc;
  if (o.rows * o.cols != x.rows) return gps_fail(h, -5, "kdiag: out must have %lld entries", (long long)x.rows);
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_kdiag_fwd_vec(h, desc, th, x, o.p);
}

int gps_kdiag_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta, const DLTensor* X,
                  const DLTensor* wv, DLTensor* dtheta_out, DLTensor* dX_out) {
  if (!h) return -1;
  const double* th;
  Mat x, w, dth, dx;
  int rc;
  if ((rc = theta_ptr(h, desc, theta, 3, &th))) return rc;
  if ((rc = gps_as_mat(h, X, 4, "X", &x, false))) return rc;
  if ((rc = gps_as_mat(h, wv, 5, "w", &w))) return rc;
  if (w.rows * w.cols != x.rows) return gps_fail(h, -5, "kdiag_bwd: w must have %lld entries", (long long)x.rows);
  if ((rc = gps_as_mat(h, dtheta_out, 6, "dtheta_out", &dth))) return rc;
  if (dth.rows * dth.cols < desc->n_theta) return gps_fail(h, -6, "dtheta_out too small");
  if (dX_out) {
    if ((rc = gps_as_mat(h, dX_out, 7, "dX_out", &dx, false))) return rc;
    if (dx.rows != x.rows || dx.cols != x.cols) return gps_fail(h, -7, "kdiag_bwd: dX shape mismatch");
  }
  GPS_CUDA(h, cudaSetDevice(h->device));
  Plan pl;
  PlanDims pd;
  if ((rc = build_plan(h, desc, x.cols, &pl, &pd))) return rc;
  if (x.rows == 0) {
    GPS_CUDA(h, cudaMemsetAsync(dth.p, 0, pl.n_theta * sizeof(double), h->stream));
    return 0;
  }
  unsigned nblk = (unsigned)((x.rows + GRAM_THREADS - 1) / GRAM_THREADS);
  double* part = (double*)gps_ws(h, WS_PARTIAL, (size_t)nblk * pl.n_theta * sizeof(double));
  if (!part) return -102;
  kdiag_kernel<<<nblk, GRAM_THREADS, 0, h->stream>>>(pl, pd, th, x.p, x.ld, x.rows, nullptr, w.p, part,
                                                     dX_out ? dx.p : nullptr, dX_out ? dx.ld : 0, (int)x.cols);
  GPS_LAUNCH_CHECK(h);
  reduce_cols_kernel<<<(pl.n_theta + 127) / 128, 128, 0, h->stream>>>(part, nblk, pl.n_theta, 1.0, dth.p, 0, 0);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

}  // extern "C"
