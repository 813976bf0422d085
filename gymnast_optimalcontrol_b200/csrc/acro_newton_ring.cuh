// acro_newton_ring.cuh - the Newton / Armijo loop (tg:298-398) for small and medium batches, where every
// problem advances at the speed of its own dependency chain and memory latency must never be exposed.
//
// One warp = one tile of 32 problems, warp-synchronous control flow (votes instead of per-thread loops), and
// all streamed operands arrive through a per-warp ring of ACRO_RING_D shared-memory stages filled by 1-D bulk
// TMA copies (cp.async.bulk ... mbarrier::complete_tx).  In the tiled layout the rows a warp needs for one
// time step of one array are one contiguous block, so a stage is 4-6 bulk copies issued by lane 0:
//     forward pass  : X[t] (1 KB) | U[t] (512 B) | K[t] (2 KB) | S[t] (512 B) | reference x,u
//     backward pass : X[t] (1 KB) | U[t] (512 B) | lin[t] (2.5 KB)            | reference x,u
// The consumer reads its own lane's column with LDS at constant offsets.  Stage t is refilled (with step
// t + D) one iteration after it was read, when the reads have provably completed, so no "empty" barriers are
// needed.  Results are written with plain predicated stores: a lane whose problem has finished, or whose
// Armijo candidate was already accepted, keeps computing in lockstep but never stores.
#pragma once
#include <cstdint>

#include "acro_device.cuh"
#include "acro_views.cuh"

namespace acro {

#define ACRO_RING_D 8  // power of two

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One lane of the (fully active) warp, chosen by the hardware.  ptxas knows that exactly one lane follows the
// branch on this predicate, so the bulk copies inside keep their uniform-register operands (a branch on
// `lane == 0` makes it wrap every copy in an ELECT / R2UR.BROADCAST loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ double lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Stage layout (byte offsets).  Per-problem references add six tiled rows, shared ones 48 bytes.
constexpr uint32_t kOffX = 0, kOffU = 1024, kOffA = 1536 /* K+S or lin: 2560 B */, kOffRef = 4096;
template <bool RPB>
__host__ __device__ constexpr uint32_t stage_bytes() { return RPB ? 4096 + 1536 : 4096 + 128; }
template <bool RPB>
__host__ __device__ constexpr uint32_t tx_bytes() { return RPB ? 4096 + 1536 : 4096 + 48; }

struct Ring {
  uint32_t data, bars;  // shared addresses of stage 0 / barrier 0
  uint32_t seq;         // fills consumed so far (warp-uniform); fills issued = seq + in-flight
  uint32_t iss;         // fills issued so far
};

// Everything a pass needs to find one warp's operands: tile base pointers (at t = 0) and per-step strides.
struct TilePtrs {
  const double *x, *u, *k, *s, *lin;  // source iterate (current), gains, feed-forward, linearisation
  const double *rx, *ru;              // reference (shared: plain arrays; per problem: tile bases)
  int64_t sx, su, sk, ss, sl;         // doubles per time step of the tiled arrays (Bp * C)
};

// Source pointers of the next fill; they walk forwards (forward pass) or backwards (backward pass) one time
// step per fill, so issuing a stage needs no multiplications.
struct FillPtrs {
  const double *x, *u, *a0, *a1, *rx, *ru;
};
template <bool RPB, bool FWD>
__device__ __forceinline__ FillPtrs fill_begin(const TilePtrs& p, int steps) {
  const int64_t t0 = FWD ? 0 : steps - 1;
  FillPtrs f;
  f.x = p.x + t0 * p.sx;
  f.u = p.u + t0 * p.su;
  f.a0 = FWD ? p.k + t0 * p.sk : p.lin + t0 * p.sl;
  f.a1 = p.s + t0 * p.ss;
  f.rx = p.rx + t0 * (RPB ? p.sx : 4);
  f.ru = p.ru + t0 * (RPB ? p.su : 2);
  return f;
}
template <bool RPB, bool FWD>
__device__ __forceinline__ void ring_fill(Ring& r, const TilePtrs& p, FillPtrs& f, int lane) {
  if (elect_one()) {
    const uint32_t st = r.iss & (ACRO_RING_D - 1);
    const uint32_t bar = r.bars + st * 8, dst = r.data + st * stage_bytes<RPB>();
    mbar_expect_tx(bar, tx_bytes<RPB>());
    bulk_g2s(dst + kOffX, f.x, 1024, bar);
    bulk_g2s(dst + kOffU, f.u, 512, bar);
    if (FWD) {
      bulk_g2s(dst + kOffA, f.a0, 2048, bar);
      bulk_g2s(dst + kOffA + 2048, f.a1, 512, bar);
    } else {
      bulk_g2s(dst + kOffA, f.a0, 2560, bar);
    }
    bulk_g2s(dst + kOffRef, f.rx, RPB ? 1024 : 32, bar);
    bulk_g2s(dst + kOffRef + (RPB ? 1024 : 32), f.ru, RPB ? 512 : 16, bar);
  }
  constexpr int dir = FWD ? 1 : -1;
  f.x += dir * p.sx;
  f.u += dir * p.su;
  f.a0 += dir * (FWD ? p.sk : p.sl);
  f.a1 += dir * p.ss;
  f.rx += dir * (RPB ? p.sx : 4);
  f.ru += dir * (RPB ? p.su : 2);
  ++r.iss;
}

struct RingStage {
  uint32_t base;  // shared address of the stage + lane * 8
  uint32_t ref;   // shared address of the reference block (+ lane * 8 when per problem)
};
template <bool RPB>
__device__ __forceinline__ RingStage ring_stage(const Ring& r, uint32_t n, int lane) {
  const uint32_t st = n & (ACRO_RING_D - 1);
  const uint32_t b = r.data + st * stage_bytes<RPB>();
  return RingStage{b + lane * 8u, b + kOffRef + (RPB ? lane * 8u : 0u)};
}
__device__ __forceinline__ uint32_t ring_bar(const Ring& r, uint32_t n) { return r.bars + (n & (ACRO_RING_D - 1)) * 8; }
__device__ __forceinline__ uint32_t ring_parity(uint32_t n) { return (n / ACRO_RING_D) & 1u; }

template <bool RPB>
__device__ __forceinline__ void lds_ref(const RingStage& s, double xr[4], double ur[2]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) xr[c] = lds(s.ref + (RPB ? c * 256 : c * 8));
#pragma unroll
  for (int c = 0; c < 2; ++c) ur[c] = lds(s.ref + (RPB ? 1024 + c * 256 : 32 + c * 8));
}

// ---------------------------------------------------------------------------------------------------------
// forward pass (tg:218-252): closed-loop rollout with step size gamma + its cost; writes the candidate
// (Xo, Uo) and the linearisation about it when `store`.
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB>
__device__ __forceinline__ double forward_ring(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                               int lane, double gamma, bool store, double* __restrict__ Xo,
                                               double* __restrict__ Uo, double* __restrict__ Lo, const double xrT[4]) {
  const int steps = N - 1;
  FillPtrs f = fill_begin<RPB, true>(p, steps);
  for (int i = 0; i < ACRO_RING_D && i < steps; ++i) ring_fill<RPB, true>(r, p, f, lane);
  double xp[4];
  StepIn in;
  {
    mbar_wait(ring_bar(r, r.seq), ring_parity(r.seq));
    const RingStage s = ring_stage<RPB>(r, r.seq, lane);
#pragma unroll
    for (int c = 0; c < 4; ++c) in.x[c] = lds(s.base + kOffX + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) in.u[c] = lds(s.base + kOffU + c * 256);
#pragma unroll
    for (int c = 0; c < 8; ++c) in.k[c] = lds(s.base + kOffA + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) in.s[c] = lds(s.base + kOffA + 2048 + c * 256);
    lds_ref<RPB>(s, in.xr, in.ur);
    ++r.seq;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) xp[c] = in.x[c];  // x+_0 = x_0
  double cost = 0.0;
  double* po_x = Xo + lane;
  double* po_u = Uo + lane;
  double* po_l = Lo + lane;
  for (int t = 0; t < steps; ++t) {
    const bool more = t + 1 < steps;
    const uint32_t ready = more ? mbar_test(ring_bar(r, r.seq), ring_parity(r.seq)) : 1u;
    double dx[4], up[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = xp[c] - in.x[c];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double kd = in.k[i * 4] * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) kd = fma(in.k[i * 4 + j], dx[j], kd);
      up[i] = (in.u[i] + kd) + gamma * in.s[i];
    }
    if (store) {
#pragma unroll
      for (int c = 0; c < 4; ++c) po_x[c * 32] = xp[c];
#pragma unroll
      for (int c = 0; c < 2; ++c) po_u[c * 32] = up[c];
    }
    double ex[4], eu[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) ex[c] = xp[c] - in.xr[c];
#pragma unroll
    for (int c = 0; c < 2; ++c) eu[c] = up[c] - in.ur[c];
    cost += quad4(ex, [&](int i, int j) { return w.Q(i, j); });
    cost += quad2(eu, [&](int i, int j) { return w.R(i, j); });
    // operands of step t+1 straight into the registers that step t no longer needs
    if (more) {
      if (!ready) mbar_wait(ring_bar(r, r.seq), ring_parity(r.seq));
      const RingStage s = ring_stage<RPB>(r, r.seq, lane);
#pragma unroll
      for (int c = 0; c < 4; ++c) in.x[c] = lds(s.base + kOffX + c * 256);
#pragma unroll
      for (int c = 0; c < 2; ++c) in.u[c] = lds(s.base + kOffU + c * 256);
#pragma unroll
      for (int c = 0; c < 8; ++c) in.k[c] = lds(s.base + kOffA + c * 256);
#pragma unroll
      for (int c = 0; c < 2; ++c) in.s[c] = lds(s.base + kOffA + 2048 + c * 256);
      lds_ref<RPB>(s, in.xr, in.ur);
      ++r.seq;
    }
    // the stage read one iteration ago is free now: refill it with step t + D
    __syncwarp();
    if (t + ACRO_RING_D < steps) ring_fill<RPB, true>(r, p, f, lane);
    double xn[4];
    LinD L;
    rk4_step_lin(m, xp, up[0], up[1], xn, L);
    if (store) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        po_l[j * 32] = L.a[0][j];
        po_l[(4 + j) * 32] = L.a[1][j];
      }
      po_l[8 * 32] = L.b[0];
      po_l[9 * 32] = L.b[1];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) xp[c] = xn[c];
    po_x += p.sx;
    po_u += p.su;
    po_l += p.sl;
  }
  double ex[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (store) po_x[c * 32] = xp[c];
    ex[c] = xp[c] - xrT[c];
  }
  cost += quad4(ex, [&](int i, int j) { return w.QT(i, j); });
  return cost;
}

// ---------------------------------------------------------------------------------------------------------
// backward pass (tg:166-216): affine Riccati sweep on the stored linearisation; writes K, S when `store`.
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB>
__device__ __forceinline__ void backward_ring(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                              int lane, bool store, double* __restrict__ K, double* __restrict__ S,
                                              const double xT[4], const double xrT[4], double& dJ_out,
                                              double& sn_out) {
  const int steps = N - 1;
  FillPtrs f = fill_begin<RPB, false>(p, steps);
  for (int i = 0; i < ACRO_RING_D && i < steps; ++i) ring_fill<RPB, false>(r, p, f, lane);
  double P[10], pv[4];
  {
    double dx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = xT[c] - xrT[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.QT2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.QT2(i, j), dx[j], s);
      pv[i] = s;
#pragma unroll
      for (int j = i; j < 4; ++j) P[sym(i, j)] = w.QT2(i, j);
    }
  }
  double x[4], u[2], xr[4], ur[2];
  LinD L;
  L.b0[0] = L.b0[1] = 0.0;
  auto load = [&](uint32_t n) {
    const RingStage s = ring_stage<RPB>(r, n, lane);
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = lds(s.base + kOffX + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) u[c] = lds(s.base + kOffU + c * 256);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      L.a[0][j] = lds(s.base + kOffA + j * 256);
      L.a[1][j] = lds(s.base + kOffA + (4 + j) * 256);
    }
    L.b[0] = lds(s.base + kOffA + 8 * 256);
    L.b[1] = lds(s.base + kOffA + 9 * 256);
    lds_ref<RPB>(s, xr, ur);
  };
  mbar_wait(ring_bar(r, r.seq), ring_parity(r.seq));
  load(r.seq);
  ++r.seq;
  double dJ = 0.0, sn = 0.0;
  const QhQ2<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R2(0, 0), w.R2(0, 1));
  double* pk = K + (steps - 1) * p.sk + lane;
  double* ps = S + (steps - 1) * p.ss + lane;
  for (int i = 0; i < steps; ++i) {
    const bool more = i + 1 < steps;
    const uint32_t ready = more ? mbar_test(ring_bar(r, r.seq), ring_parity(r.seq)) : 1u;
    double dx[4], du[2], q[4], rr[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = x[c] - xr[c];
#pragma unroll
    for (int c = 0; c < 2; ++c) du[c] = u[c] - ur[c];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double s = w.Q2(a, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.Q2(a, j), dx[j], s);
      q[a] = s;
    }
    rr[0] = fma(w.R2(0, 1), du[1], w.R2(0, 0) * du[0]);
    rr[1] = fma(w.R2(1, 1), du[1], w.R2(1, 0) * du[0]);
    double Kt[8], st[2];
    riccati_step<true, false>(P, pv, L, m.dt, Qh, col, w.R2(0, 0), w.R2(0, 1), w.R2(1, 1), q, rr, Kt, st, dJ);
    if (store) {
#pragma unroll
      for (int e = 0; e < 8; ++e) pk[e * 32] = Kt[e];
      ps[0] = st[0];
      ps[32] = st[1];
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double a = fabs(st[e]);
      sn = (a > sn || a != a) ? a : sn;  // NaN is sticky, like np.max(np.abs(sigma))
    }
    if (more) {  // operands of the next step (time index steps-2-i), straight into the registers just used
      if (!ready) mbar_wait(ring_bar(r, r.seq), ring_parity(r.seq));
      load(r.seq);
      ++r.seq;
    }
    // the stage read one iteration ago is free now: refill it
    __syncwarp();
    if (i + ACRO_RING_D < steps) ring_fill<RPB, false>(r, p, f, lane);
    pk -= p.sk;
    ps -= p.ss;
  }
  dJ_out = dJ;
  sn_out = sn;
}

// ---------------------------------------------------------------------------------------------------------
// kernel: one warp per block, block = tile of 32 problems
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB>
__global__ void __launch_bounds__(32) k_newton_ring(const __grid_constant__ NewtonArgs a) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const int64_t B = a.B, tile = blockIdx.x, b0 = tile * 32LL, b = b0 + lane;
  const bool valid = b < B;
  const int64_t bs = valid ? b : B - 1;  // padding lanes shadow the last problem and never write per-problem scalars
  const int N = a.N;
  const WV<WPB> w(a.kw, B, bs);
  Ring r;
  r.data = smem_u32(ring_smem);
  r.bars = r.data + ACRO_RING_D * stage_bytes<RPB>();
  r.seq = r.iss = 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < ACRO_RING_D; ++s) mbar_init(r.bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // doubles per time step of one tile (tile-major layout: compile-time strides)
  constexpr int64_t sx = 4 * 32, su = 2 * 32, sk = 8 * 32, ss = 2 * 32, sl = 10 * 32;
  const int64_t oN = tile * N, oM = tile * (N - 1);
  double* const tX[2] = {a.X + oN * sx, a.Xw + oN * sx};
  double* const tU[2] = {a.U + oM * su, a.Uw + oM * su};
  double* const tK = a.K + oM * sk;
  double* const tS = a.S + oM * ss;
  double* const tL = a.lin + oM * sl;
  TilePtrs p;
  p.k = tK;
  p.s = tS;
  p.lin = tL;
  p.rx = RPB ? a.rx + oN * sx : a.rx;
  p.ru = RPB ? a.ru + oM * su : a.ru;
  p.sx = sx;
  p.su = su;
  p.sk = sk;
  p.ss = ss;
  p.sl = sl;
  const RefV<RPB> ref{a.rx, a.ru, N, bs};
  double xrT[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xrT[c] = ref.X(N - 1, c);

  int it, st;
  double cost_k;
  if (a.o.init) {
    // u = 0 (or the caller's warm start), x = simulate_open_loop(x0, u), cost_k = total_cost(...)   (tg:311-319)
    double x[4];
    double* px = tX[0] + lane;
    double* pu = tU[0] + lane;
    double* pl = tL + lane;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = a.x0[c * B + bs];
      px[c * 32] = x[c];
    }
    double c_acc = 0.0;
    for (int t = 0; t < N - 1; ++t) {
      double u0 = 0.0, u1 = 0.0;
      if (a.o.init == 2) {
        u0 = pu[0];
        u1 = pu[32];
      } else {
        pu[0] = 0.0;
        pu[32] = 0.0;
      }
      double ex[4], eu[2] = {u0 - ref.U(t, 0), u1 - ref.U(t, 1)};
#pragma unroll
      for (int c = 0; c < 4; ++c) ex[c] = x[c] - ref.X(t, c);
      c_acc += quad4(ex, [&](int i, int j) { return w.Q(i, j); });
      c_acc += quad2(eu, [&](int i, int j) { return w.R(i, j); });
      double xn[4];
      LinD L;
      rk4_step_lin(a.m, x, u0, u1, xn, L);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pl[j * 32] = L.a[0][j];
        pl[(4 + j) * 32] = L.a[1][j];
      }
      pl[8 * 32] = L.b[0];
      pl[9 * 32] = L.b[1];
      px += sx;
      pu += su;
      pl += sl;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x[c] = xn[c];
        px[c * 32] = x[c];
      }
    }
    double ex[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ex[c] = x[c] - xrT[c];
    c_acc += quad4(ex, [&](int i, int j) { return w.QT(i, j); });
    cost_k = c_acc;
    it = 0;
    st = ACRO_RUNNING;
    if (a.h_cost && valid) a.h_cost[b] = cost_k;
    // the bulk copies of the first pass read what this warp has just written with ordinary stores
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  } else {
    it = a.iters[bs];
    st = a.status[bs];
    cost_k = a.cost[bs];
  }
  double dJ = a.o.init ? 0.0 : a.dJ[bs], sn = a.o.init ? 0.0 : a.sn[bs], gacc = a.o.init ? 0.0 : a.gacc[bs];
  bool run = valid && st == ACRO_RUNNING && it < a.o.max_iters;
  int cur = 0, home = 0, done = 0;
  while (__any_sync(FULL, run) && (a.o.chunk_iters <= 0 || done < a.o.chunk_iters)) {
    p.x = tX[cur];
    p.u = tU[cur];
    double* Xo = tX[cur ^ 1];
    double* Uo = tU[cur ^ 1];
    double xT[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xT[c] = p.x[(N - 1) * sx + c * 32 + lane];
    double dJn, snn;
    backward_ring<WPB, RPB>(a.m, w, N, p, r, lane, run, tK, tS, xT, xrT, dJn, snn);
    if (run) {
      dJ = dJn;
      sn = snn;
      if (a.h_sn) a.h_sn[int64_t(it) * B + b] = sn;
    }
    // K, S were written with ordinary stores and are read back by bulk copies
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    bool need = run, ok = false;
    double gamma = a.o.gamma_0, cn = 0.0;
    int tries = 0;
    for (int i = 0; i < a.o.max_line_search && __any_sync(FULL, need); ++i) {
      const double c = forward_ring<WPB, RPB>(a.m, w, N, p, r, lane, gamma, need, Xo, Uo, tL, xrT);
      if (need) {
        ++tries;
        // accept iff cost_new < cost_k + c*gamma*delta_J  (strict, NaN rejects)   tg:361
        const double thr = __dadd_rn(cost_k, __dmul_rn(__dmul_rn(a.o.c, gamma), dJ));
        if (c < thr) {
          ok = true;
          need = false;
          cn = c;
        } else {
          gamma = __dmul_rn(gamma, a.o.beta);  // tg:365
        }
      }
    }
    if (run) {
      if (a.h_ntry) a.h_ntry[int64_t(it) * B + b] = tries;
      ++it;
      if (!ok) {  // tg:367-369: keep the current iterate, stop
        st = ACRO_LINE_SEARCH_FAILED;
        if (a.h_gamma) a.h_gamma[int64_t(it - 1) * B + b] = nan("");
        home = cur;
      } else {
        cost_k = cn;
        gacc = gamma;
        home = cur ^ 1;
        if (a.h_gamma) a.h_gamma[int64_t(it - 1) * B + b] = gamma;
        if (a.h_cost) a.h_cost[int64_t(it) * B + b] = cost_k;
        if (sn < a.o.tol) st = ACRO_CONVERGED;  // tg:394-396
      }
      if (st == ACRO_RUNNING && it >= a.o.max_iters) st = ACRO_MAX_ITERS;
      run = (st == ACRO_RUNNING);
    }
    // candidates / lin were written with ordinary stores and are the next pass's bulk-copy sources
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    cur ^= 1;
    ++done;
  }
  if (valid && st == ACRO_RUNNING && it >= a.o.max_iters) st = ACRO_MAX_ITERS;
  if (home) {  // the final iterate of this problem sits in the workspace: move it home
    const double* sxp = tX[1] + lane;
    const double* sup = tU[1] + lane;
    double* dxp = tX[0] + lane;
    double* dup = tU[0] + lane;
    for (int t = 0; t < N; ++t) {
#pragma unroll
      for (int c = 0; c < 4; ++c) dxp[t * sx + c * 32] = sxp[t * sx + c * 32];
      if (t < N - 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c) dup[t * su + c * 32] = sup[t * su + c * 32];
      }
    }
  }
  if (valid) {
    a.cost[b] = cost_k;
    a.dJ[b] = dJ;
    a.sn[b] = sn;
    a.gacc[b] = gacc;
    a.iters[b] = it;
    a.status[b] = st;
  }
}

}  // namespace acro
