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

#ifndef ACRO_MODEL
// PPB: physical parameters per problem (`pb`, see model_per_problem); otherwise the shared model in the constant bank
#define ACRO_MODEL(PPB, m0, pb, B, b) \
  Model m_loc_;                        \
  if (PPB) m_loc_ = model_per_problem(m0, pb, B, b); \
  const Model& m = PPB ? m_loc_ : m0
#endif

#define ACRO_RING_D 3  // stages per ring

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
// The same wait as a real function call.  In a kernel whose blocks may hold more than two warps ptxas puts a YIELD at
// the head of every loop that contains an mbarrier phase check (try_wait or test_wait), which costs the warp-specialised
// loops of k_newton_spec a quarter of their speed; with the spin loop behind a call the loops stay YIELD-free
// (profiles/README.md, "YIELD").  NI = "no inlined phase checks".
__device__ __noinline__ void mbar_wait_call(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
template <bool NI>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity) {
  if (NI)
    mbar_wait_call(bar, parity);
  else
    mbar_wait(bar, parity);
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

// A stage holds SG consecutive time steps of every streamed array; a trajectory of a tile is contiguous
// in memory (tile-major layout), so SG steps of one array are ONE bulk copy.  Byte offsets inside a stage:
//     X: s*1024 + c*256   U: SG*1024 + s*512 + c*256   K: SG*1536 + s*2048 + c*256   S: SG*3584 + s*512 + c*256
//     (backward pass: lin: SG*1536 + s*2560 + j*256)   reference block at SG*4096
// SG is a template parameter: 16 when every SM holds at most one block (B <= 4736: 199 KB of shared memory per
// block), 4 otherwise (50 KB, four blocks per SM).
template <int SG>
struct StageOff {
  static constexpr uint32_t X = 0, U = SG * 1024, A = SG * 1536, S = SG * 3584, Ref = SG * 4096;
};
template <bool RPB, int SG>
__host__ __device__ constexpr uint32_t stage_bytes() { return SG * 4096 + (RPB ? SG * 1536 : ((SG * 48 + 127) / 128) * 128); }

struct Ring {
  uint32_t data, bars;  // shared addresses of slot 0 / barrier 0
  uint32_t base;        // stages consumed by earlier passes: stage k of this pass has sequence number base + k
};
__device__ __forceinline__ uint32_t ring_slot(uint32_t g) { return g % ACRO_RING_D; }
__device__ __forceinline__ uint32_t ring_parity(uint32_t g) { return (g / ACRO_RING_D) & 1u; }

// Tile base pointers (time step 0) of everything a pass streams.
struct TilePtrs {
  const double *x, *u, *k, *s, *lin;  // source iterate (current), gains, feed-forward, linearisation
  const double *rx, *ru;              // reference (shared: plain arrays; per problem: tile bases)
};
constexpr int64_t kSX = 4 * 32, kSU = 2 * 32, kSK = 8 * 32, kSS = 2 * 32, kSL = 10 * 32;  // doubles per time step

// Issue the bulk copies of stage k of a pass: time steps [t_lo, t_lo + cnt).
// RL ("recompute the linearisation"): the backward pass does not stream lin[t]; it linearises about (x_t, u_t) itself.
template <bool RPB, bool FWD, int SG, bool RL = false>
__device__ __forceinline__ void ring_fill(const Ring& r, const TilePtrs& p, int k, int t_lo, int cnt) {
  if (elect_one()) {
    const uint32_t g = r.base + k, slot = ring_slot(g);
    const uint32_t bar = r.bars + slot * 8, dst = r.data + slot * stage_bytes<RPB, SG>();
    const uint32_t n = (uint32_t)cnt;
    mbar_expect_tx(bar, n * ((!FWD && RL ? 1536u : 4096u) + (RPB ? 1536u : 48u)));
    bulk_g2s(dst + StageOff<SG>::X, p.x + t_lo * kSX, n * 1024, bar);
    bulk_g2s(dst + StageOff<SG>::U, p.u + t_lo * kSU, n * 512, bar);
    if (FWD) {
      bulk_g2s(dst + StageOff<SG>::A, p.k + t_lo * kSK, n * 2048, bar);
      bulk_g2s(dst + StageOff<SG>::S, p.s + t_lo * kSS, n * 512, bar);
    } else if (!RL) {
      bulk_g2s(dst + StageOff<SG>::A, p.lin + t_lo * kSL, n * 2560, bar);
    }
    if (RPB) {
      bulk_g2s(dst + StageOff<SG>::Ref, p.rx + t_lo * kSX, n * 1024, bar);
      bulk_g2s(dst + StageOff<SG>::Ref + SG * 1024, p.ru + t_lo * kSU, n * 512, bar);
    } else {
      bulk_g2s(dst + StageOff<SG>::Ref, p.rx + t_lo * 4, n * 32, bar);
      bulk_g2s(dst + StageOff<SG>::Ref + SG * 32, p.ru + t_lo * 2, n * 16, bar);
    }
  }
}

// reference of step s of a stage
template <bool RPB, int SG>
__device__ __forceinline__ void lds_ref(uint32_t stage, int s, int lane, double xr[4], double ur[2]) {
  const uint32_t b = stage + StageOff<SG>::Ref;
  if (RPB) {
#pragma unroll
    for (int c = 0; c < 4; ++c) xr[c] = lds(b + s * 1024 + c * 256 + lane * 8);
#pragma unroll
    for (int c = 0; c < 2; ++c) ur[c] = lds(b + SG * 1024 + s * 512 + c * 256 + lane * 8);
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) xr[c] = lds(b + s * 32 + c * 8);
#pragma unroll
    for (int c = 0; c < 2; ++c) ur[c] = lds(b + SG * 32 + s * 16 + c * 8);
  }
}

template <bool RPB, int SG>
__device__ __forceinline__ void lds_fwd(uint32_t stage, int s, int lane, StepIn& in) {
  const uint32_t b = stage + lane * 8;
#pragma unroll
  for (int c = 0; c < 4; ++c) in.x[c] = lds(b + StageOff<SG>::X + s * 1024 + c * 256);
#pragma unroll
  for (int c = 0; c < 2; ++c) in.u[c] = lds(b + StageOff<SG>::U + s * 512 + c * 256);
#pragma unroll
  for (int c = 0; c < 8; ++c) in.k[c] = lds(b + StageOff<SG>::A + s * 2048 + c * 256);
#pragma unroll
  for (int c = 0; c < 2; ++c) in.s[c] = lds(b + StageOff<SG>::S + s * 512 + c * 256);
  lds_ref<RPB, SG>(stage, s, lane, in.xr, in.ur);
}

// ---------------------------------------------------------------------------------------------------------
// forward pass (tg:218-252): closed-loop rollout with step size gamma + its cost; writes the candidate
// (Xo, Uo) and the linearisation about it when `store`.
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB, int SG, bool RL>
__device__ __forceinline__ double forward_ring(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                               int lane, double gamma, bool store, double* __restrict__ Xo,
                                               double* __restrict__ Uo, double* __restrict__ Lo, const double xrT[4]) {
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  for (int k = 0; k < ACRO_RING_D && k < n_stages; ++k) ring_fill<RPB, true, SG, RL>(r, p, k, k * SG, min(SG, steps - k * SG));
  double xp[4];
  StepIn in;
  mbar_wait(r.bars + ring_slot(r.base) * 8, ring_parity(r.base));
  lds_fwd<RPB, SG>(r.data + ring_slot(r.base) * stage_bytes<RPB, SG>(), 0, lane, in);
#pragma unroll
  for (int c = 0; c < 4; ++c) xp[c] = in.x[c];  // x+_0 = x_0
  TrigCarry tc;
  if (RL) tc = trig_carry_at(m, xp[0], xp[1]);
  double cost = 0.0;
  double* po_x = Xo + lane;
  double* po_u = Uo + lane;
  double* po_l = Lo + lane;
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = min(SG, steps - k * SG);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    const uint32_t nstage = r.data + ring_slot(g + 1) * stage_bytes<RPB, SG>();
    const uint32_t nbar = r.bars + ring_slot(g + 1) * 8, npar = ring_parity(g + 1);
    for (int s = 0; s < cnt; ++s) {
      const bool cross = (s + 1 == cnt) && (k + 1 < n_stages);  // the next step lives in the next stage
      const uint32_t ready = cross ? mbar_test(nbar, npar) : 1u;
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
      // operands of the next step straight into the registers this step no longer needs
      if (s + 1 < cnt) {
        lds_fwd<RPB, SG>(stage, s + 1, lane, in);
      } else if (cross) {
        if (!ready) mbar_wait(nbar, npar);
        lds_fwd<RPB, SG>(nstage, 0, lane, in);
      }
      // first step of a stage: every read of the previous stage has completed, its slot can be refilled
      if (s == 0 && k >= 1 && k - 1 + ACRO_RING_D < n_stages) {
        __syncwarp();
        const int kk = k - 1 + ACRO_RING_D;
        ring_fill<RPB, true, SG, RL>(r, p, kk, kk * SG, min(SG, steps - kk * SG));
      }
      double xn[4];
      if (RL) {
        rk4_step_inc(m, xp, up[0], up[1], xn, tc);
      } else {
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
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) xp[c] = xn[c];
      po_x += kSX;
      po_u += kSU;
      po_l += kSL;
    }
  }
  r.base += n_stages;
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
// Stage k of this pass holds the time steps [t_lo, t_hi] with t_hi = steps-1 - k*SG, walked downwards.
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB, int SG, bool RL>
__device__ __forceinline__ void backward_ring(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                              int lane, bool store, double* __restrict__ K, double* __restrict__ S,
                                              const double xT[4], const double xrT[4], double& dJ_out,
                                              double& sn_out) {
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  auto t_lo_of = [&](int k) { return max(0, steps - (k + 1) * SG); };
  auto cnt_of = [&](int k) { return (steps - k * SG) - t_lo_of(k); };
  for (int k = 0; k < ACRO_RING_D && k < n_stages; ++k) ring_fill<RPB, false, SG, RL>(r, p, k, t_lo_of(k), cnt_of(k));
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
  auto load = [&](uint32_t stage, int s) {
    const uint32_t b = stage + lane * 8;
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = lds(b + StageOff<SG>::X + s * 1024 + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) u[c] = lds(b + StageOff<SG>::U + s * 512 + c * 256);
    if (!RL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        L.a[0][j] = lds(b + StageOff<SG>::A + s * 2560 + j * 256);
        L.a[1][j] = lds(b + StageOff<SG>::A + s * 2560 + (4 + j) * 256);
      }
      L.b[0] = lds(b + StageOff<SG>::A + s * 2560 + 8 * 256);
      L.b[1] = lds(b + StageOff<SG>::A + s * 2560 + 9 * 256);
    }
    lds_ref<RPB, SG>(stage, s, lane, xr, ur);
  };
  mbar_wait(r.bars + ring_slot(r.base) * 8, ring_parity(r.base));
  load(r.data + ring_slot(r.base) * stage_bytes<RPB, SG>(), cnt_of(0) - 1);
  double dJ = 0.0, sn = 0.0;
  const QhQ2<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R2(0, 0), w.R2(0, 1));
  double* pk = K + (steps - 1) * kSK + lane;
  double* ps = S + (steps - 1) * kSS + lane;
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = cnt_of(k);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    const uint32_t nstage = r.data + ring_slot(g + 1) * stage_bytes<RPB, SG>();
    const uint32_t nbar = r.bars + ring_slot(g + 1) * 8, npar = ring_parity(g + 1);
    const int ncnt = (k + 1 < n_stages) ? cnt_of(k + 1) : 0;
    for (int s = cnt - 1; s >= 0; --s) {
      const bool cross = (s == 0) && (k + 1 < n_stages);
      const uint32_t ready = cross ? mbar_test(nbar, npar) : 1u;
      double dx[4], du[2], q[4], rr[2];
#pragma unroll
      for (int c = 0; c < 4; ++c) dx[c] = x[c] - xr[c];
#pragma unroll
      for (int c = 0; c < 2; ++c) du[c] = u[c] - ur[c];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double acc = w.Q2(a, 0) * dx[0];
#pragma unroll
        for (int j = 1; j < 4; ++j) acc = fma(w.Q2(a, j), dx[j], acc);
        q[a] = acc;
      }
      rr[0] = fma(w.R2(0, 1), du[1], w.R2(0, 0) * du[0]);
      rr[1] = fma(w.R2(1, 1), du[1], w.R2(1, 0) * du[0]);
      if (RL) {
        const LinD Lr = linearize_d(m, x, u[0], u[1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          L.a[0][j] = Lr.a[0][j];
          L.a[1][j] = Lr.a[1][j];
        }
        L.b[0] = Lr.b[0];
        L.b[1] = Lr.b[1];
      }
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
      // operands of the next (earlier) time step into the registers just used
      if (s > 0) {
        load(stage, s - 1);
      } else if (cross) {
        if (!ready) mbar_wait(nbar, npar);
        load(nstage, ncnt - 1);
      }
      // last step of the first stage-visit: refill the slot of the previous stage (all of its reads are done)
      if (s == cnt - 1 && k >= 1 && k - 1 + ACRO_RING_D < n_stages) {
        __syncwarp();
        const int kk = k - 1 + ACRO_RING_D;
        ring_fill<RPB, false, SG, RL>(r, p, kk, t_lo_of(kk), cnt_of(kk));
      }
      pk -= kSK;
      ps -= kSS;
    }
  }
  r.base += n_stages;
  dJ_out = dJ;
  sn_out = sn;
}

// ---------------------------------------------------------------------------------------------------------
// kernel: one warp per block, block = tile of 32 problems
// ---------------------------------------------------------------------------------------------------------
// PPB: every problem its own physical parameters (a.pb, see model_per_problem)
template <bool WPB, bool RPB, int SG, bool RL = false, bool PPB = false>
__global__ void __launch_bounds__(32) k_newton_ring(const __grid_constant__ NewtonArgs a) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const int64_t B = a.B, tile = blockIdx.x, b0 = tile * 32LL, b = b0 + lane;
  const bool valid = b < B;
  const int64_t bs = valid ? b : B - 1;  // padding lanes shadow the last problem and never write per-problem scalars
  const int N = a.N;
  ACRO_MODEL(PPB, a.m, a.pb, B, bs);
  const WV<WPB> w(a.kw, B, bs);
  Ring r;
  r.data = smem_u32(ring_smem);
  r.bars = r.data + ACRO_RING_D * stage_bytes<RPB, SG>();
  r.base = 0;
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
      if (RL) {
        rk4_step(m, x, u0, u1, xn);
      } else {
        LinD L;
        rk4_step_lin(m, x, u0, u1, xn, L);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pl[j * 32] = L.a[0][j];
          pl[(4 + j) * 32] = L.a[1][j];
        }
        pl[8 * 32] = L.b[0];
        pl[9 * 32] = L.b[1];
      }
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
    backward_ring<WPB, RPB, SG, RL>(m, w, N, p, r, lane, run, tK, tS, xT, xrT, dJn, snn);
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
      const double c = forward_ring<WPB, RPB, SG, RL>(m, w, N, p, r, lane, gamma, need, Xo, Uo, tL, xrT);
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
