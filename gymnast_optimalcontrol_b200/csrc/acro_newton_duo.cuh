// acro_newton_duo.cuh - the Newton / Armijo loop (tg:298-398) for batches that leave most of the chip idle
// (at most one tile of 32 problems per SM, e.g. config 2: B = 4096 = 128 tiles on 148 SMs).
//
// A warp-wide FP64 instruction occupies the FP64 pipe of its SM sub-partition for 2.25 cycles however many
// lanes are active, so a tile advances at the speed at which ONE sub-partition issues the ~600 FP64 instructions
// of a time-step pair; the other three sub-partitions of the SM idle.  This kernel therefore splits every pass
// of a tile between TWO warps on two sub-partitions of the same SM:
//
//   warp 0 ("chain")  : only what the recurrence itself needs
//                        forward : u+ = u + K (x+ - x) + gamma sigma ; x+ <- RK4(x+, u+)        (tg:218-229)
//                        backward: P-recursion, G = R + B'PB, K = -G^-1 B'PA                       (tg:195-213)
//   warp 1 ("trailer"): everything that hangs off the chain, a few steps behind
//                        forward : cost (tg:231-252), linearisation A_d, B_d about the new iterate
//                                  (dynamics.py:217-226, tg:161-164), all global stores
//                        backward: q, r (tg:89-129), g, sigma, delta_J, the costate p, max|sigma|,
//                                  all global stores; it also feeds the TMA ring for both warps
//
// The chain hands each step to the trailer through a ring of ACRO_DUO_R shared-memory slots guarded by
// full / empty mbarriers (forward: x+_t, u+_t; backward: K_t and the two scalars of the 2x2 factorisation).
// The trailer never feeds back into a pass; the two warps meet at a named barrier at the pass boundaries,
// where the chain warp (which owns the Armijo / convergence logic) publishes the next command.
// The arithmetic of both halves is expression-for-expression that of k_newton_ring.
#pragma once
#include "acro_newton_ring.cuh"

namespace acro {

#ifndef ACRO_DUO_R
#define ACRO_DUO_R 8  // hand-off slots (power of two)
#endif
#define ACRO_DUO_SLOT_BYTES 2560  // 10 rows of 32 doubles

__device__ __forceinline__ void sts(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// arrive (count 1) by lane 0 only, predicated: no branch
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t bar, int lane) {
  asm volatile("{\n\t.reg .pred P1;\n\tsetp.eq.s32 P1, %1, 0;\n\t@P1 mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar),
               "r"(lane)
               : "memory");
}
__device__ __forceinline__ void duo_bar() {
  __syncwarp();
  asm volatile("bar.sync 1, 64;" ::: "memory");
}

// NI variants of the chain loops (k_newton_spec) contain no mbarrier phase check: instead of testing the next hand-off
// slot in every step, every ACRO_DUO_LOOK-th step makes sure (blocking, behind a call) that the next ACRO_DUO_LOOK slots
// are free - the trailer frees them in order, so the one furthest ahead suffices.
#define ACRO_DUO_LOOK 4
// ... and the trailer waits for batches of ACRO_DUO_LOOK_T hand-offs.  With ACRO_DUO_R = 8 slots the chain never waits for
// the trailer as long as LOOK_T * (trailer step) + (wake-up latency) < (R - LOOK + 1) * (chain step): batches of 4 made the
// trailer the bottleneck of the backward pass (637 instead of 420 cycles per step, profiles/r2_spec_timing_*.txt), 3 do not.
#ifndef ACRO_DUO_LOOK_T
#define ACRO_DUO_LOOK_T 3
#endif
struct Hand {
  uint32_t data, full, empty;  // shared addresses: slot 0, full barrier 0, empty barrier 0
  uint32_t h;                  // hand-offs so far (identical in both warps)
  int zmask;                   // 0, read from shared memory at run time: `v & zmask` is a zero that depends on v
                               // as far as nvcc and ptxas can tell (a scheduling dependency without arithmetic)
  __device__ __forceinline__ uint32_t slot() const { return data + (h & (ACRO_DUO_R - 1)) * ACRO_DUO_SLOT_BYTES; }
  __device__ __forceinline__ uint32_t full_bar() const { return full + (h & (ACRO_DUO_R - 1)) * 8; }
  __device__ __forceinline__ uint32_t empty_bar() const { return empty + (h & (ACRO_DUO_R - 1)) * 8; }
  __device__ __forceinline__ uint32_t phase() const { return (h / ACRO_DUO_R) & 1u; }
  // hand-off number h + d: its empty barrier and the parity whose completion means "slot free"
  __device__ __forceinline__ uint32_t empty_bar_at(uint32_t d) const { return empty + ((h + d) & (ACRO_DUO_R - 1)) * 8; }
  __device__ __forceinline__ uint32_t free_parity_at(uint32_t d) const { return (((h + d) / ACRO_DUO_R) & 1u) ^ 1u; }
  // hand-off number hh (absolute): its full barrier and the parity whose completion means "produced"
  __device__ __forceinline__ uint32_t full_bar_of(uint32_t hh) const { return full + (hh & (ACRO_DUO_R - 1)) * 8; }
  __device__ __forceinline__ uint32_t full_parity_of(uint32_t hh) const { return (hh / ACRO_DUO_R) & 1u; }
};
// The trailer's side of the same idea (NI): every ACRO_DUO_LOOK-th hand-off it waits (behind a call) until the chain has
// produced the next ACRO_DUO_LOOK hand-offs - or the last one of the pass - and then reads them without waiting.
#ifndef ACRO_SPEC_TRAILER_NI
#define ACRO_SPEC_TRAILER_NI 1  // 0: the NI trailers keep the inlined per-step wait (ptxas then puts a YIELD in their loops)
#endif
// h_ok: hand-offs below this number are known to have been produced (starts at hd.h when a pass begins)
template <bool NI>
__device__ __forceinline__ void hand_wait_full(const Hand& hd, uint32_t h_last, uint32_t& h_ok) {
  if (NI && ACRO_SPEC_TRAILER_NI) {
    if (hd.h >= h_ok) {
      const uint32_t hh = min(hd.h + (ACRO_DUO_LOOK_T - 1), h_last);
      mbar_wait_call(hd.full_bar_of(hh), hd.full_parity_of(hh));
      h_ok = hh + 1u;
    }
  } else {
    mbar_wait(hd.full_bar(), hd.phase());
  }
}

enum { DUO_EXIT = 0, DUO_BACKWARD = 1, DUO_FORWARD = 2 };

// ---------------------------------------------------------------------------------------------------------
// The two halves of riccati_step<true, false> (acro_device.cuh), same expressions in the same order.
// ---------------------------------------------------------------------------------------------------------
template <int SW>
__device__ __forceinline__ void riccati_trailer_step(double p[4], const LinD& L, double dt, const Lu2Col& col,
                                                     const double K[8], double inv_u11, double qsel,
                                                     const double qv[4], const double r[2], double sig[2],
                                                     double& dJ) {
  const double g0 = r[0];
  const double g1 = r[1] + fma(L.b[1], p[3], L.b[0] * p[2]);
  const bool swap = (SW == 2) ? col.swap : (SW == 1);
  const double y0 = swap ? -g1 : -g0, y1 = swap ? -g0 : -g1;
  sig[1] = fma(-col.l, y0, y1) * inv_u11;
  sig[0] = fma(-qsel, sig[1], y0) * col.inv_p;
  dJ += fma(g1, sig[1], g0 * sig[0]);
  const double p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
  const double ap[4] = {fma(L.a[1][0], p3, fma(L.a[0][0], p2, p0)), fma(L.a[1][1], p3, fma(L.a[0][1], p2, p1)),
                        fma(L.a[1][2], p3, fma(L.a[0][2], p2, dt * p0)),
                        fma(L.a[1][3], p3, fma(L.a[0][3], p2, dt * p1))};
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = qv[i] + ap[i] + fma(K[4 + i], g1, K[i] * g0);
}

// ---------------------------------------------------------------------------------------------------------
// forward pass, chain warp
// ---------------------------------------------------------------------------------------------------------
template <int SG>
struct FwdIn {
  double x[4], u[2], k[8], s[2];
  __device__ __forceinline__ void load(uint32_t stage, int st, int lane) {
    const uint32_t b = stage + lane * 8;
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = lds(b + StageOff<SG>::X + st * 1024 + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) u[c] = lds(b + StageOff<SG>::U + st * 512 + c * 256);
#pragma unroll
    for (int c = 0; c < 8; ++c) k[c] = lds(b + StageOff<SG>::A + st * 2048 + c * 256);
#pragma unroll
    for (int c = 0; c < 2; ++c) s[c] = lds(b + StageOff<SG>::S + st * 512 + c * 256);
  }
};

template <bool RPB, int SG, bool NI = false>
__device__ __forceinline__ void duo_forward_chain(const Model& m, int N, Ring& r, Hand& hd, int lane, double gamma) {
  constexpr unsigned FULL = 0xffffffffu;
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  double xp[4];
  FwdIn<SG> in;
  mbar_wait_t<NI>(r.bars + ring_slot(r.base) * 8, ring_parity(r.base));
  in.load(r.data + ring_slot(r.base) * stage_bytes<RPB, SG>(), 0, lane);
#pragma unroll
  for (int c = 0; c < 4; ++c) xp[c] = in.x[c];  // x+_0 = x_0
  mbar_wait_t<NI>(hd.empty_bar(), hd.phase() ^ 1u);  // the hand-off slot of the first step is free (all are, between passes)
  // State carried into the next step's single "anything unusual?" branch, so that the step itself is straight-line
  // code: is the next hand-off slot free, and did the step just taken leave the range of the polynomial sincos.
  uint32_t hready = 1u;
  bool bad = false;
  TrigCarry tc = trig_carry_at(m, xp[0], xp[1]);
  const RotC rotc = RotC::held();
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = min(SG, steps - k * SG);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    const uint32_t nstage = r.data + ring_slot(g + 1) * stage_bytes<RPB, SG>();
    const uint32_t nbar = r.bars + ring_slot(g + 1) * 8, npar = ring_parity(g + 1);
    for (int s = 0; s < cnt; ++s) {
      // where the operands of the next step live: this stage, the next stage (whose bulk copies were issued at
      // least a stage ago: the wait is a formality), or nowhere (last step of the pass: reload this step's)
      const bool cross = (s + 1 == cnt) && (k + 1 < n_stages);
      const bool look = NI && (hd.h % ACRO_DUO_LOOK) == 0;
      if (__any_sync(FULL, cross || bad || (NI ? look : !hready))) {
        if (bad) {  // the last step left the range of the incremental sincos: redo it with a full sincos per stage
          double xo[4], uo[2];  // its inputs are still in the hand-off slot they were written to
          const uint32_t pslot = hd.data + ((hd.h - 1) & (ACRO_DUO_R - 1)) * ACRO_DUO_SLOT_BYTES + lane * 8;
#pragma unroll
          for (int c = 0; c < 4; ++c) xo[c] = lds(pslot + c * 256);
#pragma unroll
          for (int c = 0; c < 2; ++c) uo[c] = lds(pslot + (4 + c) * 256);
          const Vec4 o = rk4_step_redo(m, xo[0], xo[1], xo[2], xo[3], uo[0], uo[1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) xp[i] = o.v[i];
          tc = trig_carry_at(m, xp[0], xp[1]);
        }
        __syncwarp();
        if (cross) mbar_wait_t<NI>(nbar, npar);
        if (NI) {
          if (look) mbar_wait_call(hd.empty_bar_at(ACRO_DUO_LOOK - 1), hd.free_parity_at(ACRO_DUO_LOOK - 1));
        } else if (!hready) {
          mbar_wait(hd.empty_bar(), hd.phase() ^ 1u);  // the trailer is ACRO_DUO_R steps behind
        }
      }
      const uint32_t nsrc = cross ? nstage : stage;
      const int ns = (s + 1 < cnt) ? s + 1 : (cross ? 0 : s);
      const uint32_t slot = hd.slot(), fbar = hd.full_bar();
#pragma unroll
      for (int c = 0; c < 4; ++c) sts(slot + c * 256 + lane * 8, xp[c]);
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
#pragma unroll
      for (int c = 0; c < 2; ++c) sts(slot + (4 + c) * 256 + lane * 8, up[c]);
      __syncwarp();
      mbar_arrive_lane0(fbar, lane);
      ++hd.h;
      double xn[4];
      const uint32_t nebar = hd.empty_bar(), nepar = hd.phase() ^ 1u;
      bad = rk4_step_rot<false>(m, rotc, xp, up[0], up[1], xn, tc, [&]() {
        // operands of the next step and the state of the next hand-off slot, while the FP64 pipe is busy
        // (no branch in here: a branch would cut the step's straight-line code into two scheduling regions)
        in.load(nsrc, ns, lane);
        if (!NI) hready = mbar_test(nebar, nepar);
        int acc = __double2loint(in.x[0]) | __double2loint(in.x[1]) | __double2loint(in.x[2]) | __double2loint(in.x[3]) |
                  __double2loint(in.u[0]) | __double2loint(in.u[1]) | __double2loint(in.s[0]) | __double2loint(in.s[1]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc |= __double2loint(in.k[e]);
        return acc & hd.zmask;
      });
#pragma unroll
      for (int c = 0; c < 4; ++c) xp[c] = xn[c];
    }
  }
  r.base += n_stages;
  if (__any_sync(FULL, bad || !hready)) {
    if (bad) {
      double xo[4], uo[2];
      const uint32_t pslot = hd.data + ((hd.h - 1) & (ACRO_DUO_R - 1)) * ACRO_DUO_SLOT_BYTES + lane * 8;
#pragma unroll
      for (int c = 0; c < 4; ++c) xo[c] = lds(pslot + c * 256);
#pragma unroll
      for (int c = 0; c < 2; ++c) uo[c] = lds(pslot + (4 + c) * 256);
      const Vec4 o = rk4_step_redo(m, xo[0], xo[1], xo[2], xo[3], uo[0], uo[1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) xp[i] = o.v[i];
    }
    __syncwarp();
    if (!hready) mbar_wait(hd.empty_bar(), hd.phase() ^ 1u);
  }
  if (NI) mbar_wait_call(hd.empty_bar(), hd.phase() ^ 1u);
  {  // terminal state
    const uint32_t slot = hd.slot(), fbar = hd.full_bar();
#pragma unroll
    for (int c = 0; c < 4; ++c) sts(slot + c * 256 + lane * 8, xp[c]);
    __syncwarp();
    mbar_arrive_lane0(fbar, lane);
    ++hd.h;
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward pass, trailer warp: cost of the candidate, its linearisation, all stores, ring refills
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB, int SG, bool NI = false>
__device__ __forceinline__ double duo_forward_trailer(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                                      Hand& hd, int lane, bool store, double* __restrict__ Xo,
                                                      double* __restrict__ Uo, double* __restrict__ Lo,
                                                      const double xrT[4]) {
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  for (int k = 0; k < ACRO_RING_D && k < n_stages; ++k) ring_fill<RPB, true, SG>(r, p, k, k * SG, min(SG, steps - k * SG));
  double cost = 0.0;
  const uint32_t h_last = hd.h + uint32_t(steps);  // the terminal state is hand-off number steps of this pass
  uint32_t h_ok = hd.h;
  double* po_x = Xo + lane;
  double* po_u = Uo + lane;
  double* po_l = Lo + lane;
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = min(SG, steps - k * SG);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    mbar_wait_t<NI>(r.bars + ring_slot(g) * 8, ring_parity(g));
    for (int s = 0; s < cnt; ++s) {
      double xr[4], ur[2], xp[4], up[2];
      lds_ref<RPB, SG>(stage, s, lane, xr, ur);
      const uint32_t slot = hd.slot();
      hand_wait_full<NI>(hd, h_last, h_ok);
#pragma unroll
      for (int c = 0; c < 4; ++c) xp[c] = lds(slot + c * 256 + lane * 8);
#pragma unroll
      for (int c = 0; c < 2; ++c) up[c] = lds(slot + (4 + c) * 256 + lane * 8);
      __syncwarp();
      mbar_arrive_lane0(hd.empty_bar(), lane);
      ++hd.h;
      if (store) {
#pragma unroll
        for (int c = 0; c < 4; ++c) po_x[c * 32] = xp[c];
#pragma unroll
        for (int c = 0; c < 2; ++c) po_u[c * 32] = up[c];
      }
      double ex[4], eu[2];
#pragma unroll
      for (int c = 0; c < 4; ++c) ex[c] = xp[c] - xr[c];
#pragma unroll
      for (int c = 0; c < 2; ++c) eu[c] = up[c] - ur[c];
      cost += quad4(ex, [&](int i, int j) { return w.Q(i, j); });
      cost += quad2(eu, [&](int i, int j) { return w.R(i, j); });
      // first step of a stage: both warps are done with the previous stage (the chain warp read it before it
      // produced this step), its slot can be refilled
      if (s == 0 && k >= 1 && k - 1 + ACRO_RING_D < n_stages) {
        __syncwarp();
        const int kk = k - 1 + ACRO_RING_D;
        ring_fill<RPB, true, SG>(r, p, kk, kk * SG, min(SG, steps - kk * SG));
      }
      const LinD L = linearize_d(m, xp, up[0], up[1]);
      if (store) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          po_l[j * 32] = L.a[0][j];
          po_l[(4 + j) * 32] = L.a[1][j];
        }
        po_l[8 * 32] = L.b[0];
        po_l[9 * 32] = L.b[1];
      }
      po_x += kSX;
      po_u += kSU;
      po_l += kSL;
    }
  }
  r.base += n_stages;
  double xp[4], ex[4];
  {
    const uint32_t slot = hd.slot();
    if (NI && ACRO_SPEC_TRAILER_NI)
      mbar_wait_call(hd.full_bar(), hd.phase());
    else
      mbar_wait(hd.full_bar(), hd.phase());
#pragma unroll
    for (int c = 0; c < 4; ++c) xp[c] = lds(slot + c * 256 + lane * 8);
    __syncwarp();
    mbar_arrive_lane0(hd.empty_bar(), lane);
    ++hd.h;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (store) po_x[c * 32] = xp[c];
    ex[c] = xp[c] - xrT[c];
  }
  cost += quad4(ex, [&](int i, int j) { return w.QT(i, j); });
  return cost;
}

// ---------------------------------------------------------------------------------------------------------
// backward pass, chain warp: the P recursion and the gains
// ---------------------------------------------------------------------------------------------------------
template <int SG>
__device__ __forceinline__ void lds_lin(uint32_t stage, int s, int lane, LinD& L) {
  const uint32_t b = stage + lane * 8 + StageOff<SG>::A + s * 2560;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    L.a[0][j] = lds(b + j * 256);
    L.a[1][j] = lds(b + (4 + j) * 256);
  }
  L.b[0] = lds(b + 8 * 256);
  L.b[1] = lds(b + 9 * 256);
}

template <bool WPB, bool RPB, int SG, int SW, bool NI = false>
__device__ __forceinline__ void duo_backward_chain(const Model& m, const WV<WPB>& w, int N, Ring& r, Hand& hd, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  auto t_lo_of = [&](int k) { return max(0, steps - (k + 1) * SG); };
  auto cnt_of = [&](int k) { return (steps - k * SG) - t_lo_of(k); };
  double P[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = w.QT2(i, j);
  LinD L;
  L.b0[0] = L.b0[1] = 0.0;
  mbar_wait_t<NI>(r.bars + ring_slot(r.base) * 8, ring_parity(r.base));
  lds_lin<SG>(r.data + ring_slot(r.base) * stage_bytes<RPB, SG>(), cnt_of(0) - 1, lane, L);
  const QhQ2<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R2(0, 0), w.R2(0, 1));
  mbar_wait_t<NI>(hd.empty_bar(), hd.phase() ^ 1u);  // the hand-off slot of the first step is free (all are, between passes)
  uint32_t hready = 1u;
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = cnt_of(k);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    const uint32_t nstage = r.data + ring_slot(g + 1) * stage_bytes<RPB, SG>();
    const uint32_t nbar = r.bars + ring_slot(g + 1) * 8, npar = ring_parity(g + 1);
    const int ncnt = (k + 1 < n_stages) ? cnt_of(k + 1) : 0;
    for (int s = cnt - 1; s >= 0; --s) {
      // one branch per step for everything unusual, so that the step itself is straight-line code (see the
      // forward pass): the next stage's bulk copies (a formality) and a trailer that is ACRO_DUO_R steps behind
      const bool cross = (s == 0) && (k + 1 < n_stages);
      // (no vote here: with __any_sync in this loop ptxas puts a YIELD at the head of both chain loops, 50 cycles
      // per step; the condition is the same in every lane)
      if constexpr (NI) {
        const bool look = (hd.h % ACRO_DUO_LOOK) == 0;
        if (cross || look) {
          if (cross) mbar_wait_call(nbar, npar);
          if (look) mbar_wait_call(hd.empty_bar_at(ACRO_DUO_LOOK - 1), hd.free_parity_at(ACRO_DUO_LOOK - 1));
        }
      } else {
        if (cross || !hready) {
          if (cross) mbar_wait(nbar, npar);
          if (!hready) mbar_wait(hd.empty_bar(), hd.phase() ^ 1u);
        }
      }
      const uint32_t nsrc = cross ? nstage : stage;
      const int ns = (s > 0) ? s - 1 : (cross ? ncnt - 1 : s);
      double Kt[8], inv_u11, qsel;
      riccati_chain_step<SW>(P, L, m.dt, Qh, col, w.R2(0, 1), w.R2(1, 1), Kt, inv_u11, qsel);
      // the linearisation of the next (earlier) time step, tied to the hand-off stores by an opaque zero so that
      // the loads are issued while the P update keeps the FP64 pipe busy
      LinD Ln;
      Ln.b0[0] = Ln.b0[1] = 0.0;
      lds_lin<SG>(nsrc, ns, lane, Ln);
      int acc = __double2loint(Ln.b[0]) | __double2loint(Ln.b[1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc |= __double2loint(Ln.a[0][j]) | __double2loint(Ln.a[1][j]);
      const uint32_t slot = hd.slot() + uint32_t(acc & hd.zmask), fbar = hd.full_bar();
#pragma unroll
      for (int e = 0; e < 8; ++e) sts(slot + e * 256 + lane * 8, Kt[e]);
      sts(slot + 8 * 256 + lane * 8, inv_u11);
      sts(slot + 9 * 256 + lane * 8, qsel);
      __syncwarp();
      mbar_arrive_lane0(fbar, lane);
      ++hd.h;
      if (!NI) hready = mbar_test(hd.empty_bar(), hd.phase() ^ 1u);
      L = Ln;
    }
  }
  r.base += n_stages;
}

// ---------------------------------------------------------------------------------------------------------
// backward pass, trailer warp: cost gradients, sigma, delta_J, costate, stores, ring refills
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB, int SG, int SW, bool NI = false>
__device__ __forceinline__ void duo_backward_trailer(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                                     Hand& hd, int lane, bool store, double* __restrict__ K,
                                                     double* __restrict__ S, const double xT[4], const double xrT[4],
                                                     double& dJ_out, double& sn_out) {
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  auto t_lo_of = [&](int k) { return max(0, steps - (k + 1) * SG); };
  auto cnt_of = [&](int k) { return (steps - k * SG) - t_lo_of(k); };
  for (int k = 0; k < ACRO_RING_D && k < n_stages; ++k) ring_fill<RPB, false, SG>(r, p, k, t_lo_of(k), cnt_of(k));
  double pv[4];
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
    }
  }
  double dJ = 0.0, sn = 0.0;
  const uint32_t h_last = hd.h + uint32_t(steps) - 1u;  // last hand-off of this pass
  uint32_t h_ok = hd.h;
  const Lu2Col col = lu2_col(w.R2(0, 0), w.R2(0, 1));
  double* pk = K + (steps - 1) * kSK + lane;
  double* ps = S + (steps - 1) * kSS + lane;
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = cnt_of(k);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    mbar_wait_t<NI>(r.bars + ring_slot(g) * 8, ring_parity(g));
    for (int s = cnt - 1; s >= 0; --s) {
      double x[4], u[2], xr[4], ur[2];
      LinD L;
      L.b0[0] = L.b0[1] = 0.0;
      {
        const uint32_t b = stage + lane * 8;
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = lds(b + StageOff<SG>::X + s * 1024 + c * 256);
#pragma unroll
        for (int c = 0; c < 2; ++c) u[c] = lds(b + StageOff<SG>::U + s * 512 + c * 256);
      }
      lds_lin<SG>(stage, s, lane, L);
      lds_ref<RPB, SG>(stage, s, lane, xr, ur);
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
      double Kt[8], inv_u11, qsel, st[2];
      {
        const uint32_t slot = hd.slot();
        hand_wait_full<NI>(hd, h_last, h_ok);
#pragma unroll
        for (int e = 0; e < 8; ++e) Kt[e] = lds(slot + e * 256 + lane * 8);
        inv_u11 = lds(slot + 8 * 256 + lane * 8);
        qsel = lds(slot + 9 * 256 + lane * 8);
        __syncwarp();
        mbar_arrive_lane0(hd.empty_bar(), lane);
        ++hd.h;
      }
      riccati_trailer_step<SW>(pv, L, m.dt, col, Kt, inv_u11, qsel, q, rr, st, dJ);
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
      // every read of the previous stage (by both warps) has completed: refill its slot
      if (s == cnt - 1 && k >= 1 && k - 1 + ACRO_RING_D < n_stages) {
        __syncwarp();
        const int kk = k - 1 + ACRO_RING_D;
        ring_fill<RPB, false, SG>(r, p, kk, t_lo_of(kk), cnt_of(kk));
      }
      pk -= kSK;
      ps -= kSS;
    }
  }
  r.base += n_stages;
  dJ_out = dJ;
  sn_out = sn;
}

// shared-memory map
template <bool RPB, int SG>
struct DuoSmem {
  static constexpr uint32_t ring = 0;
  static constexpr uint32_t hand = ACRO_RING_D * stage_bytes<RPB, SG>();
  static constexpr uint32_t res = hand + ACRO_DUO_R * ACRO_DUO_SLOT_BYTES;  // 3 rows of 32 doubles
  static constexpr uint32_t flags = res + 3 * 256;                          // 32 ints
  static constexpr uint32_t cmd = flags + 128;                              // 2 ints + the zero word (+ padding)
  static constexpr uint32_t bars = cmd + 16;  // ring full[D], hand full[R], hand empty[R]
  static constexpr uint32_t total = bars + (ACRO_RING_D + 2 * ACRO_DUO_R) * 8;
};

// ---------------------------------------------------------------------------------------------------------
// kernel: two warps per block, block = tile of 32 problems
// ---------------------------------------------------------------------------------------------------------
// PPB: every problem its own physical parameters (a.pb, see model_per_problem): the lumped model constants live in
// registers of both warps instead of the constant bank.
template <bool WPB, bool RPB, int SG, bool PPB = false>
__global__ void __launch_bounds__(64) k_newton_duo(const __grid_constant__ NewtonArgs a) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  using SM = DuoSmem<RPB, SG>;
  const int lane = threadIdx.x & 31;
  const bool chain = threadIdx.x < 32;
  const int64_t B = a.B, tile = blockIdx.x, b0 = tile * 32LL, b = b0 + lane;
  const bool valid = b < B;
  const int64_t bs = valid ? b : B - 1;  // padding lanes shadow the last problem and never write
  const int N = a.N;
  ACRO_MODEL(PPB, a.m, a.pb, B, bs);
  const WV<WPB> w(a.kw, B, bs);
  const uint32_t sbase = smem_u32(ring_smem);
  Ring r;
  r.data = sbase + SM::ring;
  r.bars = sbase + SM::bars;
  r.base = 0;
  Hand hd;
  hd.data = sbase + SM::hand;
  hd.full = sbase + SM::bars + ACRO_RING_D * 8;
  hd.empty = hd.full + ACRO_DUO_R * 8;
  hd.h = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < ACRO_RING_D + 2 * ACRO_DUO_R; ++s) mbar_init(r.bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    reinterpret_cast<volatile int*>(ring_smem + SM::cmd)[2] = 0;
  }
  __syncthreads();
  hd.zmask = reinterpret_cast<volatile int*>(ring_smem + SM::cmd)[2];
  volatile double* const res = reinterpret_cast<volatile double*>(ring_smem + SM::res);
  volatile int* const flags = reinterpret_cast<volatile int*>(ring_smem + SM::flags);
  volatile int* const cmd = reinterpret_cast<volatile int*>(ring_smem + SM::cmd);

  constexpr int64_t sx = 4 * 32, su = 2 * 32, sk = 8 * 32, ss = 2 * 32, sl = 10 * 32;
  const int64_t oN = tile * N, oM = tile * (N - 1);
  double* const tX[2] = {a.X + oN * sx, a.Xw + oN * sx};
  double* const tU[2] = {a.U + oM * su, a.Uw + oM * su};
  double* const tK = a.K + oM * sk;
  double* const tS = a.S + oM * ss;
  double* const tL = a.lin + oM * sl;
  const RefV<RPB> ref{a.rx, a.ru, N, bs};
  double xrT[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xrT[c] = ref.X(N - 1, c);
  // pivot row of the 2x2 factorisation for shared weights (the same for the whole batch, from the constant bank)
  const bool swap_shared = fabs(a.kw.R2[1]) > fabs(a.kw.R2[0]);

  if (!chain) {
    // ------------------------------------------------------------------------------ trailer warp
    TilePtrs p;
    p.k = tK;
    p.s = tS;
    p.lin = tL;
    p.rx = RPB ? a.rx + oN * sx : a.rx;
    p.ru = RPB ? a.ru + oM * su : a.ru;
    for (;;) {
      duo_bar();  // A: command published
      asm volatile("fence.proxy.async;" ::: "memory");
      const int c = cmd[0], cur = cmd[1];
      if (c == DUO_EXIT) break;
      const bool store = flags[lane] != 0;
      p.x = tX[cur];
      p.u = tU[cur];
      if (c == DUO_BACKWARD) {
        double xT[4], dJ, sn;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) xT[cc] = p.x[(N - 1) * sx + cc * 32 + lane];
        if (WPB)
          duo_backward_trailer<WPB, RPB, SG, 2>(m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
        else if (swap_shared)
          duo_backward_trailer<WPB, RPB, SG, 1>(m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
        else
          duo_backward_trailer<WPB, RPB, SG, 0>(m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
        res[lane] = dJ;
        res[32 + lane] = sn;
      } else {
        const double cst = duo_forward_trailer<WPB, RPB, SG>(m, w, N, p, r, hd, lane, store, tX[cur ^ 1], tU[cur ^ 1], tL, xrT);
        res[64 + lane] = cst;
      }
      // what this warp stored with ordinary stores is the source of the next pass's bulk copies
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      duo_bar();  // B: pass complete, results published
    }
    return;
  }

  // ---------------------------------------------------------------------------------- chain warp
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
      rk4_step_lin(m, x, u0, u1, xn, L);
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
    // the trailer's bulk copies read what this warp has just written with ordinary stores
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
    // ---- backward pass
    flags[lane] = run ? 1 : 0;
    if (lane == 0) {
      cmd[0] = DUO_BACKWARD;
      cmd[1] = cur;
    }
    duo_bar();  // A
    if (WPB)
      duo_backward_chain<WPB, RPB, SG, 2>(m, w, N, r, hd, lane);
    else if (swap_shared)
      duo_backward_chain<WPB, RPB, SG, 1>(m, w, N, r, hd, lane);
    else
      duo_backward_chain<WPB, RPB, SG, 0>(m, w, N, r, hd, lane);
    duo_bar();  // B
    if (run) {
      dJ = res[lane];
      sn = res[32 + lane];
      if (a.h_sn) a.h_sn[int64_t(it) * B + b] = sn;
    }
    // ---- Armijo line search
    bool need = run, ok = false;
    double gamma = a.o.gamma_0, cn = 0.0;
    int tries = 0;
    for (int i = 0; i < a.o.max_line_search && __any_sync(FULL, need); ++i) {
      flags[lane] = need ? 1 : 0;
      if (lane == 0) {
        cmd[0] = DUO_FORWARD;
        cmd[1] = cur;
      }
      duo_bar();  // A
      duo_forward_chain<RPB, SG>(m, N, r, hd, lane, gamma);
      duo_bar();  // B
      const double c = res[64 + lane];
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
    cur ^= 1;
    ++done;
  }
  if (lane == 0) cmd[0] = DUO_EXIT;
  duo_bar();  // A: releases the trailer
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
