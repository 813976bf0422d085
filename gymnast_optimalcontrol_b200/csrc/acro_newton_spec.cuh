// acro_newton_spec.cuh - k_newton_duo plus SPECULATIVE, PARALLEL Armijo candidates (tg:344-369).
//
// In the back-tracking regime (the reference's default gamma_0 = 1) a tile of 32 problems needs a forward pass for every
// candidate step size ANY of its problems still has to try: the pass count of a tile is the maximum over its lanes
// (about 14 at gamma_0 = 1 on config 2, against a mean of 3.2 per problem), and the passes of k_newton_duo run one
// after the other on two of the four sub-partitions of an SM.  The candidates gamma_0 beta^j of one line search are
// independent of each other, so this kernel evaluates up to eight of them AT ONCE:
//
//   block = 8 warps = one tile.  Backward pass and a lone candidate: warps 0 / 1 are the chain / trailer warps of
//   k_newton_duo (same device functions), the others wait at the pass barrier.
//   Speculative round: warp c rolls out candidate j0 + c (u+ = u + K (x+ - x) + gamma_c sigma, x+ <- RK4, tg:218-229),
//   accumulates its cost (tg:231-252) and stores its trajectory into candidate buffer c; all warps read the operands
//   (x, u, K, sigma, reference) from the ONE TMA-fed ring (full barriers as before, one "empty" mbarrier per stage that
//   every reader arrives on; warp 0 refills a stage once all readers have left it).
//   Selection: warp 0 walks the candidates in the reference's sequential order (gamma *= beta, strict <, NaN rejects):
//   every lane takes its FIRST accepted candidate, exactly what the sequential loop would have picked.
//   Commit: the accepted candidate of every lane is gathered from its buffer, linearised (dynamics.py:217-226,
//   tg:161-164; no recurrence, so the eight warps take every eighth time step) and written where the next backward pass
//   expects the iterate.
//
// How many candidates a round evaluates follows the previous iteration of the tile (`speculate` = 0), or is fixed.
#pragma once
#include "acro_newton_duo.cuh"

namespace acro {

#define ACRO_SPEC_W 8  // warps per block = candidates per round at most

enum { SPEC_EXIT = 0, SPEC_BACKWARD = 1, SPEC_FORWARD_DUO = 2, SPEC_FORWARD_SPEC = 3, SPEC_COMMIT = 4 };

// ACRO_SPEC_TIMING (profiles/spec_timing.py): block 0 logs clock64 at the pass boundaries of warps 0 and 1 into the buffer
// passed as params_b: records of 4 x int64 {warp * 16 + command, after barrier A, pass function returned, after barrier B}
#ifdef ACRO_SPEC_TIMING
#define SPEC_T_DECL long long spec_t_[3]; int spec_t_n_ = 0; (void)spec_t_n_;
#define SPEC_T(i) spec_t_[i] = clock64()
#define SPEC_T_FLUSH(warp_, cmd_)                                                                             \
  if (blockIdx.x == 0 && lane == 0 && a.pb && spec_t_n_ < 2000) {                                             \
    long long* tb_ = reinterpret_cast<long long*>(const_cast<double*>(a.pb)) + ((warp_) * 2000 + spec_t_n_) * 4; \
    tb_[0] = (warp_) * 16 + (cmd_);                                                                           \
    tb_[1] = spec_t_[0];                                                                                      \
    tb_[2] = spec_t_[1];                                                                                      \
    tb_[3] = spec_t_[2];                                                                                      \
    ++spec_t_n_;                                                                                              \
  }
#else
#define SPEC_T_DECL
#define SPEC_T(i)
#define SPEC_T_FLUSH(warp_, cmd_)
#endif

__device__ __forceinline__ void spec_bar() {
  __syncwarp();
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double ldcg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// shared-memory map: the duo map + candidate results and the empty barriers of the operand ring
template <bool RPB, int SG>
struct SpecSmem {
  static constexpr uint32_t ring = 0;
  static constexpr uint32_t hand = ACRO_RING_D * stage_bytes<RPB, SG>();
  static constexpr uint32_t res = hand + ACRO_DUO_R * ACRO_DUO_SLOT_BYTES;  // rows of 32 doubles: dJ, sn, cost (duo path)
  static constexpr uint32_t cand = res + 3 * 256;                           // cost of candidate c: row c
  static constexpr uint32_t gbase = cand + ACRO_SPEC_W * 256;               // per lane: the next step size to try
  static constexpr uint32_t flags = gbase + 256;                            // 32 ints: store / need flags
  static constexpr uint32_t sel = flags + 128;                              // 32 ints: accepted candidate buffer or -1
  static constexpr uint32_t cmd = sel + 128;                                // 4 ints + the zero word (+ padding)
  static constexpr uint32_t bars = cmd + 32;  // ring full[D], hand full[R], hand empty[R], ring empty[D]
  static constexpr uint32_t total = bars + (2 * ACRO_RING_D + 2 * ACRO_DUO_R) * 8;
};

// ---------------------------------------------------------------------------------------------------------
// one speculative candidate: closed-loop rollout with step size gamma (per lane), its cost, its trajectory
// ---------------------------------------------------------------------------------------------------------
template <int SG>
struct SpecIn {
  double x[4], u[2], k[8], s[2], xr[4], ur[2];
  template <bool RPB>
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
    lds_ref<RPB, SG>(stage, st, lane, xr, ur);
  }
};

template <bool WPB, bool RPB, int SG>
__device__ __forceinline__ double spec_forward(const Model& m, const WV<WPB>& w, int N, const TilePtrs& p, Ring& r,
                                               uint32_t empty_bars, bool producer, int lane, double gamma, bool store,
                                               double* __restrict__ Xc, double* __restrict__ Uc, const double xrT[4],
                                               int zmask) {
  constexpr unsigned FULL = 0xffffffffu;
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
  if (producer)
    for (int k = 0; k < ACRO_RING_D && k < n_stages; ++k) ring_fill<RPB, true, SG>(r, p, k, k * SG, min(SG, steps - k * SG));
  double xp[4], xo[4] = {0.0, 0.0, 0.0, 0.0}, uo[2] = {0.0, 0.0};
  SpecIn<SG> in;
  mbar_wait_call(r.bars + ring_slot(r.base) * 8, ring_parity(r.base));
  in.template load<RPB>(r.data + ring_slot(r.base) * stage_bytes<RPB, SG>(), 0, lane);
#pragma unroll
  for (int c = 0; c < 4; ++c) xp[c] = in.x[c];  // x+_0 = x_0
  bool bad = false;
  TrigCarry tc = trig_carry_at(m, xp[0], xp[1]);
  const RotC rotc = RotC::held();
  double cost = 0.0;
  double* po_x = Xc + lane;
  double* po_u = Uc + lane;
  // No mbarrier phase check is inlined in this loop (ptxas would put a YIELD at its head, see mbar_wait_call): warp 0
  // refills the slot of stage k-1 at a fixed step of stage k, behind a blocking wait for the other readers - they run
  // the same instruction stream and are at most a few steps apart.
  for (int k = 0; k < n_stages; ++k) {
    const int cnt = min(SG, steps - k * SG);
    const uint32_t g = r.base + k;
    const uint32_t stage = r.data + ring_slot(g) * stage_bytes<RPB, SG>();
    const uint32_t nstage = r.data + ring_slot(g + 1) * stage_bytes<RPB, SG>();
    const uint32_t nbar = r.bars + ring_slot(g + 1) * 8, npar = ring_parity(g + 1);
    const int fill_s = min(4, cnt - 1);
    for (int s = 0; s < cnt; ++s) {
      const bool cross = (s + 1 == cnt) && (k + 1 < n_stages);
      const bool left = (s == 0) && (k >= 1);  // this warp has read the last operand of stage k-1
      const bool fill = producer && (s == fill_s) && (k >= 1) && (k - 1 + ACRO_RING_D < n_stages);
      if (__any_sync(FULL, cross || bad || left || fill)) {
        if (bad) {  // the last step left the range of the incremental sincos: redo it with a full sincos per stage
          const Vec4 o = rk4_step_redo(m, xo[0], xo[1], xo[2], xo[3], uo[0], uo[1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) xp[i] = o.v[i];
          tc = trig_carry_at(m, xp[0], xp[1]);
        }
        __syncwarp();
        if (left) mbar_arrive_lane0(empty_bars + ring_slot(g - 1) * 8, lane);
        if (fill) {
          // every reader has left stage k-1 (wait for the stragglers): refill its slot with stage k-1+D
          mbar_wait_call(empty_bars + ring_slot(g - 1) * 8, uint32_t((k - 1) / ACRO_RING_D) & 1u);
          const int kk = k - 1 + ACRO_RING_D;
          ring_fill<RPB, true, SG>(r, p, kk, kk * SG, min(SG, steps - kk * SG));
        }
        if (cross) mbar_wait_call(nbar, npar);
      }
      const uint32_t nsrc = cross ? nstage : stage;
      const int ns = (s + 1 < cnt) ? s + 1 : (cross ? 0 : s);
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
#pragma unroll
      for (int c = 0; c < 4; ++c) xo[c] = xp[c];
      uo[0] = up[0];
      uo[1] = up[1];
      double xn[4];
      bad = rk4_step_rot<false>(m, rotc, xp, up[0], up[1], xn, tc, [&]() {
        // operands of the next step (and, in warp 0, whether the stage to refill has been left by every reader),
        // issued while the FP64 pipe is busy
        in.template load<RPB>(nsrc, ns, lane);
        int acc = __double2loint(in.x[0]) | __double2loint(in.x[1]) | __double2loint(in.x[2]) | __double2loint(in.x[3]) |
                  __double2loint(in.u[0]) | __double2loint(in.u[1]) | __double2loint(in.s[0]) | __double2loint(in.s[1]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc |= __double2loint(in.k[e]);
        return acc & zmask;
      });
#pragma unroll
      for (int c = 0; c < 4; ++c) xp[c] = xn[c];
      po_x += kSX;
      po_u += kSU;
    }
  }
  if (__any_sync(FULL, bad)) {
    if (bad) {
      const Vec4 o = rk4_step_redo(m, xo[0], xo[1], xo[2], xo[3], uo[0], uo[1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) xp[i] = o.v[i];
    }
    __syncwarp();
  }
  // (the last stages are never refilled, nobody waits for their empty barriers: no arrival needed)
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
// commit: gather every lane's accepted candidate, linearise about it, write the new iterate
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void spec_commit(const Model& m, int N, int warp, int lane, bool commit,
                                            const double* __restrict__ xsrc, const double* __restrict__ usrc,
                                            double* __restrict__ Xo, double* __restrict__ Uo, double* __restrict__ Lo) {
  // xsrc / usrc: per lane, time step 0 of the trajectory to commit (its accepted candidate buffer; lanes with nothing to
  // commit point at valid data and never store).  The gathers come from DRAM (the candidate buffers were just written
  // and are several times the L2): every warp keeps the operands of its next ACRO_COMMIT_AHEAD steps in flight.
  constexpr int AH = 3;
  const int steps = N - 1;
  double x[AH][4], u[AH][2];
  auto fetch = [&](int slot, int t) {
    if (t < steps) {
#pragma unroll
      for (int c = 0; c < 4; ++c) x[slot][c] = ldcg(xsrc + int64_t(t) * kSX + c * 32);
#pragma unroll
      for (int c = 0; c < 2; ++c) u[slot][c] = ldcg(usrc + int64_t(t) * kSU + c * 32);
    }
  };
#pragma unroll
  for (int a = 0; a < AH; ++a) fetch(a, warp + a * ACRO_SPEC_W);
  for (int t0 = warp; t0 < steps; t0 += AH * ACRO_SPEC_W) {
#pragma unroll
    for (int a = 0; a < AH; ++a) {
      const int t = t0 + a * ACRO_SPEC_W;
      if (t < steps) {
        const double xs[4] = {x[a][0], x[a][1], x[a][2], x[a][3]};
        const double u0 = u[a][0], u1 = u[a][1];
        fetch(a, t + AH * ACRO_SPEC_W);  // the slot is free again: its next occupant
        const LinD L = linearize_d(m, xs, u0, u1);
        if (commit) {
          double* px = Xo + int64_t(t) * kSX + lane;
          double* pu = Uo + int64_t(t) * kSU + lane;
          double* pl = Lo + int64_t(t) * kSL + lane;
#pragma unroll
          for (int c = 0; c < 4; ++c) px[c * 32] = xs[c];
          pu[0] = u0;
          pu[32] = u1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pl[j * 32] = L.a[0][j];
            pl[(4 + j) * 32] = L.a[1][j];
          }
          pl[8 * 32] = L.b[0];
          pl[9 * 32] = L.b[1];
        }
      }
    }
  }
  if (warp == steps % ACRO_SPEC_W && commit) {  // terminal state
#pragma unroll
    for (int c = 0; c < 4; ++c) Xo[int64_t(steps) * kSX + c * 32 + lane] = ldcg(xsrc + int64_t(steps) * kSX + c * 32);
  }
}

// ---------------------------------------------------------------------------------------------------------
// kernel: eight warps per block, block = tile of 32 problems
// ---------------------------------------------------------------------------------------------------------
template <bool WPB, bool RPB, int SG>
__global__ void __launch_bounds__(ACRO_SPEC_W * 32, 1) k_newton_spec(const __grid_constant__ NewtonArgs a) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  using SM = SpecSmem<RPB, SG>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t B = a.B, tile = blockIdx.x, b0 = tile * 32LL, b = b0 + lane;
  const bool valid = b < B;
  const int64_t bs = valid ? b : B - 1;  // padding lanes shadow the last problem and never write
  const int N = a.N;
  const int steps = N - 1, n_stages = (steps + SG - 1) / SG;
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
  const uint32_t empty_bars = hd.empty + ACRO_DUO_R * 8;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < 2 * ACRO_RING_D + 2 * ACRO_DUO_R; ++s) mbar_init(r.bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    reinterpret_cast<volatile int*>(ring_smem + SM::cmd)[4] = 0;
  }
  __syncthreads();
  hd.zmask = reinterpret_cast<volatile int*>(ring_smem + SM::cmd)[4];
  volatile double* const res = reinterpret_cast<volatile double*>(ring_smem + SM::res);
  volatile double* const cand = reinterpret_cast<volatile double*>(ring_smem + SM::cand);
  volatile double* const gbase = reinterpret_cast<volatile double*>(ring_smem + SM::gbase);
  volatile int* const flags = reinterpret_cast<volatile int*>(ring_smem + SM::flags);
  volatile int* const sel = reinterpret_cast<volatile int*>(ring_smem + SM::sel);
  volatile int* const cmd = reinterpret_cast<volatile int*>(ring_smem + SM::cmd);

  constexpr int64_t sx = 4 * 32, su = 2 * 32, sk = 8 * 32, ss = 2 * 32, sl = 10 * 32;
  const int64_t oN = tile * N, oM = tile * (N - 1);
  const int64_t tiles = (B + 31) / 32;
  double* const tX[2] = {a.X + oN * sx, a.Xw + oN * sx};
  double* const tU[2] = {a.U + oM * su, a.Uw + oM * su};
  double* const tK = a.K + oM * sk;
  double* const tS = a.S + oM * ss;
  double* const tL = a.lin + oM * sl;
  // candidate buffers: a.spec_ws = Xc[8][tiles][N][4][32] then Uc[8][tiles][N-1][2][32]
  double* const cX0 = a.spec_ws + oN * sx;
  double* const cU0 = a.spec_ws + int64_t(ACRO_SPEC_W) * tiles * N * sx + oM * su;
  const int64_t cXs = tiles * N * sx, cUs = tiles * (N - 1) * su;  // stride between candidate buffers
  const RefV<RPB> ref{a.rx, a.ru, N, bs};
  double xrT[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xrT[c] = ref.X(N - 1, c);
  const bool swap_shared = fabs(a.kw.R2[1]) > fabs(a.kw.R2[0]);
  TilePtrs p;
  p.k = tK;
  p.s = tS;
  p.lin = tL;
  p.rx = RPB ? a.rx + oN * sx : a.rx;
  p.ru = RPB ? a.ru + oM * su : a.ru;

  if (warp != 0) {
    // ------------------------------------------------------------------------------ workers (warp 1 = the duo trailer)
    SPEC_T_DECL
    for (;;) {
      spec_bar();  // A: command published
      SPEC_T(0);
      asm volatile("fence.proxy.async;" ::: "memory");
      const int c = cmd[0], cur = cmd[1], J = cmd[2];
      if (c == SPEC_EXIT) break;
      p.x = tX[cur];
      p.u = tU[cur];
      if (c == SPEC_BACKWARD) {
        if (warp == 1) {
          const bool store = flags[lane] != 0;
          double xT[4], dJ, sn;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) xT[cc] = p.x[(N - 1) * sx + cc * 32 + lane];
          if (WPB)
            duo_backward_trailer<WPB, RPB, SG, 2, true>(a.m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
          else if (swap_shared)
            duo_backward_trailer<WPB, RPB, SG, 1, true>(a.m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
          else
            duo_backward_trailer<WPB, RPB, SG, 0, true>(a.m, w, N, p, r, hd, lane, store, tK, tS, xT, xrT, dJ, sn);
          res[lane] = dJ;
          res[32 + lane] = sn;
        } else {
          r.base += n_stages;
        }
      } else if (c == SPEC_FORWARD_DUO) {
        if (warp == 1) {
          const bool store = flags[lane] != 0;
          const double cst = duo_forward_trailer<WPB, RPB, SG, true>(a.m, w, N, p, r, hd, lane, store, tX[cur ^ 1], tU[cur ^ 1], tL, xrT);
          res[64 + lane] = cst;
        } else {
          r.base += n_stages;
        }
      } else if (c == SPEC_FORWARD_SPEC) {
        if (warp < J) {
          const bool store = flags[lane] != 0;
          double gamma = gbase[lane];
          for (int i = 0; i < warp; ++i) gamma = __dmul_rn(gamma, a.o.beta);  // tg:365, sequential products
          const double cst = spec_forward<WPB, RPB, SG>(a.m, w, N, p, r, empty_bars, false, lane, gamma, store, cX0 + warp * cXs,
                                                        cU0 + warp * cUs, xrT, hd.zmask);
          cand[warp * 32 + lane] = cst;
        } else {
          r.base += n_stages;
        }
      } else {  // SPEC_COMMIT
        const int sj = sel[lane];
        const bool commit = sj >= 0;
        const double* xs = commit ? cX0 + sj * cXs + lane : tX[cur] + lane;
        const double* us = commit ? cU0 + sj * cUs + lane : tU[cur] + lane;
        spec_commit(a.m, N, warp, lane, commit, xs, us, tX[cur ^ 1], tU[cur ^ 1], tL);
      }
      // what this warp stored with ordinary stores is read by other warps and by the next pass's bulk copies
      SPEC_T(1);
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      spec_bar();  // B: pass complete, results published
      SPEC_T(2);
      if (warp == 1) { SPEC_T_FLUSH(1, c) }
    }
    return;
  }

  // ---------------------------------------------------------------------------------- warp 0: chain + control
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
  int pred = 1;  // candidates the previous line search of this tile needed (maximum over its lanes)

  auto publish = [&](int c, int J) {
    if (lane == 0) {
      cmd[0] = c;
      cmd[1] = cur;
      cmd[2] = J;
    }
  };
  SPEC_T_DECL
  auto after_pass = [&]() {
    SPEC_T(1);
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    spec_bar();  // B
    SPEC_T(2);
    SPEC_T_FLUSH(0, cmd[0])
  };

  while (__any_sync(FULL, run) && (a.o.chunk_iters <= 0 || done < a.o.chunk_iters)) {
    p.x = tX[cur];
    p.u = tU[cur];
    // ---- backward pass (warps 0, 1)
    flags[lane] = run ? 1 : 0;
    publish(SPEC_BACKWARD, 0);
    spec_bar();  // A
    SPEC_T(0);
    if (WPB)
      duo_backward_chain<WPB, RPB, SG, 2, true>(a.m, w, N, r, hd, lane);
    else if (swap_shared)
      duo_backward_chain<WPB, RPB, SG, 1, true>(a.m, w, N, r, hd, lane);
    else
      duo_backward_chain<WPB, RPB, SG, 0, true>(a.m, w, N, r, hd, lane);
    after_pass();
    if (run) {
      dJ = res[lane];
      sn = res[32 + lane];
      if (a.h_sn) a.h_sn[int64_t(it) * B + b] = sn;
    }
    // ---- Armijo line search: rounds of J candidates
    bool need = run, ok = false, commit = false;
    double gamma = a.o.gamma_0, cn = 0.0;
    int tries = 0, tried = 0, acc_buf = -1;  // tried: candidates evaluated by the tile so far
    while (tried < a.o.max_line_search && __any_sync(FULL, need)) {
      int J = a.o.speculate > 0 ? a.o.speculate : (pred - tried > 0 ? pred - tried : 4);
      J = min(min(J, ACRO_SPEC_W), a.o.max_line_search - tried);
      flags[lane] = need ? 1 : 0;
      if (J == 1) {
        // a lone candidate: the chain / trailer pair writes it (and its linearisation) straight to the other buffer
        publish(SPEC_FORWARD_DUO, 1);
        spec_bar();  // A
        SPEC_T(0);
        duo_forward_chain<RPB, SG, true>(a.m, N, r, hd, lane, gamma);
        after_pass();
        const double c = res[64 + lane];
        if (need) {
          ++tries;
          // accept iff cost_new < cost_k + c*gamma*delta_J  (strict, NaN rejects)   tg:361
          const double thr = __dadd_rn(cost_k, __dmul_rn(__dmul_rn(a.o.c, gamma), dJ));
          if (c < thr) {
            ok = true;
            need = false;
            cn = c;
            commit = false;  // already where it belongs
          } else {
            gamma = __dmul_rn(gamma, a.o.beta);  // tg:365
          }
        }
      } else {
        gbase[lane] = gamma;
        if (lane == 0) {  // the stage-release barriers count the readers of this round
#pragma unroll
          for (int s = 0; s < ACRO_RING_D; ++s) {
            mbar_inval(empty_bars + s * 8);
            mbar_init(empty_bars + s * 8, uint32_t(J));
          }
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        publish(SPEC_FORWARD_SPEC, J);
        spec_bar();  // A
        SPEC_T(0);
        const double c0 = spec_forward<WPB, RPB, SG>(a.m, w, N, p, r, empty_bars, true, lane, gamma, need, cX0, cU0, xrT, hd.zmask);
        cand[lane] = c0;
        after_pass();
        for (int j = 0; j < J; ++j) {  // the reference's order: first accepted wins
          const double c = cand[j * 32 + lane];
          if (need) {
            ++tries;
            const double thr = __dadd_rn(cost_k, __dmul_rn(__dmul_rn(a.o.c, gamma), dJ));
            if (c < thr) {
              ok = true;
              need = false;
              cn = c;
              commit = true;
              acc_buf = j;
            } else {
              gamma = __dmul_rn(gamma, a.o.beta);
            }
          }
        }
      }
      tried += J;
    }
    // ---- commit the candidates accepted in speculative rounds
    if (__any_sync(FULL, commit)) {
      sel[lane] = commit ? acc_buf : -1;
      publish(SPEC_COMMIT, 0);
      spec_bar();  // A
      SPEC_T(0);
      const double* xs = commit ? cX0 + acc_buf * cXs + lane : tX[cur] + lane;
      const double* us = commit ? cU0 + acc_buf * cUs + lane : tU[cur] + lane;
      spec_commit(a.m, N, 0, lane, commit, xs, us, tX[cur ^ 1], tU[cur ^ 1], tL);
      after_pass();
    }
    if (run) {
      pred = tries;
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
    } else {
      pred = 0;
    }
    pred = max(1, __reduce_max_sync(FULL, pred));
    cur ^= 1;
    ++done;
  }
  publish(SPEC_EXIT, 0);
  spec_bar();  // A: releases the workers
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
