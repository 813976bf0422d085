/*
 * acro_abi.h - C ABI of libacro_b200.so: the batched acrobot optimal-control hot path on
 * NVIDIA B200 (sm_100a), FP64 throughout.
 *
 * The reference (francescoolivieri/Gymnast_OptimalControl) has no FFI layer: its boundary
 * is a set of module-level Python functions.  Each entry point below names the reference
 * function it replaces (file:line, relative to the reference root).  A maintainer binds
 * them with ctypes; gymnast_optimalcontrol_b200/_abi.py is that binding.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative ACRO_E_* code otherwise, never throws,
 *    and is asynchronous on the given stream (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream).  There is no global mutable state apart from a thread-local
 *    error string (acro_last_error_string) and the launch counter; no environment variable is read.
 *  - All array pointers are caller-owned DEVICE pointers to FP64 (or int32 where stated)
 *    unless a parameter is documented as HOST.  The library allocates nothing.
 *  - Point batches are structure-of-arrays with the problem index fastest:
 *        state batch      x[c][b]            c in 0..3            -> x[c*B + b]
 *        input batch      u[c][b]            c in 0..1
 *        per-problem scalars, int32 flags    v[b]
 *    TIME-INDEXED batch arrays (trajectories, gains, ...) are TILED structure-of-arrays, tile-major:
 *        A[tile][t][c][lane],   tile = b / 32, lane = b % 32,   T time steps, C components
 *        element (t, c, b)  ->  A[((tile*T + t)*C + c)*32 + lane]
 *    so the whole trajectory of a warp's 32 problems is one contiguous block: the C rows of a time step
 *    are C*256 consecutive, 128-byte aligned bytes (component c at the constant offset c*256) and
 *    consecutive time steps follow each other: coalesced, vectorisable, streamed with one pointer per
 *    array, and any run of consecutive steps is movable by a single bulk copy.  Buffers hold
 *    32*ceil(B/32)*T*C doubles; padding lanes are never read for results.
 *    Below, a tiled array with T time steps and C components is written  name {T x C}:
 *        state trajectory X {N x 4}     input trajectory U {N-1 x 2}
 *        gains            K {N-1 x 8}   (K_t is 2x4 row-major, component i*4+j)
 *        feed-forward     S {N-1 x 2}   (sigma_t)
 *    and point batches keep the bracket notation  x [4][B].
 *    acro_pack_soa / acro_unpack_soa convert from / to batch-major (B, T, C) arrays.
 *  - "Shared" reference data (one trajectory for the whole batch) is plain row-major
 *    [t][c], i.e. exactly the NumPy arrays of the reference: x_ref (N,4), u_ref (N-1,2),
 *    K_reg (N-1,2,4).
 */
#ifndef ACRO_ABI_H
#define ACRO_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACRO_ABI_VERSION 2

/* error codes */
#define ACRO_OK 0
#define ACRO_E_INVALID (-1) /* bad argument (null pointer, non-positive size, ...) */
#define ACRO_E_CUDA (-2)    /* a CUDA runtime call failed; see acro_last_error_string() */
#define ACRO_E_NODEVICE (-3)

/* per-problem solver status (int32), mirrors the exits of newton_Algorithm
 * (trajectory_generation.py:329-396) */
#define ACRO_RUNNING 0            /* not finished (resume with another call) */
#define ACRO_CONVERGED 1          /* max|sigma| < tol after an accepted step (tg:394-396) */
#define ACRO_MAX_ITERS 2          /* loop ran max_iters times (tg:329) */
#define ACRO_LINE_SEARCH_FAILED 3 /* 20 rejected candidates (tg:367-369); iterate kept */

/* Physical parameters: dynamics.py:15-61 (params_1/2/3), dt: dynamics.py:173.
 * actuated_tau1 = 0 reproduces the plant of dynamics.py:205 (tau_1 forced to 0);
 * 1 gives the fully-actuated model of fully_actuated_ref_gen.py:20-73. */
typedef struct AcroParams {
  double m1, m2, l1, lc1, l2, lc2, I1, I2, g, f1, f2;
  double dt;
  int32_t actuated_tau1;
  int32_t reserved;
} AcroParams;

/* Cost weights.  Q, R, QT are HOST arrays (row-major 4x4, 2x2, 4x4, symmetric), the
 * module-level constants of trajectory_generation.py:16-18 / trajectory_tracking.py:38-39,
 * 173-175.  If the *_b device pointers are non-NULL they override them per problem:
 * Q_b[e][b] (e = 0..15), R_b[e][b] (e = 0..3), QT_b[e][b]. */
typedef struct AcroWeights {
  double Q[16];
  double R[4];
  double QT[16];
  const double* Q_b;
  const double* R_b;
  const double* QT_b;
} AcroWeights;

/* Reference trajectory handed to the optimiser / tracker.
 * per_problem = 0: x (N,4) and u (N-1,2) row-major, shared by the batch.
 * per_problem = 1: x {N x 4}, u {N-1 x 2}. */
typedef struct AcroRef {
  const double* x;
  const double* u;
  int32_t per_problem;
  int32_t reserved;
} AcroRef;

/* Arguments of newton_Algorithm (trajectory_generation.py:298) */
typedef struct AcroNewtonOpts {
  int32_t max_iters;    /* total cap on iterations per problem */
  int32_t chunk_iters;  /* at most this many iterations in THIS call (0 = no limit) */
  int32_t max_line_search; /* 20 in the reference (tg:345) */
  int32_t init;         /* 1: start from u = 0, roll out from x0, compute the initial cost
                           (tg:311-319); 2: the same but keep the caller's U as the initial inputs
                           (warm start); 0: resume - X, U, lin_ws, cost, delta_J, sigma_norm,
                           gamma_acc, iters, status must be what the previous call left */
  double tol;           /* tg:394 */
  double beta;          /* tg:365 */
  double c;             /* tg:361 */
  double gamma_0;       /* tg:344 */
  /* Kernel selection (all 0 = automatic, by batch size and SM count of the current device).  Every variant
   * computes the same iteration; they differ in how a tile of 32 problems is mapped to warps, and agree to
   * rounding (1e-12) but not bit for bit (different sin/cos evaluation order), so pin `kernel`/`stage_steps`
   * when results must not depend on the batch size of the call (sharding across GPUs). */
  int32_t kernel;       /* ACRO_NEWTON_AUTO / _DUO / _RING / _THREAD / _SPEC */
  int32_t stage_steps;  /* time steps per TMA stage of the ring kernels: 0 auto, or 2, 4, 8, 16 */
  int32_t recompute_lin; /* large-batch ring kernel: 0 auto, 1 backward pass recomputes the linearisation
                            (304 B per problem-step-iteration), 2 it streams the one the forward pass stored (464 B) */
  int32_t speculate;    /* Armijo candidates evaluated in parallel per round by ACRO_NEWTON_SPEC: 0 auto
                            (from the previous iteration's number of tries), 1..8 fixed */
  double* spec_ws;      /* DEVICE workspace of ACRO_NEWTON_SPEC (candidate trajectories), 128-byte aligned,
                            acro_newton_spec_ws_doubles(B, N) doubles, or NULL.  With kernel = ACRO_NEWTON_AUTO the
                            speculative kernel is chosen when this workspace is given, the batch is at most one tile per
                            SM and gamma_0 >= 0.5 (a line search that starts with a large step back-tracks) */
} AcroNewtonOpts;
#define ACRO_NEWTON_AUTO 0
#define ACRO_NEWTON_DUO 1    /* two warps per tile (recurrence + trailer), at most two tiles per SM */
#define ACRO_NEWTON_RING 2   /* one warp per tile, TMA-fed ring */
#define ACRO_NEWTON_THREAD 3 /* one thread per problem, register prefetch (no alignment requirement) */
#define ACRO_NEWTON_SPEC 4   /* eight warps per tile: like DUO, plus speculative parallel Armijo candidates */

const char* acro_version(void);
const char* acro_last_error_string(void);
/* Number of kernels this library has launched in the calling process (for bench.py). */
int64_t acro_launch_count(void);

/* ---- D1-D3: dynamics.py ------------------------------------------------------------ */
/* continuous_dynamics(xx, uu)  dynamics.py:197-213.  x [4][B], u [2][B] -> xdot [4][B] */
int acro_continuous_dynamics(const AcroParams* p, int64_t B, const double* x, const double* u,
                             double* xdot, void* stream);
/* dynamics(xx, uu): one RK4 step  dynamics.py:177-195.  -> xnext [4][B] */
int acro_rk4_step(const AcroParams* p, int64_t B, const double* x, const double* u, double* xnext,
                  void* stream);
/* Calculate_A_B_matrixes(x_t, u_t)  dynamics.py:217-226  (discrete = 0), optionally followed
 * by discretize_linearization  trajectory_generation.py:161-164  (discrete = 1).
 * A [16][B] (4x4 row-major), Bm [8][B] (4x2 row-major). */
int acro_linearize(const AcroParams* p, int64_t B, const double* x, const double* u, double* A,
                   double* Bm, int discrete, void* stream);

/* ---- reference builders ------------------------------------------------ */
/* compute_equilibrium(u_target, theta_guess)  trajectory_generation.py:22-39 for a batch: solves the gravity balance
 * G(theta1, theta2) = u_target by Newton's method from theta_guess (the reference uses SciPy's hybr on the same two
 * equations).  u_target [2][B], theta_guess [2][B] -> theta [2][B] (x_e = [theta1, theta2, 0, 0], u_e = u_target),
 * n_iter [B] int32: iterations used, or -max_iter if max|residual| >= tol after max_iter iterations (the reference
 * raises RuntimeError there).  params_b: per-problem physical parameters [11][B] or NULL. */
int acro_equilibrium(const AcroParams* p, const double* params_b, int64_t B, const double* u_target,
                     const double* theta_guess, double tol, int max_iter, double* theta, int32_t* n_iter, void* stream);

/* ---- G1-G11: trajectory_generation.py ---------------------------------------------- */
/* simulate_open_loop(x0, u_traj)  tg:74-87.  x0 [4][B], U {N-1 x 2} (NULL = zeros)
 * -> X {N x 4} */
int acro_rollout_open_loop(const AcroParams* p, int64_t B, int N, const double* x0, const double* U,
                           double* X, void* stream);
/* total_cost(x_traj, u_traj, x_ref, u_ref, Q, R, Q_T)  tg:231-252 -> cost [B] */
int acro_total_cost(const AcroWeights* w, int64_t B, int N, const double* X, const double* U,
                    const AcroRef* ref, double* cost, void* stream);
/* compute_costate_trajectory  tg:138-159 -> lam {N x 4}.  (Its result is not used by the
 * reference's Newton loop; exposed for completeness.) */
int acro_costate(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X,
                 const double* U, const AcroRef* ref, double* lam, void* stream);
/* derivatives_Cost(x, x_ref, u, u_ref, Q, R, Q_T, terminal)  tg:89-114 for a batch of points:
 * x, x_ref [4][B], u, u_ref [2][B] -> l [B], grad_x [4][B] = 2Q(x-x_ref) (2Q_T if terminal),
 * grad_u [2][B] = 2R(u-u_ref) (untouched if terminal).  The Hessians are the constants 2Q, 2R, 2Q_T. */
int acro_cost_derivatives(const AcroWeights* w, int64_t B, const double* x, const double* x_ref,
                          const double* u, const double* u_ref, int terminal, double* l, double* grad_x,
                          double* grad_u, void* stream);
/* discretize_linearization(Ac, Bc, dt)  tg:161-164: Ad = I + dt Ac, Bd = dt Bc on [16][B], [8][B]. */
int acro_discretize(int64_t B, const double* Ac, const double* Bc, double dt, double* Ad, double* Bd,
                    void* stream);
/* build_stage_lists(x_traj, u_traj, x_ref, u_ref, lambda_seq)  tg:166-181 (lambda_seq is ignored by the
 * reference, tg:116-129): -> A {N-1 x 16}, Bm {N-1 x 8}, q {N-1 x 4}, r {N-1 x 2}, q_T [4][B].
 * The quadratic blocks are the constants 2Q, 2R, S = 0 and 2Q_T. */
int acro_stage_lists(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X,
                     const double* U, const AcroRef* ref, double* A, double* Bm, double* q, double* r,
                     double* q_T, void* stream);
/* calculate_K_and_sigma(A_list, B_list, Q_list, R_list, S_list, q_list, r_list, Q_T_block, q_T)
 * tg:183-216 on caller-supplied dense lists: A {T x 16}, Bm {T x 8}, Q {T x 16}, R {T x 4},
 * S_cross {T x 8} (NULL = 0), q {T x 4}, r {T x 2}, Q_T [16][B], q_T [4][B]
 * -> K {T x 8}, S {T x 2}, delta_J [B]. */
int acro_riccati_lists(int64_t B, int T, const double* A, const double* Bm, const double* Q, const double* R,
                       const double* S_cross, const double* q, const double* r, const double* Q_T,
                       const double* q_T, double* K, double* S, double* delta_J, void* stream);
/* build_stage_lists + calculate_K_and_sigma fused  tg:166-216:
 * -> K {N-1 x 8}, S {N-1 x 2}, delta_J [B] (expected_reduction), sigma_norm [B] = max|sigma| */
int acro_riccati_affine(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X,
                        const double* U, const AcroRef* ref, double* K, double* S, double* delta_J,
                        double* sigma_norm, void* stream);
/* forward_closed_loop_update + total_cost for a set of G step sizes  tg:218-252.
 * One rollout per (problem b, candidate g).  gammas: [G][B] if gammas_per_problem else [G].
 * cost out [G][B].  Xn [G]{N x 4} and Un [G]{N-1 x 2} may be NULL (cost only). */
int acro_closed_loop_rollout_cost(const AcroParams* p, const AcroWeights* w, int64_t B, int N,
                                  const double* X, const double* U, const double* K, const double* S,
                                  const AcroRef* ref, int G, const double* gammas,
                                  int gammas_per_problem, double* Xn, double* Un, double* cost,
                                  void* stream);
/* The Armijo test of tg:352-365 on precomputed candidate costs: first g (in order) with
 * cost_cand[g][b] < cost_k[b] + c*gamma[g]*delta_J[b] (strict; NaN rejects) -> accepted [B]
 * (int32, -1 = none). */
int acro_armijo_select(int64_t B, int G, const double* cost_k, const double* delta_J,
                       const double* gammas, int gammas_per_problem, const double* cost_cand,
                       double c, int32_t* accepted, void* stream);
/* newton_Algorithm  tg:298-398, one problem per thread, whole loop on the device.
 * In/out: X {N x 4}, U {N-1 x 2} (current iterate; written by init), cost [B],
 * iters [B] int32, status [B] int32.  x0 [4][B] is read when opts->init.
 * Workspace: Xw, Uw same sizes as X, U; lin_ws {N-1 x 10} holds the discrete linearisation about the
 * current iterate (written by the rollouts, read by the next backward pass; part of the state to keep when
 * resuming).  Out: K, S of the last computed iteration
 * (evaluated on the pre-update trajectory, as the reference returns them), delta_J [B],
 * sigma_norm [B], gamma_acc [B] (last accepted step).
 * Optional history (NULL to skip): hist_cost [(max_iters+1)][B], hist_sigma_norm
 * [max_iters][B], hist_gamma [max_iters][B], hist_ntry [max_iters][B] int32. */
int acro_newton_solve(const AcroParams* p, const AcroWeights* w, const AcroNewtonOpts* opts, int64_t B,
                      int N, const double* x0, const AcroRef* ref, double* X, double* U, double* Xw,
                      double* Uw, double* lin_ws, double* K, double* S, double* cost, double* delta_J,
                      double* sigma_norm, double* gamma_acc, int32_t* iters, int32_t* status,
                      double* hist_cost, double* hist_sigma_norm, double* hist_gamma,
                      int32_t* hist_ntry, void* stream);
/* Size of AcroNewtonOpts.spec_ws in doubles (8 candidate copies of X and U). */
int64_t acro_newton_spec_ws_doubles(int64_t B, int N);
/* Name of the kernel acro_newton_solve would launch for these options and this batch on the current device
 * (e.g. "acro::k_newton_duo<false,false,16>"), for logs and bench.py; launches nothing. */
int acro_newton_describe(const AcroNewtonOpts* opts, int64_t B, int ref_per_problem, int weights_per_problem,
                         int params_per_problem, char* buf, int buf_len);
/* The step-size sweep of plot_armijo_line_search  tg:257-264: P base iterates (X,U,K,S with
 * batch size P) x S_n shared step sizes steps[S_n] -> cost [S_n][P]. */
int acro_stepsize_sweep(const AcroParams* p, const AcroWeights* w, int64_t P, int N, const double* X,
                        const double* U, const double* K, const double* S, const AcroRef* ref, int S_n,
                        const double* steps, double* cost, void* stream);

/* ---- T1-T5: trajectory_tracking.py ------------------------------------------------- */
/* solve_LQR_tracking(x_opt, u_opt)  tt:170-203 with weights w->Q (Q_reg), w->R (R_reg) and
 * P_T = 2 Q_reg.  traj = the trajectory linearised about (AcroRef layout; shared => B = 1
 * problem and K is (N-1,2,4) row-major, per-problem => K {N-1 x 8}). */
int acro_lqr_gains(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const AcroRef* traj,
                   double* K, void* stream);
/* ---- per-problem physical parameters (dynamics.py:15-61 defines three parameter sets,
 * set_params dynamics.py:117-144 selects one for the whole program).  params_b [11][B] (device): rows m1, m2, l1,
 * lc1, l2, lc2, I1, I2, g, f1, f2 per problem; dt and actuated_tau1 from p.  params_b = NULL is the shared-parameter
 * call above.  Everything else as in the function without the suffix. */
int acro_continuous_dynamics_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x,
                                const double* u, double* xdot, void* stream);
int acro_rk4_step_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x, const double* u,
                     double* xnext, void* stream);
int acro_linearize_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x, const double* u,
                      double* A, double* Bm, int discrete, void* stream);
int acro_rollout_open_loop_pp(const AcroParams* p, const double* params_b, int64_t B, int N, const double* x0,
                              const double* U, double* X, void* stream);
/* acro_newton_solve with per-problem physical parameters (one thread per problem; params_b = NULL: acro_newton_solve) */
int acro_newton_solve_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, const AcroNewtonOpts* opts,
                         int64_t B, int N, const double* x0, const AcroRef* ref, double* X, double* U, double* Xw,
                         double* Uw, double* lin_ws, double* K, double* S, double* cost, double* delta_J,
                         double* sigma_norm, double* gamma_acc, int32_t* iters, int32_t* status, double* hist_cost,
                         double* hist_sigma_norm, double* hist_gamma, int32_t* hist_ntry, void* stream);
/* the plant of every problem has its own parameters, the gains are the caller's (e.g. those of the nominal model:
 * tracking under model mismatch) */
int acro_lqr_track_pp(const AcroParams* p, const double* params_b, int64_t B, int N, const AcroRef* traj,
                      const double* K, const double* x0, double* Xt, double* Ut, void* stream);

/* simulate_tracking(x_opt, u_opt, K_reg, x0_perturbed)  tt:206-216.
 * traj/K shared (K (N-1,2,4)) or per-problem (K {N-1 x 8}) following traj->per_problem.
 * x0 [4][B] -> Xt {N x 4}, Ut {N-1 x 2}. */
int acro_lqr_track(const AcroParams* p, int64_t B, int N, const AcroRef* traj, const double* K,
                   const double* x0, double* Xt, double* Ut, void* stream);
/* compute_P_inf(A, B, Q, R)  tt:144-165.  A [16][B], Bm [8][B], weights from w->Q, w->R
 * (or per problem) -> P [16][B], n_iter [B] int32 (max_iter reached => n_iter = -max_iter). */
int acro_p_inf(const AcroWeights* w, int64_t B, const double* A, const double* Bm, int max_iter,
               double tol, double* P, int32_t* n_iter, void* stream);
/* solver_mpc(x0, A_list, B_list, Q, R, Q_T, T_pred)  tt:73-140 restated as the
 * equality-constrained LQ problem it is.  x0 [4][B]; window A_w {T_pred-1 x 16},
 * B_w {T_pred-1 x 8}; terminal weight QT [16][B]; w->Q, w->R stage weights.
 * -> U0 [2][B], X_opt {T_pred x 4}, U_opt {T_pred x 2} (X_opt/U_opt may be NULL),
 * K_ws {T_pred-1 x 8}: the gains of the sweep (out; required when X_opt/U_opt are asked for,
 * else may be NULL). */
int acro_mpc_solve(const AcroWeights* w, int64_t B, int T_pred, const double* x0, const double* A_w,
                   const double* B_w, const double* QT, double* U0, double* X_opt, double* U_opt,
                   double* K_ws, void* stream);
/* solve_mpc_tracking(x0, x_ref, u_ref, T)  tt:8-69: T-1 receding-horizon solves per problem,
 * each a (T_pred-1)-step Riccati sweep over the sliding window of the linearisation about
 * the reference (padded with the linearisation about x_f, u_f), then one plant step.
 * ref shared: the first-move gains are computed once per time step by one sweep each
 * (K0 [(T-1)][8] workspace/out, row-major) and applied to every problem;
 * ref per problem: every problem runs its own sweeps.  lin_ws is a workspace for the compact
 * linearisation: [N-1][10] (shared) or {N-1 x 10} (per problem).
 * x_f HOST [4], u_f HOST [2]  (tt:33-34).
 * QT_inf: DEVICE terminal weight (from acro_p_inf), [16] or, if qt_per_problem, [16][B].
 * -> Xr {T x 4}, Ur {T-1 x 2}; n_solves (HOST int64*, may be NULL) = Riccati sweeps run. */
int acro_mpc_track(const AcroParams* p, const AcroWeights* w, int64_t B, int N, int T, int T_pred,
                   const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                   int qt_per_problem, const double* x0, double* K0, double* lin_ws, double* Xr,
                   double* Ur, int64_t* n_solves, void* stream);
/* the same with physical parameters per problem (params_b [11][B] as in acro_rk4_step_pp; dynamics.py:15-61): every
 * problem linearises its reference, pads its window about (x_f, u_f) and steps its plant with its own model.  The
 * reference must be in the per-problem layout; QT_inf is normally per problem too (acro_linearize_pp at x_f, then
 * acro_p_inf).  params_b NULL: acro_mpc_track with a per-problem reference. */
int acro_mpc_track_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                      int T_pred, const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                      int qt_per_problem, const double* x0, double* lin_ws, double* Xr, double* Ur,
                      int64_t* n_solves, void* stream);

/* ---- measurement helper ------------------------------------------------------------- */
/* solve_mpc_tracking with the input box of trajectory_tracking.py:87-91, 102-104, 112-114 switched on
 * (`test_constraints`): every receding-horizon step solves
 *     min sum_{j<T_pred-1} x_j'Q x_j + u_j'R u_j + x_{T_pred-1}' Q_T x_{T_pred-1},  x_{j+1} = A_j x_j + B_j u_j,
 *     -tau_max <= u_j + u_ref[t+j] <= tau_max
 * exactly (primal active-set method on Riccati sweeps, csrc/acro_mpc_box.cuh) and applies u_ref[t] + u_0 to the plant.
 * R must be diagonal, weights shared.  Reference shared or per problem; lin_ws as in acro_mpc_track;
 * ws: acro_mpc_box_ws_doubles(B, T, T_pred) doubles of scratch (per-problem part, then the tables a shared reference
 * uses: gains of the empty working set, (T-1)(T_pred-1) x 4, and the padded window rows, (T+T_pred-3) x 11).  max_iter <= 0: 6 (T_pred-1) + 20 iterations per step.
 * Out: Xr {T x 4}, Ur {T-1 x 2}; optional n_sweeps [B] (active-set iterations of the problem), n_active [T-1][B]
 * (inputs at a bound in the solution of step t), status [B] (1 if a step hit max_iter, else 0). */
int64_t acro_mpc_box_ws_doubles(int64_t B, int T, int T_pred);
int acro_mpc_track_box(const AcroParams* p, const AcroWeights* w, int64_t B, int N, int T, int T_pred,
                       const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                       int qt_per_problem, const double* x0, double tau_max, int max_iter, double* lin_ws, double* ws,
                       double* Xr, double* Ur, int32_t* n_sweeps, int32_t* n_active, int32_t* status, void* stream);
/* the same with physical parameters per problem (params_b [11][B]; per-problem reference layout, QT_inf normally per
 * problem): own linearisation, own padding about (x_f, u_f), own plant */
int acro_mpc_track_box_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                          int T_pred, const AcroRef* ref, const double* x_f, const double* u_f,
                          const double* QT_inf, int qt_per_problem, const double* x0, double tau_max, int max_iter,
                          double* lin_ws, double* ws, double* Xr, double* Ur, int32_t* n_sweeps, int32_t* n_active,
                          int32_t* status, void* stream);

/* Runs blocks x threads threads, each doing iters x 8 independent dependent-chain DFMAs
 * (16 flops per thread per iteration); out [blocks*threads].  Timed by bench.py to get the
 * achievable FP64 pipe peak of the device it runs on. */
int acro_bench_fp64_peak(int blocks, int threads, int iters, double* out, void* stream);
/* `chains` (1, 2, 4 or 8) independent dependent-DFMA chains of `iters` links per thread, executed only by
 * lanes < active_lanes of every warp; cycles[blocks] (int64) = clock64 ticks of thread 0 of each block.
 * cycles/iters with one chain and one warp per SM sub-partition = DFMA latency. */
int acro_bench_fp64_chain(int blocks, int threads, int iters, int chains, int active_lanes, double* out,
                          long long* cycles, void* stream);

/* ---- layout helpers ------------------------------------------------------------------ */
/* batch-major src (B, T, C) row-major  ->  tiled dst [tile][t][c][lane] (padding lanes zero-filled) */
int acro_pack_soa(int64_t B, int T, int C, const double* src, double* dst, void* stream);
/* tiled src [tile][t][c][lane]  ->  batch-major dst (B, T, C) row-major */
int acro_unpack_soa(int64_t B, int T, int C, const double* src, double* dst, void* stream);
/* plain matrix transpose src (rows, cols) -> dst (cols, rows): state batches (B, 4) <-> [4][B] */
int acro_transpose(int64_t rows, int64_t cols, const double* src, double* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACRO_ABI_H */
