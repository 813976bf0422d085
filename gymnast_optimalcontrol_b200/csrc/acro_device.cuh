// acro_device.cuh - register-resident device functions of the acrobot hot path (FP64).
//
// One thread owns one problem: the 4-vector state, the 2x4 lower block of the discrete
// linearisation, the symmetric 4x4 Riccati matrix P (10 registers) and the costate p live in
// registers; nothing here touches memory.  The structure of the model is used instead of
// dense 4x4 algebra:
//   A_d = I + dt*A_c has top rows [1 0 dt 0; 0 1 0 dt]   (trajectory_generation.py:161-164,
//   B_d = dt*B_c has zero top rows, and column 0 is zero   dynamics.py:153-164)
// unless the fully-actuated variant is selected (tau1 != 0).
// Reference lines are quoted per function (tg = trajectory_generation.py, tt =
// trajectory_tracking.py).
#pragma once
#include <cuda_runtime.h>

namespace acro {

// Lumped model constants, derived on the host from AcroParams (dynamics.py:64-90):
//   M11 = a1 + 2 h cos(th2), M12 = a3 + h cos(th2), M22 = a3,
//   G   = [g1 sin(th1) + g2 sin(th1+th2) ; g2 sin(th1+th2)]
struct Model {
  double a1, h, a3, g1, g2, f1, f2, dt, tau1;
  double h2;    // 2 h
  double det0;  // a1 a3 - a3^2 : det M = det0 - h^2 cos^2(th2)
  double hsq;   // h^2
  // sin/cos constants (see sincos2): they travel in the kernel-parameter constant bank so that they reach
  // the DFMAs through uniform registers (LDCU.128, two constants per instruction) instead of being rebuilt
  // from 32-bit immediates before every use.
  double tc[16];
};

// Per-problem physical parameters (dynamics.py:15-61: params_1/2/3 are three such sets): rows m1, m2, l1, lc1, l2,
// lc2, I1, I2, g, f1, f2 of an (11, B) array.  dt, the actuation flag and the trig constants stay those of the shared
// model.  Same expressions as make_model() on the host.
#define ACRO_N_PHYS 11
__device__ __forceinline__ Model model_per_problem(const Model& base, const double* __restrict__ pb, int64_t B, int64_t b) {
  const double m1 = pb[0 * B + b], m2 = pb[1 * B + b], l1 = pb[2 * B + b], lc1 = pb[3 * B + b], lc2 = pb[5 * B + b];
  const double I1 = pb[6 * B + b], I2 = pb[7 * B + b], g = pb[8 * B + b];
  Model m = base;
  m.a1 = I1 + I2 + lc1 * lc1 * m1 + m2 * (l1 * l1 + lc2 * lc2);
  m.h = m2 * l1 * lc2;
  m.a3 = I2 + lc2 * lc2 * m2;
  m.g1 = g * (lc1 * m1 + m2 * l1);
  m.g2 = g * m2 * lc2;
  m.f1 = pb[9 * B + b];
  m.f2 = pb[10 * B + b];
  m.h2 = 2.0 * m.h;
  m.det0 = m.a1 * m.a3 - m.a3 * m.a3;
  m.hsq = m.h * m.h;
  return m;
}

// 2/pi, -pi/2 split in two, 1.5*2^52, then the minimax coefficients of fdlibm's __kernel_sin / __kernel_cos
// on [-pi/4, pi/4]:  sin r = r + r^3 (S1 + z (S2 + ... z S6)),  cos r = 1 - z/2 + z^2 (C1 + z (C2 + ... z C6)),  z = r^2
#define ACRO_TRIG_CONSTANTS                                                                                  \
  {6.36619772367581382433e-01, -1.57079632679489655800e+00, -6.12323399573676603587e-17, 6755399441055744.0, \
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,                     \
   2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,                      \
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,                      \
   -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}

struct Trig {
  double s1, c1, s2, c2, s12, c12;
};

// Largest |angle| the polynomial path handles: up to here round(x 2/pi) fits the low word of the 1.5*2^52
// trick and the two-FMA Cody-Waite reduction (first FMA exact) leaves an error below 1e-24.
#define ACRO_TRIG_FAST_MAX 1.0e9

// x with its sign bit XORed with bit 31 of `bits`
__device__ __forceinline__ double flip_sign(double x, int bits) {
  return __hiloint2double(__double2hiint(x) ^ (bits & 0x80000000), __double2loint(x));
}

// sin and cos of two angles in lockstep (~1 ulp).  No branch: for |x| > ACRO_TRIG_FAST_MAX the result is
// meaningless but finite, and callers that may see such angles check and redo with the library routine;
// inf / nan propagate to nan like the library.
__device__ __forceinline__ void sincos2(const Model& m, double xa, double xb, double& sa, double& ca, double& sb,
                                        double& cb) {
  const double qma = fma(xa, m.tc[0], m.tc[3]), qmb = fma(xb, m.tc[0], m.tc[3]);
  const int ja = __double2loint(qma), jb = __double2loint(qmb);
  const double qa = qma - m.tc[3], qb = qmb - m.tc[3];
  double ra = fma(qa, m.tc[1], xa), rb = fma(qb, m.tc[1], xb);
  ra = fma(qa, m.tc[2], ra);
  rb = fma(qb, m.tc[2], rb);
  const double za = ra * ra, zb = rb * rb;
  double psa = fma(za, m.tc[9], m.tc[8]), psb = fma(zb, m.tc[9], m.tc[8]);
  double pca = fma(za, m.tc[15], m.tc[14]), pcb = fma(zb, m.tc[15], m.tc[14]);
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    psa = fma(za, psa, m.tc[4 + k]);
    psb = fma(zb, psb, m.tc[4 + k]);
    pca = fma(za, pca, m.tc[10 + k]);
    pcb = fma(zb, pcb, m.tc[10 + k]);
  }
  const double sra = fma(ra * za, psa, ra), srb = fma(rb * zb, psb, rb);
  const double cra = fma(za * za, pca, fma(za, -0.5, 1.0)), crb = fma(zb * zb, pcb, fma(zb, -0.5, 1.0));
  const double s0a = (ja & 1) ? cra : sra, c0a = (ja & 1) ? sra : cra;
  const double s0b = (jb & 1) ? crb : srb, c0b = (jb & 1) ? srb : crb;
  // quadrant signs: flip the sign bit with integer logic instead of spending FP64-pipe slots on negations
  sa = flip_sign(s0a, ja << 30);
  ca = flip_sign(c0a, (ja + 1) << 30);
  sb = flip_sign(s0b, jb << 30);
  cb = flip_sign(c0b, (jb + 1) << 30);
}

// FAST = true: polynomial path (caller guarantees or checks |angles| <= ACRO_TRIG_FAST_MAX);
// FAST = false: CUDA library sincos (Payne-Hanek reduction for huge arguments).
template <bool FAST>
__device__ __forceinline__ Trig trig_of(const Model& m, double th1, double th2) {
  Trig t;
  if (FAST) {
    sincos2(m, th1, th2, t.s1, t.c1, t.s2, t.c2);
  } else {
    sincos(th1, &t.s1, &t.c1);
    sincos(th2, &t.s2, &t.c2);
  }
  // angle-addition instead of a third sincos: 4 flops, error ~2 ulp
  t.s12 = fma(t.s1, t.c2, t.c1 * t.s2);
  t.c12 = fma(t.c1, t.c2, -(t.s1 * t.s2));
  return t;
}

// 1/d for a normal, finite d: hardware approximation (about 20 bits) + two Newton steps, no special-case
// branch.  Used for det M, which is bounded away from 0 and infinity by the physical parameters.
__device__ __forceinline__ double rcp_nr(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// Pieces of one evaluation of the equations of motion that the Jacobian reuses.
struct Eom {
  double M11, M12, M22, inv_det, dd1, dd2;
};

// qdd = M^-1 (tau - (C+F) qd - G), tau = [tau1*u0, u1]   (dynamics.py:197-213, 64-90)
// ACT = false: the caller knows that the first joint is not actuated (m.tau1 == 0; the Newton solver refuses the
// fully-actuated plant), so the torque term and its run-time select disappear.  x + 0.0 == x except for the sign of
// a zero, so the result is the same number.
template <bool ACT = true>
__device__ __forceinline__ Eom eom(const Model& m, const Trig& t, double w1, double w2, double u0,
                                   double u1) {
  Eom e;
  e.M11 = fma(m.h2, t.c2, m.a1);
  e.M12 = fma(m.h, t.c2, m.a3);
  e.M22 = m.a3;
  const double hs2 = m.h * t.s2;
  const double grav2 = m.g2 * t.s12;
  const double grav1 = fma(m.g1, t.s1, grav2);
  // r1 = tau1 + h s2 w2 w1 + h s2 (w1+w2) w2 - f1 w1 - G1 = tau1 + h s2 w2 (2 w1 + w2) - (f1 w1 + G1)
  // r2 = u1 - h s2 w1^2 - f2 w2 - G2                        = (u1 - h s2 w1 w1) - (f2 w2 + G2)
  double r1 = fma(hs2 * w2, fma(2.0, w1, w2), -fma(m.f1, w1, grav1));
  if (ACT) r1 += (m.tau1 != 0.0) ? m.tau1 * u0 : 0.0;
  const double r2 = fma(-(hs2 * w1), w1, u1) - fma(m.f2, w2, grav2);
  // det M = M11 M22 - M12^2 = (a1 a3 - a3^2) - h^2 cos^2(th2): two dependent operations after cos(th2)
  const double det = fma(-m.hsq, t.c2 * t.c2, m.det0);
  e.inv_det = rcp_nr(det);
  e.dd1 = (e.M22 * r1 - e.M12 * r2) * e.inv_det;
  e.dd2 = (e.M11 * r2 - e.M12 * r1) * e.inv_det;
  return e;
}

// continuous_dynamics(xx, uu)  dynamics.py:197-213
template <bool FAST>
__device__ __forceinline__ void f_eval_t(const Model& m, const double x[4], double u0, double u1, double f[4]) {
  const Trig t = trig_of<FAST>(m, x[0], x[1]);
  const Eom e = eom(m, t, x[2], x[3], u0, u1);
  f[0] = x[2];
  f[1] = x[3];
  f[2] = e.dd1;
  f[3] = e.dd2;
}

// Range tracking on the high words (integer pipe): hi(|x|) as an int orders finite doubles like their magnitude.
// "Too large" = finite and above ACRO_TRIG_FAST_MAX; inf / nan are left to propagate through the fast path.
#define ACRO_TRIG_FAST_MAX_HI 0x41cdcd65  /* high word of 1.0e9 */
__device__ __forceinline__ int abs_hi(double x) { return __double2hiint(x) & 0x7fffffff; }
__device__ __forceinline__ bool hi_too_large(int h) { return h > ACRO_TRIG_FAST_MAX_HI && h < 0x7ff00000; }
__device__ __forceinline__ bool angles_ok(double a, double b) { return !hi_too_large(max(abs_hi(a), abs_hi(b))); }

__device__ __forceinline__ void f_eval(const Model& m, const double x[4], double u0, double u1, double f[4]) {
  if (angles_ok(x[0], x[1]))
    f_eval_t<true>(m, x, u0, u1, f);
  else
    f_eval_t<false>(m, x, u0, u1, f);
}

// dynamics(xx, uu): classic RK4, zero-order hold on u   dynamics.py:177-195.
// Returns the largest high word of |angle| any of the four stages saw.
template <bool FAST>
__device__ __forceinline__ int rk4_step_t(const Model& m, const double x[4], double u0, double u1, double xn[4]) {
  const double h = m.dt, hh = 0.5 * m.dt, h6 = m.dt * (1.0 / 6.0);
  double k1[4], k2[4], k3[4], k4[4], y[4];
  int amax = max(abs_hi(x[0]), abs_hi(x[1]));
  f_eval_t<FAST>(m, x, u0, u1, k1);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(hh, k1[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k2);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(hh, k2[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k3);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(h, k3[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = fma(2.0, k3[i], fma(2.0, k2[i], k1[i])) + k4[i];
    xn[i] = fma(h6, s, x[i]);  // x + (dt/6) (k1 + 2 k2 + 2 k3 + k4)   dynamics.py:193
  }
  return amax;
}

// One branch per step instead of one per sin/cos: run the polynomial path, and only if some stage angle
// was beyond its range (a diverging rollout on its way to overflow) redo the step with the library routine.
// the library-sincos version is a real function call: it keeps its registers and code out of the hot kernels
// (arguments and results by value, so the fast path never has to keep its arrays addressable in local memory)
struct Vec4 {
  double v[4];
};
__device__ __noinline__ Vec4 rk4_step_slow(const Model& m, double x0, double x1, double x2, double x3, double u0,
                                           double u1) {
  const double x[4] = {x0, x1, x2, x3};
  Vec4 o;
  rk4_step_t<false>(m, x, u0, u1, o.v);
  return o;
}
__device__ __forceinline__ void rk4_step(const Model& m, const double x[4], double u0, double u1, double xn[4]) {
  const int amax = rk4_step_t<true>(m, x, u0, u1, xn);
  if (hi_too_large(amax)) {
    const Vec4 o = rk4_step_slow(m, x[0], x[1], x[2], x[3], u0, u1);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = o.v[i];
  }
}

// Lower two rows of the continuous Jacobians (dynamics.py:153-170, 217-226):
//   ac[r][j] = A_c[2+r][j],  bc0[r] = B_c[2+r][0] (zero unless fully actuated),
//   bc1[r] = B_c[2+r][1].   Rows 0-1 of A_c are [0 0 1 0; 0 0 0 1], of B_c zero.
struct LinC {
  double ac[2][4];
  double bc0[2];
  double bc1[2];
};

// Jacobian from the pieces of an evaluation of the equations of motion at the same (x, u)
__device__ __forceinline__ LinC jacobian_from(const Model& m, const Trig& t, const Eom& e, double w1, double w2) {
  const double hs2 = m.h * t.s2, hc2 = m.h * t.c2;
  const double gc12 = m.g2 * t.c12;
  // columns of d rhs / d x_j  minus (dM/dx_j) qdd   (only theta2 changes M)
  double d0[4], d1[4];
  d0[0] = -fma(m.g1, t.c1, gc12);
  d1[0] = -gc12;
  d0[1] = hc2 * (w1 * w2 + (w1 + w2) * w2) - gc12 + hs2 * (2.0 * e.dd1 + e.dd2);
  d1[1] = -hc2 * w1 * w1 - gc12 + hs2 * e.dd1;
  d0[2] = 2.0 * hs2 * w2 - m.f1;
  d1[2] = -2.0 * hs2 * w1;
  d0[3] = 2.0 * hs2 * (w1 + w2);
  d1[3] = -m.f2;
  LinC L;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    L.ac[0][j] = (e.M22 * d0[j] - e.M12 * d1[j]) * e.inv_det;
    L.ac[1][j] = (e.M11 * d1[j] - e.M12 * d0[j]) * e.inv_det;
  }
  L.bc1[0] = -e.M12 * e.inv_det;
  L.bc1[1] = e.M11 * e.inv_det;
  const bool act = (m.tau1 != 0.0);
  L.bc0[0] = act ? m.tau1 * e.M22 * e.inv_det : 0.0;
  L.bc0[1] = act ? -m.tau1 * e.M12 * e.inv_det : 0.0;
  return L;
}

template <bool FAST>
__device__ __forceinline__ LinC linearize_c_t(const Model& m, const double x[4], double u0, double u1) {
  const Trig t = trig_of<FAST>(m, x[0], x[1]);
  const Eom e = eom(m, t, x[2], x[3], u0, u1);
  return jacobian_from(m, t, e, x[2], x[3]);
}

__device__ __noinline__ LinC linearize_c_slow(const Model& m, double x0, double x1, double x2, double x3, double u0,
                                              double u1) {
  const double x[4] = {x0, x1, x2, x3};
  return linearize_c_t<false>(m, x, u0, u1);
}
__device__ __forceinline__ LinC linearize_c(const Model& m, const double x[4], double u0, double u1) {
  if (angles_ok(x[0], x[1])) return linearize_c_t<true>(m, x, u0, u1);
  return linearize_c_slow(m, x[0], x[1], x[2], x[3], u0, u1);
}

// Forward-Euler discretisation (tg:161-164) of the lower block:
//   a[r][j] = A_d[2+r][j] = delta + dt*ac[r][j],  b[r] = B_d[2+r][1] = dt*bc1[r].
struct LinD {
  double a[2][4];
  double b[2];
  double b0[2];  // column 0 of B_d (zero unless fully actuated)
};

__device__ __forceinline__ LinD discretize(const LinC& c, double dt) {
  LinD d;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int j = 0; j < 4; ++j) d.a[r][j] = (j == r + 2 ? 1.0 : 0.0) + dt * c.ac[r][j];
    d.b[r] = dt * c.bc1[r];
    d.b0[r] = dt * c.bc0[r];
  }
  return d;
}

__device__ __forceinline__ LinD linearize_d(const Model& m, const double x[4], double u0, double u1) {
  return discretize(linearize_c(m, x, u0, u1), m.dt);
}

// RK4 step that also returns the discrete linearisation about (x, u): the first stage already evaluates the
// sines, cosines and accelerations the Jacobian needs, so the backward pass of the next Newton iteration can
// read A_d, B_d instead of redoing two sincos and one evaluation of the equations of motion per step.
template <bool FAST>
__device__ __forceinline__ int rk4_step_lin_t(const Model& m, const double x[4], double u0, double u1,
                                                 double xn[4], LinD& L) {
  const double h = m.dt, hh = 0.5 * m.dt, h6 = m.dt * (1.0 / 6.0);
  double k1[4], k2[4], k3[4], k4[4], y[4];
  int amax = max(abs_hi(x[0]), abs_hi(x[1]));
  {
    const Trig t = trig_of<FAST>(m, x[0], x[1]);
    const Eom e = eom(m, t, x[2], x[3], u0, u1);
    k1[0] = x[2];
    k1[1] = x[3];
    k1[2] = e.dd1;
    k1[3] = e.dd2;
    L = discretize(jacobian_from(m, t, e, x[2], x[3]), m.dt);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(hh, k1[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k2);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(hh, k2[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k3);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma(h, k3[i], x[i]);
  amax = max(amax, max(abs_hi(y[0]), abs_hi(y[1])));
  f_eval_t<FAST>(m, y, u0, u1, k4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = fma(2.0, k3[i], fma(2.0, k2[i], k1[i])) + k4[i];
    xn[i] = fma(h6, s, x[i]);  // x + (dt/6) (k1 + 2 k2 + 2 k3 + k4)   dynamics.py:193
  }
  return amax;
}

struct StepLin {
  double xn[4];
  LinD L;
};
__device__ __noinline__ StepLin rk4_step_lin_slow(const Model& m, double x0, double x1, double x2, double x3, double u0,
                                                  double u1) {
  const double x[4] = {x0, x1, x2, x3};
  StepLin o;
  rk4_step_lin_t<false>(m, x, u0, u1, o.xn, o.L);
  return o;
}
__device__ __forceinline__ void rk4_step_lin(const Model& m, const double x[4], double u0, double u1, double xn[4],
                                             LinD& L) {
  const int amax = rk4_step_lin_t<true>(m, x, u0, u1, xn, L);
  if (hi_too_large(amax)) {
    const StepLin o = rk4_step_lin_slow(m, x[0], x[1], x[2], x[3], u0, u1);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = o.xn[i];
    L = o.L;
  }
}

// sincos2 for NA angles in lockstep: 2*NA independent polynomial chains cover the 8-cycle DFMA latency.
// `jz` is added to the quadrant index: callers pass 0, or an opaque zero that carries a scheduling dependency.
template <int NA>
__device__ __forceinline__ void sincos_n(const Model& m, const double x[NA], double s[NA], double c[NA], int jz = 0) {
  double qm[NA], q[NA], r[NA], z[NA], ps[NA], pc[NA];
  int j[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) qm[a] = fma(x[a], m.tc[0], m.tc[3]);
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    j[a] = __double2loint(qm[a]) + jz;
    q[a] = qm[a] - m.tc[3];
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) r[a] = fma(q[a], m.tc[1], x[a]);
#pragma unroll
  for (int a = 0; a < NA; ++a) r[a] = fma(q[a], m.tc[2], r[a]);
#pragma unroll
  for (int a = 0; a < NA; ++a) z[a] = r[a] * r[a];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    ps[a] = fma(z[a], m.tc[9], m.tc[8]);
    pc[a] = fma(z[a], m.tc[15], m.tc[14]);
  }
#pragma unroll
  for (int k = 3; k >= 0; --k) {
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      ps[a] = fma(z[a], ps[a], m.tc[4 + k]);
      pc[a] = fma(z[a], pc[a], m.tc[10 + k]);
    }
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const double sr = fma(r[a] * z[a], ps[a], r[a]);
    const double cr = fma(z[a] * z[a], pc[a], fma(z[a], -0.5, 1.0));
    const double s0 = (j[a] & 1) ? cr : sr, c0 = (j[a] & 1) ? sr : cr;
    s[a] = flip_sign(s0, j[a] << 30);
    c[a] = flip_sign(c0, (j[a] + 1) << 30);
  }
}

__device__ __forceinline__ Trig trig_from(double s1, double c1, double s2, double c2) {
  Trig t;
  t.s1 = s1;
  t.c1 = c1;
  t.s2 = s2;
  t.c2 = c2;
  t.s12 = fma(t.s1, t.c2, t.c1 * t.s2);
  t.c12 = fma(t.c1, t.c2, -(t.s1 * t.s2));
  return t;
}

// The RK4 step of rk4_step_t<true>, same expressions, written in the order of its data dependencies instead of
// stage by stage: the angles of a stage depend on the velocities of the previous stage, i.e. on the accelerations
// of the stage before that, so the sines and cosines of stage s+1 do not wait for the equations of motion of
// stage s.  Stages 1 and 2 (whose angles need nothing but x) share one lockstep sincos of four angles; the
// sincos of stages 3 and 4 stand before the equations of motion of stages 2 and 3, which fill their latency.
// `mid` is a hook for the caller's loads of the next step: issued in the middle of the step they hide behind FP64
// instructions.  It returns an (opaque) zero derived from what it loaded, which is added to the quadrant index of
// the stage-4 sines: the instruction scheduler, which would otherwise push loads nobody needs before the next
// step to the very end of this one (where nothing is left to overlap them with), has to complete them by then.
template <class Mid>
__device__ __forceinline__ int rk4_step_overlap(const Model& m, const double x[4], double u0, double u1,
                                                double xn[4], Mid mid) {
  const double h = m.dt, hh = 0.5 * m.dt, h6 = m.dt * (1.0 / 6.0);
  const double th[4] = {x[0], x[1], fma(hh, x[2], x[0]), fma(hh, x[3], x[1])};
  int amax = max(max(abs_hi(th[0]), abs_hi(th[1])), max(abs_hi(th[2]), abs_hi(th[3])));
  double sn[4], cs[4];
  sincos_n<4>(m, th, sn, cs);
  const Trig t1 = trig_from(sn[0], cs[0], sn[1], cs[1]);
  const Trig t2 = trig_from(sn[2], cs[2], sn[3], cs[3]);
  const Eom e1 = eom<false>(m, t1, x[2], x[3], u0, u1);
  const double w21 = fma(hh, e1.dd1, x[2]), w22 = fma(hh, e1.dd2, x[3]);
  const double th3a = fma(hh, w21, x[0]), th3b = fma(hh, w22, x[1]);
  amax = max(amax, max(abs_hi(th3a), abs_hi(th3b)));
  const Trig t3 = trig_of<true>(m, th3a, th3b);
  const int jz = mid();
  const Eom e2 = eom<false>(m, t2, w21, w22, u0, u1);
  const double w31 = fma(hh, e2.dd1, x[2]), w32 = fma(hh, e2.dd2, x[3]);
  const double th4a = fma(h, w31, x[0]), th4b = fma(h, w32, x[1]);
  amax = max(amax, max(abs_hi(th4a), abs_hi(th4b)));
  const double th4[2] = {th4a, th4b};
  double s4[2], c4[2];
  sincos_n<2>(m, th4, s4, c4, jz);
  const Trig t4 = trig_from(s4[0], c4[0], s4[1], c4[1]);
  const Eom e3 = eom<false>(m, t3, w31, w32, u0, u1);
  const double w41 = fma(h, e3.dd1, x[2]), w42 = fma(h, e3.dd2, x[3]);
  const Eom e4 = eom<false>(m, t4, w41, w42, u0, u1);
  const double k1[4] = {x[2], x[3], e1.dd1, e1.dd2}, k2[4] = {w21, w22, e2.dd1, e2.dd2};
  const double k3[4] = {w31, w32, e3.dd1, e3.dd2}, k4[4] = {w41, w42, e4.dd1, e4.dd2};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = fma(2.0, k3[i], fma(2.0, k2[i], k1[i])) + k4[i];
    xn[i] = fma(h6, s, x[i]);  // x + (dt/6) (k1 + 2 k2 + 2 k3 + k4)   dynamics.py:193
  }
  return amax;
}

// ---------------------------------------------------------------------------------------
// RK4 step with incremental sines and cosines.
//
// With dt = 0.02 the four stage angles of a step and the first angle of the next step lie within a fraction of a
// radian of each other, so only ONE stage per step needs a full (range-reduced, degree-13/14) sincos: stage 2,
// whose angle x + dt/2 * omega is known when the step starts and whose sines are not needed before the second
// evaluation of the equations of motion.  The others follow by rotating a known (sin, cos) pair through the small
// angle difference d, with sin d and cos d - 1 from short Taylor polynomials (no range reduction, no quadrant
// logic, half the dependency chain):
//     stage 1 of a step  <- stage 4 of the previous step   |d| = O(dt^2 * second difference of the accelerations)
//     stage 3            <- stage 2                        |d| = dt^2/4 * |acceleration|
//     stage 4            <- stage 2                        |d| ~ dt/2 * |omega|
// Truncation errors are below 1e-19 (small: |d| <= 2^-6, terms to d^7 / d^6) and 3e-17 (medium: |d| <= 2^-3,
// terms to d^9 / d^10); a rotated pair is never rotated more than twice before the next full sincos, so nothing
// accumulates.  A step whose differences exceed the bounds (|omega| > 12 rad/s, |acceleration| > 150 rad/s^2:
// a diverging candidate) reports `bad` and the caller redoes it with rk4_step_redo.
// ---------------------------------------------------------------------------------------
struct TrigCarry {
  double th[2], s[2], c[2];  // stage-4 angles of the previous step, their sines and cosines
};

// Taylor coefficients of sin d and cos d - 1.  As literals the compiler rebuilds them from 32-bit immediates in
// every step of a loop (forty MOV / LDC instructions per time step, 9 % of the chain warp's step); RotC::held()
// launders them through an empty asm so that they stay in registers across the loop.  The rollout kernels, which
// are compiled for three blocks per SM, keep the literals (RotC::literals()).
struct RotC {
  double s3, s5, s7, s9, c2, c4, c6, c8, c10;
  __device__ __forceinline__ static RotC literals() {
    return RotC{-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0,
                -1.0 / 3628800.0};
  }
  __device__ __forceinline__ static RotC held() {
    RotC r = literals();
    asm volatile("" : "+d"(r.s3), "+d"(r.s5), "+d"(r.s7), "+d"(r.s9));
    asm volatile("" : "+d"(r.c2), "+d"(r.c4), "+d"(r.c6), "+d"(r.c8), "+d"(r.c10));
    return r;
  }
};

// (s, c) = (sin, cos)(theta)  ->  (sin, cos)(theta + d)
__device__ __forceinline__ void rot_small(const RotC& k, double d, double s, double c, double& so, double& co) {
  const double d2 = d * d;
  double ps = fma(d2, k.s7, k.s5);
  double pc = fma(d2, k.c6, k.c4);
  ps = fma(d2, ps, k.s3);
  pc = fma(d2, pc, k.c2);
  const double sd = fma(d * d2, ps, d);  // sin d
  const double cm = d2 * pc;             // cos d - 1
  so = s + fma(s, cm, c * sd);
  co = c + fma(c, cm, -(s * sd));
}
__device__ __forceinline__ void rot_medium(const RotC& k, double d, double s, double c, double& so, double& co) {
  const double d2 = d * d;
  double ps = fma(d2, k.s9, k.s7);
  double pc = fma(d2, k.c10, k.c8);
  ps = fma(d2, ps, k.s5);
  pc = fma(d2, pc, k.c6);
  ps = fma(d2, ps, k.s3);
  pc = fma(d2, pc, k.c4);
  pc = fma(d2, pc, k.c2);
  const double sd = fma(d * d2, ps, d);
  const double cm = d2 * pc;
  so = s + fma(s, cm, c * sd);
  co = c + fma(c, cm, -(s * sd));
}
#define ACRO_ROT_SMALL_HI 0x3f900000   /* high word of 2^-6 */
#define ACRO_ROT_MEDIUM_HI 0x3fc00000  /* high word of 2^-3 */

// x + 0 in a way that carries a scheduling dependency on `jz` (an opaque zero, see rk4_step_overlap)
__device__ __forceinline__ double tie(double x, int jz) {
  return __hiloint2double(__double2hiint(x) + jz, __double2loint(x));
}

// Returns true when the step left the validity range of the incremental formulas (then xn and tc are garbage).
template <bool ACT, class Mid>
__device__ __forceinline__ bool rk4_step_rot(const Model& m, const RotC& rc, const double x[4], double u0, double u1,
                                             double xn[4], TrigCarry& tc, Mid mid) {
  const double h = m.dt, hh = 0.5 * m.dt, h6 = m.dt * (1.0 / 6.0);
  // stage 2: the one full sincos
  const double th2[2] = {fma(hh, x[2], x[0]), fma(hh, x[3], x[1])};
  double s2[2], c2[2];
  sincos_n<2>(m, th2, s2, c2);
  // stage 1 from the previous step's stage 4
  const double d1[2] = {x[0] - tc.th[0], x[1] - tc.th[1]};
  double s1[2], c1[2];
  rot_small(rc, d1[0], tc.s[0], tc.c[0], s1[0], c1[0]);
  rot_small(rc, d1[1], tc.s[1], tc.c[1], s1[1], c1[1]);
  const Trig t1 = trig_from(s1[0], c1[0], s1[1], c1[1]);
  const Eom e1 = eom<ACT>(m, t1, x[2], x[3], u0, u1);
  const double w21 = fma(hh, e1.dd1, x[2]), w22 = fma(hh, e1.dd2, x[3]);
  // stage 3 from stage 2
  const double d3[2] = {fma(hh, w21, x[0]) - th2[0], fma(hh, w22, x[1]) - th2[1]};
  double s3[2], c3[2];
  rot_small(rc, d3[0], s2[0], c2[0], s3[0], c3[0]);
  rot_small(rc, d3[1], s2[1], c2[1], s3[1], c3[1]);
  const Trig t3 = trig_from(s3[0], c3[0], s3[1], c3[1]);
  const int jz = mid();
  const Trig t2 = trig_from(s2[0], c2[0], s2[1], c2[1]);
  const Eom e2 = eom<ACT>(m, t2, w21, w22, u0, u1);
  const double w31 = fma(hh, e2.dd1, x[2]), w32 = fma(hh, e2.dd2, x[3]);
  // stage 4 from stage 2
  const double th4[2] = {fma(h, w31, x[0]), fma(h, w32, x[1])};
  const double d4[2] = {tie(th4[0] - th2[0], jz), th4[1] - th2[1]};
  double s4[2], c4[2];
  rot_medium(rc, d4[0], s2[0], c2[0], s4[0], c4[0]);
  rot_medium(rc, d4[1], s2[1], c2[1], s4[1], c4[1]);
  const Trig t4 = trig_from(s4[0], c4[0], s4[1], c4[1]);
  const Eom e3 = eom<ACT>(m, t3, w31, w32, u0, u1);
  const double w41 = fma(h, e3.dd1, x[2]), w42 = fma(h, e3.dd2, x[3]);
  const Eom e4 = eom<ACT>(m, t4, w41, w42, u0, u1);
  const double k1[4] = {x[2], x[3], e1.dd1, e1.dd2}, k2[4] = {w21, w22, e2.dd1, e2.dd2};
  const double k3[4] = {w31, w32, e3.dd1, e3.dd2}, k4[4] = {w41, w42, e4.dd1, e4.dd2};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = fma(2.0, k3[i], fma(2.0, k2[i], k1[i])) + k4[i];
    xn[i] = fma(h6, s, x[i]);  // x + (dt/6) (k1 + 2 k2 + 2 k3 + k4)   dynamics.py:193
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    tc.th[a] = th4[a];
    tc.s[a] = s4[a];
    tc.c[a] = c4[a];
  }
  const int small = max(max(abs_hi(d1[0]), abs_hi(d1[1])), max(abs_hi(d3[0]), abs_hi(d3[1])));
  const int medium = max(abs_hi(d4[0]), abs_hi(d4[1]));
  const int range = max(abs_hi(th2[0]), abs_hi(th2[1]));
  // (non-finite differences have high words >= 0x7ff00000 and land here too)
  return small >= ACRO_ROT_SMALL_HI || medium >= ACRO_ROT_MEDIUM_HI || range > ACRO_TRIG_FAST_MAX_HI;
}

// start of a pass, or after a redone step: carry = (angles of x, their sines and cosines), so that the next
// step's stage-1 rotation is through d = 0 exactly
__device__ __noinline__ TrigCarry trig_carry_at(const Model& m, double th1, double th2) {
  TrigCarry tc;
  tc.th[0] = th1;
  tc.th[1] = th2;
  if (angles_ok(th1, th2)) {
    sincos2(m, th1, th2, tc.s[0], tc.c[0], tc.s[1], tc.c[1]);
  } else {
    sincos(th1, &tc.s[0], &tc.c[0]);
    sincos(th2, &tc.s[1], &tc.c[1]);
  }
  return tc;
}
// the step again, with a full sincos per stage (and the library routine beyond its range)
__device__ __noinline__ Vec4 rk4_step_redo(const Model& m, double x0, double x1, double x2, double x3, double u0,
                                           double u1) {
  const double x[4] = {x0, x1, x2, x3};
  Vec4 o;
  rk4_step(m, x, u0, u1, o.v);
  return o;
}

// One step of a rollout with the incremental sincos and its fallback; `tc` is carried from step to step
// (initialise with trig_carry_at(m, x[0], x[1])).  One thread per problem: the redo is an ordinary divergent branch.
__device__ __forceinline__ void rk4_step_inc(const Model& m, const double x[4], double u0, double u1, double xn[4],
                                             TrigCarry& tc) {
  const bool bad = rk4_step_rot<true>(m, RotC::literals(), x, u0, u1, xn, tc, []() { return 0; });
  if (bad) {
    const Vec4 o = rk4_step_redo(m, x[0], x[1], x[2], x[3], u0, u1);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = o.v[i];
    tc = trig_carry_at(m, xn[0], xn[1]);
  }
}

// ---------------------------------------------------------------------------------------
// Symmetric 4x4 in 10 registers: index of (i,j), i<=j
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ constexpr int sym(int i, int j) {
  return (i <= j) ? (i * 4 - (i * (i - 1)) / 2 + (j - i)) : (j * 4 - (j * (j - 1)) / 2 + (i - j));
}

// 2x2 solve with partial pivoting (what LAPACK dgesv does for np.linalg.solve / inv at
// tg:203-204, tt:157,200).  Factor once, apply to several right-hand sides.
struct Lu2 {
  double inv_p, l, u01, inv_u11;
  bool swap;
};

__device__ __forceinline__ Lu2 lu2(double g00, double g01, double g10, double g11) {
  Lu2 f;
  f.swap = fabs(g10) > fabs(g00);
  const double p = f.swap ? g10 : g00, q = f.swap ? g11 : g01;
  const double r = f.swap ? g00 : g10, s = f.swap ? g01 : g11;
  f.inv_p = 1.0 / p;
  f.l = r * f.inv_p;
  f.u01 = q;
  f.inv_u11 = 1.0 / fma(-f.l, q, s);
  return f;
}

__device__ __forceinline__ void lu2_solve(const Lu2& f, double b0, double b1, double& x0, double& x1) {
  const double y0 = f.swap ? b1 : b0, y1 = f.swap ? b0 : b1;
  x1 = fma(-f.l, y0, y1) * f.inv_u11;
  x0 = fma(-f.u01, x1, y0) * f.inv_p;
}

// ---------------------------------------------------------------------------------------
// One step of the backward Riccati recursion on the structured linearisation.
//
// AFFINE = true : calculate_K_and_sigma (tg:183-216)
//     G = Rh + B'PB, F = B'PA, g = r + B'p, K = -G^-1 F, sigma = -G^-1 g, dJ += g'sigma,
//     P <- Qh + A'PA - K'GK,  p <- q + A'p - K'G sigma
// AFFINE = false: solve_LQR_tracking / the MPC sweep (tt:191-201, tt:80-117)
//     K = -(Rh + B'PB)^-1 B'PA,  P <- Qh + A'PA + (A'PB) K
// Both P updates are the same expression: G K = -F gives  -K'GK = K'F = (A'PB) K, and
// -K'G sigma = K'g.  Qh/Rh are the blocks as used by the caller (2Q, 2R for Newton; Q, R
// for tracking).  P is symmetric (10 values); K is 2x4 row-major.
// ---------------------------------------------------------------------------------------
// First column of G = Rh + B'PB when the first actuator does not exist: (Rh00, Rh01), constant over the sweep.
struct Lu2Col {
  double inv_p, l;
  bool swap;
};
__device__ __forceinline__ Lu2Col lu2_col(double g00, double g10) {
  Lu2Col c;
  c.swap = fabs(g10) > fabs(g00);
  c.inv_p = 1.0 / (c.swap ? g10 : g00);
  c.l = (c.swap ? g00 : g10) * c.inv_p;
  return c;
}

template <bool AFFINE, bool ACT0, class QH>
__device__ __forceinline__ void riccati_step(double P[10], double p[4], const LinD& L, double dt,
                                             const QH& Qh, const Lu2Col& col, double Rh00, double Rh01, double Rh11,
                                             const double qv[4], const double r[2], double K[8],
                                             double sig[2], double& dJ) {
  // M = P * A_d    (A_d rows 0-1 are [1 0 dt 0; 0 1 0 dt])
  double M[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double pi0 = P[sym(i, 0)], pi1 = P[sym(i, 1)], pi2 = P[sym(i, 2)], pi3 = P[sym(i, 3)];
    M[i][0] = fma(pi3, L.a[1][0], fma(pi2, L.a[0][0], pi0));
    M[i][1] = fma(pi3, L.a[1][1], fma(pi2, L.a[0][1], pi1));
    M[i][2] = fma(pi3, L.a[1][2], fma(pi2, L.a[0][2], dt * pi0));
    M[i][3] = fma(pi3, L.a[1][3], fma(pi2, L.a[0][3], dt * pi1));
  }
  // S = A_d' * M (upper triangle)
  double S[10];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double top = (i < 2) ? M[i][j] : dt * M[i - 2][j];
      S[sym(i, j)] = fma(L.a[1][i], M[3][j], fma(L.a[0][i], M[2][j], top));
    }
  }
  // F = B_d' M (rows of F: input 0 and input 1), B'PB
  double F1[4], F0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    F1[j] = fma(L.b[1], M[3][j], L.b[0] * M[2][j]);
    F0[j] = ACT0 ? fma(L.b0[1], M[3][j], L.b0[0] * M[2][j]) : 0.0;
  }
  // P b for the two input columns (only rows 2,3 of P b are needed for B'PB)
  const double Pb1_2 = fma(P[sym(2, 3)], L.b[1], P[sym(2, 2)] * L.b[0]);
  const double Pb1_3 = fma(P[sym(3, 3)], L.b[1], P[sym(2, 3)] * L.b[0]);
  double G00 = Rh00, G01 = Rh01, G11 = Rh11 + fma(L.b[1], Pb1_3, L.b[0] * Pb1_2);
  if (ACT0) {
    const double Pb0_2 = fma(P[sym(2, 3)], L.b0[1], P[sym(2, 2)] * L.b0[0]);
    const double Pb0_3 = fma(P[sym(3, 3)], L.b0[1], P[sym(2, 3)] * L.b0[0]);
    G00 += fma(L.b0[1], Pb0_3, L.b0[0] * Pb0_2);
    G01 += fma(L.b[1], Pb0_3, L.b[0] * Pb0_2);
  }
  // 2x2 solve G [K | sigma] = -[F | g] by LU with partial pivoting (what LAPACK does).  Without the first
  // actuator the first column of G is the constant (Rh00, Rh01): pivot choice, 1/pivot and the multiplier come
  // precomputed in `col`, F has a zero first row, and only ONE reciprocal (of u11) sits on the P -> P chain.
  double g0 = 0.0, g1 = 0.0;
  if (ACT0) {
    const Lu2 lu = lu2(G00, G01, G01, G11);
#pragma unroll
    for (int j = 0; j < 4; ++j) lu2_solve(lu, -F0[j], -F1[j], K[j], K[4 + j]);
    if (AFFINE) {
      g0 = r[0] + fma(L.b0[1], p[3], L.b0[0] * p[2]);
      g1 = r[1] + fma(L.b[1], p[3], L.b[0] * p[2]);
      lu2_solve(lu, -g0, -g1, sig[0], sig[1]);
    }
  } else {
    const double q = col.swap ? G11 : G01, s = col.swap ? G01 : G11;
    const double inv_u11 = rcp_nr(fma(-col.l, q, s));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // right-hand side (0, -F1[j])
      const double x1 = (col.swap ? col.l * F1[j] : -F1[j]) * inv_u11;
      K[4 + j] = x1;
      K[j] = (col.swap ? fma(-q, x1, -F1[j]) : -(q * x1)) * col.inv_p;
    }
    if (AFFINE) {
      g0 = r[0];
      g1 = r[1] + fma(L.b[1], p[3], L.b[0] * p[2]);
      const double y0 = col.swap ? -g1 : -g0, y1 = col.swap ? -g0 : -g1;
      sig[1] = fma(-col.l, y0, y1) * inv_u11;
      sig[0] = fma(-q, sig[1], y0) * col.inv_p;
    }
  }
  if (AFFINE) {
    dJ += fma(g1, sig[1], g0 * sig[0]);
    // p <- q + A_d' p + K' g
    const double p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
    const double ap[4] = {fma(L.a[1][0], p3, fma(L.a[0][0], p2, p0)), fma(L.a[1][1], p3, fma(L.a[0][1], p2, p1)),
                          fma(L.a[1][2], p3, fma(L.a[0][2], p2, dt * p0)),
                          fma(L.a[1][3], p3, fma(L.a[0][3], p2, dt * p1))};
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = qv[i] + ap[i] + fma(K[4 + i], g1, K[i] * g0);
  }
  // P <- Qh + S + K'F
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = i; j < 4; ++j) {
      double kf = K[4 + i] * F1[j];
      if (ACT0) kf = fma(K[i], F0[j], kf);
      P[sym(i, j)] = Qh(i, j) + S[sym(i, j)] + kf;
    }
  }
}

// The P half of riccati_step<*, false> alone (same expressions in the same order): what the recursion itself needs.
// Used by the chain warp of k_newton_duo and by the MPC sweeps.
// SW: which row of the constant first column of G is the pivot (lu2_col): 0 = the first, 1 = the second (rows
// swapped), 2 = decided per lane at run time (per-problem weights).  With shared weights the choice is the same
// for every problem and every step, so the caller branches once per pass and the selects (two dozen per step)
// disappear from the step.
// ROW0 = false skips the first row of K (the gain of the input that does not act on the plant, needed by the callers
// that output K but not by the P recursion itself: the inner steps of an MPC sweep).
template <int SW, bool ROW0 = true, class QH>
__device__ __forceinline__ void riccati_chain_step(double P[10], const LinD& L, double dt, const QH& Qh,
                                                   const Lu2Col& col, double Rh01, double Rh11, double K[8],
                                                   double& inv_u11, double& qsel) {
  const bool swap = (SW == 2) ? col.swap : (SW == 1);
  double M[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double pi0 = P[sym(i, 0)], pi1 = P[sym(i, 1)], pi2 = P[sym(i, 2)], pi3 = P[sym(i, 3)];
    M[i][0] = fma(pi3, L.a[1][0], fma(pi2, L.a[0][0], pi0));
    M[i][1] = fma(pi3, L.a[1][1], fma(pi2, L.a[0][1], pi1));
    M[i][2] = fma(pi3, L.a[1][2], fma(pi2, L.a[0][2], dt * pi0));
    M[i][3] = fma(pi3, L.a[1][3], fma(pi2, L.a[0][3], dt * pi1));
  }
  double S[10];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double top = (i < 2) ? M[i][j] : dt * M[i - 2][j];
      S[sym(i, j)] = fma(L.a[1][i], M[3][j], fma(L.a[0][i], M[2][j], top));
    }
  }
  double F1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) F1[j] = fma(L.b[1], M[3][j], L.b[0] * M[2][j]);
  const double Pb1_2 = fma(P[sym(2, 3)], L.b[1], P[sym(2, 2)] * L.b[0]);
  const double Pb1_3 = fma(P[sym(3, 3)], L.b[1], P[sym(2, 3)] * L.b[0]);
  const double G01 = Rh01, G11 = Rh11 + fma(L.b[1], Pb1_3, L.b[0] * Pb1_2);
  const double q = swap ? G11 : G01, s = swap ? G01 : G11;
  inv_u11 = rcp_nr(fma(-col.l, q, s));
  qsel = q;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double x1 = (swap ? col.l * F1[j] : -F1[j]) * inv_u11;
    K[4 + j] = x1;
    if (ROW0) K[j] = (swap ? fma(-q, x1, -F1[j]) : -(q * x1)) * col.inv_p;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double kf = K[4 + i] * F1[j];
      P[sym(i, j)] = Qh(i, j) + S[sym(i, j)] + kf;
    }
  }
}

// Dense variant for caller-supplied A (4x4) and B (4x2): compute_P_inf (tt:144-165) and
// solver_mpc (tt:73-140) take arbitrary matrices.  P <- Q + A'PA + (A'PB)K, K = -(R+B'PB)^-1 B'PA.
template <class QH>
__device__ __forceinline__ void riccati_step_dense(double P[10], const double A[16], const double Bm[8],
                                                   const QH& Qh, double R00, double R01, double R11,
                                                   double K[8]) {
  double M[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = P[sym(i, 0)] * A[j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(P[sym(i, k)], A[k * 4 + j], s);
      M[i][j] = s;
    }
  double S[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      double s = A[i] * M[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(A[k * 4 + i], M[k][j], s);
      S[sym(i, j)] = s;
    }
  double F[2][4], PB[4][2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = Bm[c] * M[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(Bm[k * 2 + c], M[k][j], s);
      F[c][j] = s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = P[sym(i, 0)] * Bm[c];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(P[sym(i, k)], Bm[k * 2 + c], s);
      PB[i][c] = s;
    }
  }
  double G[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = a; c < 2; ++c) {
      double s = Bm[a] * PB[0][c];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(Bm[k * 2 + a], PB[k][c], s);
      G[a][c] = s;
    }
  const Lu2 lu = lu2(R00 + G[0][0], R01 + G[0][1], R01 + G[0][1], R11 + G[1][1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) lu2_solve(lu, -F[0][j], -F[1][j], K[j], K[4 + j]);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j)
      P[sym(i, j)] = Qh(i, j) + S[sym(i, j)] + fma(K[4 + i], F[1][j], K[i] * F[0][j]);
}

// General dense affine step for caller-supplied lists: calculate_K_and_sigma (tg:183-216) with
// arbitrary A (4x4), B (4x2), Q_t (4x4 sym), R_t (2x2 sym), S_t (2x4), q_t, r_t:
//   G = R + B'PB, F = S + B'PA, g = r + B'p, K = -G^-1 F, sigma = -G^-1 g, dJ += g'sigma,
//   P <- Q + A'PA - K'GK (= Q + A'PA + K'F),  p <- q + A'p - K'G sigma (= q + A'p + K'g)
__device__ __forceinline__ void riccati_step_lists(double P[10], double p[4], const double A[16],
                                                   const double Bm[8], const double Q[16], const double R[4],
                                                   const double Sx[8], const double q[4], const double r[2],
                                                   double K[8], double sig[2], double& dJ) {
  double M[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = P[sym(i, 0)] * A[j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(P[sym(i, k)], A[k * 4 + j], s);
      M[i][j] = s;
    }
  double APA[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      double s = A[i] * M[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(A[k * 4 + i], M[k][j], s);
      APA[sym(i, j)] = s;
    }
  double F[2][4], PB[4][2], g[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = Sx[c * 4 + j];
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fma(Bm[k * 2 + c], M[k][j], s);
      F[c][j] = s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = P[sym(i, 0)] * Bm[c];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(P[sym(i, k)], Bm[k * 2 + c], s);
      PB[i][c] = s;
    }
    double s = r[c];
#pragma unroll
    for (int k = 0; k < 4; ++k) s = fma(Bm[k * 2 + c], p[k], s);
    g[c] = s;
  }
  double G[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = a; c < 2; ++c) {
      double s = R[a * 2 + c];
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fma(Bm[k * 2 + a], PB[k][c], s);
      G[a][c] = s;
    }
  const Lu2 lu = lu2(G[0][0], G[0][1], G[0][1], G[1][1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) lu2_solve(lu, -F[0][j], -F[1][j], K[j], K[4 + j]);
  lu2_solve(lu, -g[0], -g[1], sig[0], sig[1]);
  dJ += fma(g[1], sig[1], g[0] * sig[0]);
  double pn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = q[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) s = fma(A[k * 4 + i], p[k], s);
    pn[i] = s + fma(K[4 + i], g[1], K[i] * g[0]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = pn[i];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j)
      P[sym(i, j)] = Q[i * 4 + j] + APA[sym(i, j)] + fma(K[4 + i], F[1][j], K[i] * F[0][j]);
}

}  // namespace acro
