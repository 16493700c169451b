// fields.cuh - registered vector-field device functors.
//
// A functor plays the role of the Python callable inside ODETerm(vector_field)
// (diffrax/_term.py:174-211): f(t, y, args) -> dy/dt.  `t` arrives in user time, i.e. already
// multiplied by `direction` the way WrapTerm.vf does (_term.py:738-740).  Parameters are the
// Python-float `args`; they are weakly typed in JAX, so they are rounded ONCE to the working
// dtype on the host (make<R>) and live in kernel-parameter constant memory.
//
// Interface (duck-typed, used by ensemble_kernel.cuh):
//   static constexpr int kId, kDim;  static constexpr bool kSde;
//   template <class R> struct P;                                   // POD parameters
//   template <class R> static P<R> make(const double *p, int n);    // host
//   template <class R> static __device__ void eval(const P<R>&, R t, const R (&y)[kDim], R (&f)[kDim]);
//   (kSde) template <class R> static __device__ R diffusion(const P<R>&, R t);  // additive scalar noise
#pragma once
#include "common.cuh"

namespace dfx {

// dy = -lambda * y   (test/test_integrate.py:60, test_saveat_solution.py:29-41)
template <int D>
struct DecayField {
  static constexpr int kId = DFX_FIELD_DECAY;
  static constexpr int kDim = D;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 1;
  template <class R> struct P { R lambda; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[D], R (&f)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) f[i] = -p.lambda * y[i];
  }
};

// benchmarks/lotka_volterra.py:13-20: [a*x + b*x*y, c*y + d*x*y]
struct LotkaVolterraField {
  static constexpr int kId = DFX_FIELD_LOTKA_VOLTERRA;
  static constexpr int kDim = 2;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 4;
  template <class R> struct P { R a, b, c, d; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0], (R)p[1], (R)p[2], (R)p[3]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[2], R (&f)[2]) {
    const R x = y[0], yy = y[1];
    f[0] = p.a * x + (p.b * x) * yy;
    f[1] = p.c * yy + (p.d * x) * yy;
  }
};

// Lorenz-63 (BASELINE config 2): [sigma (y - x), x (rho - z) - y, x y - beta z]
struct LorenzField {
  static constexpr int kId = DFX_FIELD_LORENZ;
  static constexpr int kDim = 3;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 3;
  template <class R> struct P { R sigma, rho, beta; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0], (R)p[1], (R)p[2]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[3], R (&f)[3]) {
    f[0] = p.sigma * (y[1] - y[0]);
    f[1] = y[0] * (p.rho - y[2]) - y[1];
    f[2] = y[0] * y[1] - p.beta * y[2];
  }
};

// Planar circular restricted three-body problem in the rotating frame (BASELINE config 3).
// State (x, y, vx, vy); primaries at (-mu, 0) and (1 - mu, 0).
struct Cr3bpField {
  static constexpr int kId = DFX_FIELD_CR3BP;
  static constexpr int kDim = 4;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 1;
  template <class R> struct P { R mu, mup; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0], (R)1 - (R)p[0]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[4], R (&f)[4]) {
    const R x = y[0], yy = y[1], vx = y[2], vy = y[3];
    const R dx1 = x + p.mu, dx2 = x - p.mup;
    const R r1s = dx1 * dx1 + yy * yy, r2s = dx2 * dx2 + yy * yy;
    const R w1 = p.mup / (r1s * r_sqrt(r1s)), w2 = p.mu / (r2s * r_sqrt(r2s));  // (1-mu)/r1^3, mu/r2^3
    f[0] = vx;
    f[1] = vy;
    f[2] = x + R(2) * vy - w1 * dx1 - w2 * dx2;
    f[3] = yy - R(2) * vx - (w1 + w2) * yy;
  }
};

// y0' = y1, y1' = -w0^2 y0 + A sin(w t): a time-dependent field that exercises the stage times
// t0 + c_i dt and the "exactly t1 when c_i == 1" rule (runge_kutta.py:1023).
struct ForcedOscField {
  static constexpr int kId = DFX_FIELD_FORCED_OSC;
  static constexpr int kDim = 2;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 3;
  template <class R> struct P { R w0sq, amp, w; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0], (R)p[1], (R)p[2]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R t, const R (&y)[2], R (&f)[2]) {
    f[0] = y[1];
    f[1] = -p.w0sq * y[0] + p.amp * r_sin(p.w * t);
  }
};

// van der Pol oscillator: y0' = y1, y1' = mu (1 - y0^2) y1 - y0
struct VdpField {
  static constexpr int kId = DFX_FIELD_VDP;
  static constexpr int kDim = 2;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 1;
  template <class R> struct P { R mu; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[2], R (&f)[2]) {
    f[0] = y[1];
    f[1] = p.mu * (R(1) - y[0] * y[0]) * y[1] - y[0];
  }
};

// Ornstein-Uhlenbeck (BASELINE config 5): dy = theta (mu - y) dt + sigma dW, additive scalar noise.
// As an ODE field (levy_area == none) only the drift is used.
struct OuField {
  static constexpr int kId = DFX_FIELD_OU;
  static constexpr int kDim = 1;
  static constexpr bool kSde = true;
  static constexpr int kNumParams = 3;
  template <class R> struct P { R theta, mu, sigma; };
  template <class R> static P<R> make(const double *p, int) { return P<R>{(R)p[0], (R)p[1], (R)p[2]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[1], R (&f)[1]) {
    f[0] = p.theta * (p.mu - y[0]);
  }
  template <class R> static __device__ __forceinline__ R diffusion(const P<R> &p, R) { return p.sigma; }
};

}  // namespace dfx
