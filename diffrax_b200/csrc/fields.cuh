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
//   template <class R> static P<R> make(const double *p, int n, const void *weights);  // host; weights = device pointer
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0]}; }
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0], (R)p[1], (R)p[2], (R)p[3]}; }
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0], (R)p[1], (R)p[2]}; }
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0], (R)1 - (R)p[0]}; }
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0], (R)p[1], (R)p[2]}; }
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
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0]}; }
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
  // additive noise sigma + sigma_t * t: the optional 4th parameter gives time-dependent diffusion (the
  // getting-started.md:63-84 example dy = -y dt + t/10 dw; exercises ShARK's g(t1) - g(t0) term, srk.py:612-618)
  template <class R> struct P { R theta, mu, sigma, sigma_t; };
  template <class R> static P<R> make(const double *p, int n, const void *) { return P<R>{(R)p[0], (R)p[1], (R)p[2], n >= 4 ? (R)p[3] : R(0)}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[1], R (&f)[1]) {
    f[0] = p.theta * (p.mu - y[0]);
  }
  template <class R> static __device__ __forceinline__ R diffusion(const P<R> &p, R t) { return p.sigma + p.sigma_t * t; }
};

// D independent OU components, each driven by its own Brownian motion: VirtualBrownianTree(shape=(D,)) with a diagonal
// diffusion (vf_prod = g (.) dW).  Same field id as OuField; the launcher key is (field, dim).
template <int D>
struct OuDiagField {
  static constexpr int kId = DFX_FIELD_OU;
  static constexpr int kDim = D;
  static constexpr int kNoise = D;
  static constexpr bool kSde = true;
  static constexpr int kNumParams = 3;
  template <class R> using P = OuField::P<R>;
  template <class R> static P<R> make(const double *p, int n, const void *w) { return OuField::make<R>(p, n, w); }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[D], R (&f)[D]) {
#pragma unroll
    for (int c = 0; c < D; ++c) f[c] = p.theta * (p.mu - y[c]);
  }
  template <class R> static __device__ __forceinline__ R diffusion(const P<R> &p, R t) { return p.sigma + p.sigma_t * t; }
};

// D Ornstein-Uhlenbeck components driven through a constant D x M diffusion MATRIX by VirtualBrownianTree(shape=(M,)):
// ControlTerm(lambda t, y, args: G, bm) with G of shape (D, M), whose prod is tensordot(G, dW) (_term.py:267-268, 417-427).
// Field id DFX_FIELD_OU_MATRIX + M; params [theta, mu, G row-major (D*M)].
template <int D, int M>
struct OuMatrixField {
  static constexpr int kId = DFX_FIELD_OU_MATRIX + M;
  static constexpr int kDim = D;
  static constexpr int kNoise = M;
  static constexpr bool kMatrixNoise = true;
  static constexpr bool kSde = true;
  static constexpr int kNumParams = 2 + D * M;
  template <class R> struct P { R theta, mu; R g[D][M]; };
  template <class R> static P<R> make(const double *p, int, const void *) {
    P<R> o;
    o.theta = (R)p[0]; o.mu = (R)p[1];
    for (int c = 0; c < D; ++c)
      for (int j = 0; j < M; ++j) o.g[c][j] = (R)p[2 + c * M + j];
    return o;
  }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[D], R (&f)[D]) {
#pragma unroll
    for (int c = 0; c < D; ++c) f[c] = p.theta * (p.mu - y[c]);
  }
  // row c of G . w  (tensordot over the Brownian axis, ascending j)
  template <class R> static __device__ __forceinline__ R noise_prod(const P<R> &p, R, const R (&w)[M], int c) {
    R acc = R(0);
#pragma unroll
    for (int j = 0; j < M; ++j) acc += p.g[c][j] * w[j];
    return acc;
  }
};

// Geometric Brownian motion, dy = mu y dt + sigma y dW: STATE-DEPENDENT (multiplicative) diagonal diffusion,
// ControlTerm(lambda t, y, args: sigma * y, bm) with a scalar Brownian motion driving every component (tensordot with a 0-d
// control).  Heun converges to the Stratonovich solution (heun.py:24-33); ShARK / SRK need additive noise and are refused.
template <int D>
struct GbmField {
  static constexpr int kId = DFX_FIELD_GBM;
  static constexpr int kDim = D;
  static constexpr bool kStateNoise = true;
  static constexpr bool kSde = true;
  static constexpr int kNumParams = 2;
  template <class R> struct P { R mu, sigma; };
  template <class R> static P<R> make(const double *p, int, const void *) { return P<R>{(R)p[0], (R)p[1]}; }
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &p, R, const R (&y)[D], R (&f)[D]) {
#pragma unroll
    for (int c = 0; c < D; ++c) f[c] = p.mu * y[c];
  }
  template <class R> static __device__ __forceinline__ R noise_prod(const P<R> &p, R, const R (&y)[D], const R (&w)[1], int c) {
    return (p.sigma * y[c]) * w[0];
  }
};

// Neural-ODE vector field (BASELINE config 4): eqx.nn.MLP(d -> W -> W -> d) with softplus hidden activations and
// a tanh output (docs/examples/neural_ode.ipynb cell 5; benchmarks/small_neural_ode.py:25-28), evaluated per thread
// on the FP32 CUDA cores.  This is the exact-fp32 reference implementation of the field inside the generic ensemble
// kernel; the tensor-core (tcgen05, 3xTF32) kernel lives in mlp_kernel.cuh.
// Weights (device memory, fp32/fp64 as R, eqx Linear layout (out, in) row-major):
//   W1[W][D], b1[W], W2[W][W], b2[W], W3[D][W], b3[D]
// Every lane reads the same weight at the same time, so the loads are L1 broadcasts (one transaction per warp).
template <int D, int W>
struct MlpField {
  static constexpr int kId = DFX_FIELD_MLP;
  static constexpr int kDim = D;
  static constexpr int kWidth = W;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = 2;  // [width, depth] for checking
  static constexpr long long kNumWeights = (long long)W * D + W + (long long)W * W + W + (long long)D * W + D;
  template <class R> struct P { const R *w; };
  template <class R> static P<R> make(const double *, int, const void *weights) { return P<R>{(const R *)weights}; }

  template <class R> static __device__ __forceinline__ R softplus(R x) {
    // jax.nn.softplus = logaddexp(x, 0) = max(x, 0) + log1p(exp(-|x|))
    return r_max(x, R(0)) + r_log1p(r_exp(-r_abs(x)));
  }

  template <class R>
  static __device__ __noinline__ void eval(const P<R> &p, R, const R (&y)[D], R (&f)[D]) {
    const R *W1 = p.w, *b1 = W1 + W * D, *W2 = b1 + W, *b2 = W2 + W * W, *W3 = b2 + W, *b3 = W3 + D * W;
    R h1[W];
#pragma unroll
    for (int o = 0; o < W; ++o) {
      R acc = R(0);
#pragma unroll
      for (int k = 0; k < D; ++k) acc += __ldg(W1 + o * D + k) * y[k];
      h1[o] = softplus(acc + __ldg(b1 + o));
    }
    R out[D];
#pragma unroll
    for (int c = 0; c < D; ++c) out[c] = R(0);
#pragma unroll 1
    for (int o = 0; o < W; ++o) {
      const R *row = W2 + o * W;
      R acc = R(0);
#pragma unroll
      for (int k = 0; k < W; ++k) acc += __ldg(row + k) * h1[k];
      const R h2 = softplus(acc + __ldg(b2 + o));
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] += __ldg(W3 + c * W + o) * h2;
    }
#pragma unroll
    for (int c = 0; c < D; ++c) f[c] = r_tanh(out[c] + __ldg(b3 + c));
  }
};

}  // namespace dfx
