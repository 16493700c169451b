// interp.cuh - the solvers' local dense interpolants.
//
// Replaces: LocalLinearInterpolation (_local_interpolation.py:29-47),
// ThirdOrderHermitePolynomialInterpolation.from_k (50-92), FourthOrderPolynomialInterpolation
// (95-139) with Dopri5's c_mid (dopri5.py:36-47), _Tsit5Interpolation.evaluate (tsit5.py:116-156)
// and _Dopri8Interpolation.evaluate (dopri8.py:294-303).  Expression order follows the
// reference so that theta == 0 reproduces y0 bit-exactly (test_global_interpolation.py:346).
#pragma once
#include "common.cuh"
#include "tableaux.cuh"

namespace dfx {

enum { kInterpLinear = 0, kInterpHermite = 1, kInterpDopri5 = 2, kInterpTsit5 = 3, kInterpDopri8 = 4 };

template <class R> struct TabConst;
template <> struct TabConst<double> {
  static __device__ __forceinline__ double cmid(int i) { return kDopri5Cmid_f64[i]; }
  static __device__ __forceinline__ double d8(int i, int m) { return kDopri8Eval_f64[i * 6 + m]; }
};
template <> struct TabConst<float> {
  static __device__ __forceinline__ float cmid(int i) { return kDopri5Cmid_f32[i]; }
  static __device__ __forceinline__ float d8(int i, int m) { return kDopri8Eval_f32[i * 6 + m]; }
};

// Forward-mode dual number: the interpolants below are written once, generically in the type of theta; instantiated with
// Dual<R> they return the derivative the reference obtains from jax.jvp of `evaluate` (AbstractPath.derivative, _path.py).
template <class R> struct Dual { R v, d; };
template <class R> __device__ __forceinline__ Dual<R> operator+(Dual<R> a, Dual<R> b) { return {a.v + b.v, a.d + b.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator+(Dual<R> a, R b) { return {a.v + b, a.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator+(R a, Dual<R> b) { return {a + b.v, b.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator-(Dual<R> a, Dual<R> b) { return {a.v - b.v, a.d - b.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator-(Dual<R> a, R b) { return {a.v - b, a.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator*(Dual<R> a, Dual<R> b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <class R> __device__ __forceinline__ Dual<R> operator*(Dual<R> a, R b) { return {a.v * b, a.d * b}; }
template <class R> __device__ __forceinline__ Dual<R> operator*(R a, Dual<R> b) { return {a * b.v, a * b.d}; }
template <class R> __device__ __forceinline__ R zero_like(R) { return R(0); }
template <class R> __device__ __forceinline__ Dual<R> zero_like(Dual<R>) { return {R(0), R(0)}; }
template <class R> __device__ __forceinline__ Dual<R> &operator+=(Dual<R> &a, Dual<R> b) { a = a + b; return a; }

// The interpolant of kind KIND at theta = th (T = R: value; T = Dual<R>: value and d/dtheta).  k is [S][D] (ignored for linear).
// K: anything spelling an element k[j][c] (a plain R[S][D] array, or the shared-memory view of ensemble_kernel.cuh)
template <int KIND, class R, int S, int D, class T, class K>
__device__ __forceinline__ void interp_core(const T th, const R (&y0)[D], const R (&y1)[D], const K &k, T (&out)[D]) {
  if constexpr (KIND == kInterpLinear) {
#pragma unroll
    for (int c = 0; c < D; ++c) out[c] = y0[c] + th * (y1[c] - y0[c]);
  } else if constexpr (KIND == kInterpHermite) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const R k0 = k[0][c], k1 = k[S - 1][c];
      const R a = k0 + k1 + R(2) * y0[c] - R(2) * y1[c];
      const R b = R(-2) * k0 - k1 - R(3) * y0[c] + R(3) * y1[c];
      T p = R(0) * th + a;   // jnp.polyval: Horner from zero
      p = p * th + b;
      p = p * th + k0;
      p = p * th + y0[c];
      out[c] = p;
    }
  } else if constexpr (KIND == kInterpDopri5) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      R acc = R(0);
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j != 1) acc += TabConst<R>::cmid(j) * k[j][c];   // c_mid[1] == 0
      const R ymid = y0[c] + acc;
      const R f0 = k[0][c], f1 = k[S - 1][c];
      const R a = R(2) * (f1 - f0) - R(8) * (y1[c] + y0[c]) + R(16) * ymid;
      const R b = R(5) * f0 - R(3) * f1 + R(18) * y0[c] + R(14) * y1[c] - R(32) * ymid;
      const R cc = f1 - R(4) * f0 - R(11) * y0[c] - R(5) * y1[c] + R(16) * ymid;
      T p = R(0) * th + a;
      p = p * th + b;
      p = p * th + cc;
      p = p * th + f0;
      p = p * th + y0[c];
      out[c] = p;
    }
  } else if constexpr (KIND == kInterpTsit5) {
    const T x = th, x2 = x * x;
    T b[7];
    b[0] = R(-1.0530884977290216) * x * (x - R(1.3299890189751412)) * (x2 - R(1.4364028541716351) * x + R(0.7139816917074209));
    b[1] = R(0.1017) * x2 * (x2 - R(2.1966568338249754) * x + R(1.2949852507374631));
    b[2] = R(2.490627285651252793) * x2 * (x2 - R(2.38535645472061657) * x + R(1.57803468208092486));
    b[3] = R(-16.54810288924490272) * (x - R(1.21712927295533244)) * (x - R(0.61620406037800089)) * x2;
    b[4] = R(47.37952196281928122) * (x - R(1.203071208372362603)) * (x - R(0.658047292653547382)) * x2;
    b[5] = R(-34.87065786149660974) * (x - R(1.2)) * (x - R(0.666666666666666667)) * x2;
    b[6] = R(2.5) * (x - R(1)) * (x - R(0.6)) * x2;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T acc = zero_like(th);
#pragma unroll
      for (int j = 0; j < 7; ++j) acc += b[j] * k[j][c];
      out[c] = y0[c] + acc;
    }
  } else {  // kInterpDopri8
    T w[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) {
      if (dopri8_eval_row_nonzero(j)) {
        T p = R(0) * th + TabConst<R>::d8(j, 0);
#pragma unroll
        for (int m = 1; m < 6; ++m) p = p * th + TabConst<R>::d8(j, m);
        w[j] = p * th;
      } else {
        w[j] = zero_like(th);
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T acc = zero_like(th);
#pragma unroll
      for (int j = 0; j < 14; ++j)
        if (dopri8_eval_row_nonzero(j)) acc += w[j] * k[j][c];
      out[c] = y0[c] + acc;
    }
  }
}

// Evaluate the interpolant of kind KIND on [t0, t1] at time t.
template <int KIND, class R, int S, int D, class K>
__device__ __forceinline__ void interp_eval(R t0, R t1, const R (&y0)[D], const R (&y1)[D], const K &k,
                                            R t, R (&out)[D]) {
  interp_core<KIND, R, S, D, R>(linear_rescale(t0, t, t1), y0, y1, k, out);
}

// d/dt of the interpolant at time t: the tangent of `evaluate` w.r.t. t (AbstractPath.derivative via jax.jvp), including
// linear_rescale's own tangent where(t0 == t1, 0, 1 / (t1 - t0)) (_misc.py:71-82).
template <int KIND, class R, int S, int D, class K>
__device__ __forceinline__ void interp_deriv(R t0, R t1, const R (&y0)[D], const R (&y1)[D], const K &k,
                                             R t, R (&out)[D]) {
  const Dual<R> th{linear_rescale(t0, t, t1), (t0 == t1) ? R(0) : R(1) / (t1 - t0)};
  Dual<R> o[D];
  interp_core<KIND, R, S, D, Dual<R>>(th, y0, y1, k, o);
#pragma unroll
  for (int c = 0; c < D; ++c) out[c] = o[c].d;
}

}  // namespace dfx
