"""Handles for the registered device functors (csrc/fields.cuh).

A handle stands where the reference takes a Python callable ``vector_field(t, y, args)``
(diffrax/_term.py:174-211): ``ODETerm(fields.Lorenz(10., 28., 8/3))``.  The numeric arguments
are what ``args`` / closed-over Python floats are in the reference; they are rounded once to
the working dtype on the host (JAX weak typing).  SDE functors expose ``.drift`` and
``.diffusion`` for ``MultiTerm(ODETerm(f.drift), ControlTerm(f.diffusion, bm))``.
"""
from __future__ import annotations

from . import _lib


class FieldPart:
    def __init__(self, field, part):
        self.field, self.part = field, part


class Field:
    name = ""
    dim = 0          # 0 == any (templated on y0's last dimension)
    is_sde = False

    @property
    def field_id(self):
        return _lib.FIELD_IDS[self.name]

    @property
    def drift(self):
        return FieldPart(self, "drift")

    @property
    def diffusion(self):
        if not self.is_sde:
            raise AttributeError(f"{type(self).__name__} has no diffusion")
        return FieldPart(self, "diffusion")

    def params(self):
        return []

    def weights(self, xp, dtype):
        return None

    def cuda_twin(self, dim):
        """This functor written as a `CudaField` (the same statements as csrc/fields.cuh, so the same bits), or None.  It is what
        `diffeqsolve(..., args=[N, n_params])` runs on a built-in functor: the per-trajectory-parameter kernel variant exists for
        generated functors (their parameters are a plain array `p[]`)."""
        return None

    def cpp_functor(self, dim, bm_dim):
        """The csrc/fields.cuh functor type for a state of dimension `dim` driven by a Brownian motion of shape `(bm_dim,)`
        (0: shape () or none); None when there is none.  Used to instantiate (field, solver, dtype) combinations that are not
        among the prebuilt kernels on first use - `ensure_builtin_kernel`."""
        return None


class LinearDecay(Field):
    """dy/dt = -lam * y  (test/test_integrate.py:60, test_saveat_solution.py:29-41); kernels prebuilt for d in {1,2,3},
    other dimensions (up to 8) are instantiated on first use."""
    name = "decay"

    def __init__(self, lam=1.0):
        self.lam = float(lam)

    def params(self):
        return [self.lam]

    def cuda_twin(self, dim):
        return CudaField(dim, " ".join(f"f[{i}] = -p[0] * y[{i}];" for i in range(dim)), params=self.params())

    def cpp_functor(self, dim, bm_dim):
        return f"::dfx::DecayField<{dim}>"


class LotkaVolterra(Field):
    """benchmarks/lotka_volterra.py:13-20: [a x + b x y, c y + d x y]."""
    name, dim = "lotka_volterra", 2

    def __init__(self, a=1.5, b=-1.0, c=-3.0, d=1.0):
        self.p = [float(a), float(b), float(c), float(d)]

    def params(self):
        return self.p

    def cuda_twin(self, dim):
        return CudaField(2, "const R x = y[0], yy = y[1]; f[0] = p[0] * x + (p[1] * x) * yy; f[1] = p[2] * yy + (p[3] * x) * yy;",
                         params=self.params())

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::LotkaVolterraField"


class Lorenz(Field):
    """Lorenz-63: [sigma (y-x), x (rho - z) - y, x y - beta z]."""
    name, dim = "lorenz", 3

    def __init__(self, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
        self.p = [float(sigma), float(rho), float(beta)]

    def params(self):
        return self.p

    def cuda_twin(self, dim):
        return CudaField(3, "f[0] = p[0] * (y[1] - y[0]); f[1] = y[0] * (p[1] - y[2]) - y[1]; f[2] = y[0] * y[1] - p[2] * y[2];",
                         params=self.params(), min_blocks_per_sm=6)

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::LorenzField"


class CR3BP(Field):
    """Planar circular restricted three-body problem, rotating frame, state (x, y, vx, vy)."""
    name, dim = "cr3bp", 4

    def __init__(self, mu=0.012277471):
        self.mu = float(mu)

    def params(self):
        return [self.mu]

    def cuda_twin(self, dim):
        return CudaField(4, "const R mu = p[0], mup = (R)1 - p[0]; const R x = y[0], yy = y[1], vx = y[2], vy = y[3]; "
                            "const R dx1 = x + mu, dx2 = x - mup; const R r1s = dx1 * dx1 + yy * yy, r2s = dx2 * dx2 + yy * yy; "
                            "const R w1 = mup / (r1s * ::dfx::r_sqrt(r1s)), w2 = mu / (r2s * ::dfx::r_sqrt(r2s)); "
                            "f[0] = vx; f[1] = vy; f[2] = x + R(2) * vy - w1 * dx1 - w2 * dx2; f[3] = yy - R(2) * vx - (w1 + w2) * yy;",
                         params=self.params())

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::Cr3bpField"


class ForcedOscillator(Field):
    """y0' = y1, y1' = -w0^2 y0 + A sin(w t)."""
    name, dim = "forced_osc", 2

    def __init__(self, w0sq=1.0, amp=1.0, w=2.0):
        self.p = [float(w0sq), float(amp), float(w)]

    def params(self):
        return self.p

    def cuda_twin(self, dim):
        return CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0] + p[1] * ::dfx::r_sin(p[2] * t);", params=self.params())

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::ForcedOscField"


class VanDerPol(Field):
    name, dim = "vdp", 2

    def __init__(self, mu=1.0):
        self.mu = float(mu)

    def params(self):
        return [self.mu]

    def cuda_twin(self, dim):
        return CudaField(2, "f[0] = y[1]; f[1] = p[0] * (R(1) - y[0] * y[0]) * y[1] - y[0];", params=self.params())

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::VdpField"


class OrnsteinUhlenbeck(Field):
    """dy = theta (mu - y) dt + (sigma + sigma_t t) dW (additive noise, optionally growing linearly in time).  With a state of
    dimension m > 1 the components are independent and driven by VirtualBrownianTree(shape=(m,)) (diagonal diffusion)."""
    name, dim, is_sde = "ou", 0, True

    def __init__(self, theta=1.0, mu=0.0, sigma=0.5, sigma_t=0.0):
        self.p = [float(theta), float(mu), float(sigma)] + ([float(sigma_t)] if sigma_t else [])

    def params(self):
        return self.p

    def cuda_twin(self, dim):
        p4 = (self.p + [0.0])[:4]      # [theta, mu, sigma, sigma_t]
        return CudaField(dim, " ".join(f"f[{i}] = p[0] * (p[1] - y[{i}]);" for i in range(dim)), params=p4, diffusion="p[2] + p[3] * t")

    def cpp_functor(self, dim, bm_dim):
        return "::dfx::OuField" if dim == 1 else f"::dfx::OuDiagField<{dim}>"


class GeometricBrownianMotion(Field):
    """dy = mu y dt + sigma y dW: state-dependent (multiplicative) diagonal diffusion with a scalar Brownian motion -
    ``ControlTerm(lambda t, y, args: sigma * y, VirtualBrownianTree(..., shape=(), ...))``.  Euler / Heun (Heun converges to the
    Stratonovich solution); kernels for state dimension 1 and 2."""
    name, dim, is_sde = "gbm", 0, True

    def __init__(self, mu=0.1, sigma=0.2):
        self.p = [float(mu), float(sigma)]

    def params(self):
        return self.p

    def cpp_functor(self, dim, bm_dim):
        return f"::dfx::GbmField<{dim}>"


class OrnsteinUhlenbeckMatrix(Field):
    """dy = theta (mu - y) dt + G dW with a constant ``[d, m]`` diffusion MATRIX ``G`` and an m-dimensional Brownian motion:
    ``ControlTerm(lambda t, y, args: G, VirtualBrownianTree(..., shape=(m,), ...))``, whose product is
    ``tensordot(G, dW)`` (_term.py:267-268, 417-427).  Kernels: (d, m) in {(2, 2), (3, 2), (2, 3)}."""
    name, is_sde = "ou_matrix", True

    def __init__(self, theta, mu, G):
        import numpy as np
        self.G = np.asarray(G, np.float64)
        if self.G.ndim != 2:
            raise ValueError("G must be a [d, m] matrix")
        self.dim, self.m = int(self.G.shape[0]), int(self.G.shape[1])
        self.p = [float(theta), float(mu)] + [float(v) for v in self.G.ravel()]

    @property
    def field_id(self):
        return _lib.FIELD_OU_MATRIX + self.m

    def params(self):
        return self.p

    def cpp_functor(self, dim, bm_dim):
        return f"::dfx::OuMatrixField<{self.dim}, {self.m}>"


class MLP(Field):
    """Neural-ODE vector field: ``eqx.nn.MLP(in=d, out=d, width, depth=2, activation=softplus, final_activation=tanh)``
    (docs/examples/neural_ode.ipynb cell 5).  ``layers`` is ``[(W1, b1), (W2, b2), (W3, b3)]`` with eqx ``Linear``
    weight shapes ``(out, in)``.  d=4, width=128, fp32 runs on the tcgen05 tensor-core kernel; other sizes (d <= 8, width <= 256) and
    fp64 run the per-thread functor, instantiated on first use."""
    name, dim = "mlp", 4

    def __init__(self, layers):
        import numpy as np
        self.layers = [(np.asarray(W), np.asarray(b)) for W, b in layers]
        if len(self.layers) != 3:
            raise ValueError("MLP functor supports depth=2 (three Linear layers)")
        self.width = self.layers[0][0].shape[0]
        self.dim = self.layers[0][0].shape[1]
        self._cache = {}

    @staticmethod
    def init(key_seed: int, d: int = 4, width: int = 128, dtype="float32"):
        """eqx.nn.Linear initialisation: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases (NumPy RNG)."""
        import numpy as np
        rng = np.random.default_rng(key_seed)
        layers = []
        for fan_in, fan_out in ((d, width), (width, width), (width, d)):
            lim = 1.0 / np.sqrt(fan_in)
            layers.append((rng.uniform(-lim, lim, (fan_out, fan_in)).astype(dtype), rng.uniform(-lim, lim, fan_out).astype(dtype)))
        return MLP(layers)

    def params(self):
        return [float(self.width), 2.0]

    @property
    def field_id(self):
        # the prebuilt kernels (tensor-core path included) are for d = 4, width 128; any other size is the generic per-thread
        # functor MlpField<d, width> under an id of its own, instantiated on first use (ensure_builtin_kernel)
        if (self.dim, self.width) == (4, 128):
            return _lib.FIELD_IDS["mlp"]
        return _lib.FIELD_USER + (1 << 28) + self.dim * 4096 + self.width

    def cpp_functor(self, dim, bm_dim):
        if (self.dim, self.width) == (4, 128):
            return "::dfx::MlpField<4, 128>"
        if not (1 <= self.dim <= 8 and 1 <= self.width <= 256):
            return None
        return (f"UserMlp; struct UserMlp : ::dfx::MlpField<{self.dim}, {self.width}> "
                f"{{ static constexpr int kId = {self.field_id}; }}")

    def flat(self, dtype):
        import numpy as np
        return np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in self.layers]).astype(dtype)

    def oracle_params(self):
        """[width, depth, W1, b1, W2, b2, W3, b3] as doubles (the oracle's field_params layout)."""
        import numpy as np
        return np.concatenate([[float(self.width), 2.0], self.flat(np.float64)])

    def weights(self, xp, dtype):
        import numpy as np
        key = (type(xp).__name__, str(getattr(xp, "device", "")), str(dtype))
        if key not in self._cache:
            np_dt = np.float32 if "32" in str(dtype) else np.float64
            self._cache[key] = xp.asarray(self.flat(np_dt), dtype)
        return self._cache[key]


# --------------------------------------------------------------------------------------
# User-supplied vector fields, compiled on first use
# --------------------------------------------------------------------------------------
_SOLVER_CPP = {0: "::dfx::Tsit5", 1: "::dfx::Dopri5", 2: "::dfx::Dopri8", 3: "::dfx::Heun", 4: "::dfx::Bosh3",
               5: "::dfx::Midpoint", 6: "::dfx::Ralston", 7: "::dfx::EulerSolver", 8: "::dfx::SharkSolver"}
_HALF = 0x100

class UserEvent:
    """Condition function number `index` of a `CudaField(events=[...])`, for `Event(cond_fn=...)` (_event.py:13-118): a
    real-valued expression in `t`, `y[i]`, `p[i]`; the solve stops on the step where it changes sign."""
    kind = _lib.EVENT_USER

    def __init__(self, field, index):
        self.field, self.index = field, int(index)

    def params(self, d, ctrl):
        return [float(self.index)]


_USER_TU = r"""// generated by diffrax_b200.fields.CudaField - do not edit
@DEFINES@
#include "launch.cuh"
namespace {
@PREAMBLE@
struct UserField {
  static constexpr int kId = @ID@;
  static constexpr int kDim = @DIM@;
  static constexpr bool kSde = @SDE@;
  static constexpr int kNumParams = @NP@;
@ARGS_TRAIT@
@NOISE_TRAITS@
  template <class R> struct P { R p[@NP1@]; };
  template <class R> static P<R> make(const double *q, int n, const void *) {
    P<R> o;
    for (int i = 0; i < @NP1@; ++i) o.p[i] = (i < n && i < @NP@) ? (R)q[i] : R(0);
    return o;
  }
  // vector_field(t, y, args) of ODETerm (_term.py:174-211): y[kDim] -> f[kDim]; p[] are the bound `args`
  template <class R>
  static __device__ __forceinline__ void eval(const P<R> &P_, R t, const R (&y)[@DIM@], R (&f)[@DIM@]) {
    [[maybe_unused]] const R *p = P_.p;
    (void)t;
@DRIFT@
  }
@NOISE_FNS@
};
[[maybe_unused]] constexpr bool kUseSde = UserField::kSde;
DFX_REGISTER(@REAL@, UserField, @SOLVER@, @LEVY@)
}  // namespace
"""


class CudaField(Field):
    """A vector field written by the user as CUDA C++ statements - what a Python ``vector_field(t, y, args)`` is to the
    reference's ``ODETerm`` (_term.py:174-211).  The solver kernels are templates over the field functor, so the statements
    are compiled into a kernel of their own (nvcc, sm_100a) the first time a (solver, dtype) combination is asked for
    (~10 s; cached under ``diffrax_b200/lib/user/`` by content hash) and registered through ``dfx_register_launcher``.

    ``drift``: statements that assign ``f[0..dim)`` from ``t``, ``y[0..dim)`` and the parameters ``p[0..n_params)``; ``R`` is the
    working type (``double`` / ``float``).  Example - the damped pendulum::

        pend = CudaField(2, "f[0] = y[1]; f[1] = -p[0] * sin(y[0]) - p[1] * y[1];", params=[9.81, 0.1])
        sol = diffeqsolve(ODETerm(pend), Dopri5(), 0.0, 10.0, None, y0, stepsize_controller=PIDController(1e-8, 1e-8))

    SDEs (``MultiTerm(ODETerm(f.drift), ControlTerm(f.diffusion, VirtualBrownianTree(...)))``), either
    * ``diffusion="<expr>"``: additive noise ``g(t)`` (an expression in ``t`` and ``p``) times a Brownian motion of shape ``()``
      (dim 1) or ``(dim,)`` (diagonal) - Euler / Heun / ShARK and ``HalfSolver`` of the latter two; or
    * ``noise="<statements>"``, ``noise_dim=m``: the general ``ControlTerm.prod`` (_term.py:417-427): assign ``gx[0..dim)``, the
      product ``g(t, y) . x`` with the Brownian increment ``x[0..m)`` (``m = 1`` for shape ``()``) - Euler / Heun (Stratonovich).

    ``wide=True``: a WIDE state (``dim`` up to 1024), one trajectory per warp (csrc/wide_kernel.cuh): the state is spread over
    the 32 lanes, and ``drift`` is written per component - statements that assign ``fi``, component ``i`` of f, from ``i``, the
    whole state ``y[0..D)``, ``p`` and ``t``; e.g. Lorenz-96: ``"fi = (y[(i + 1) % D] - y[(i + D - 2) % D]) * y[(i + D - 1) % D] - y[i] + p[0];"``.
    Parameter sweeps: ``diffeqsolve(..., args=A)`` with ``A`` of shape ``[N, len(params)]`` gives trajectory ``i`` the parameters
    ``p[] = A[i]`` (what a vmapped ``args`` is to the reference, _integrate.py:896); ``params`` then only fixes their number.
    ``events=["<expr>", ...]``: real-valued condition functions in ``t``, ``y``, ``p`` for ``Event(field.event(i), ...)`` - the
    reference's arbitrary ``cond_fn(t, y, args)`` (_event.py:13-118); with a root finder the crossing is located on the step's
    interpolant.  ``preamble``: device helper functions / constants placed before the functor.  ``min_blocks_per_sm``: occupancy target handed
    to ptxas (``__launch_bounds__``) instead of the register-budget heuristic of csrc/ensemble_kernel.cuh - e.g. 6 for a
    3-dimensional fp64 field with a 7-stage solver (what the built-in Lorenz/Dopri5 kernel uses)."""
    is_user = True

    def __init__(self, dim, drift, *, params=(), diffusion=None, noise=None, noise_dim=None, preamble="", name=None,
                 min_blocks_per_sm=None, events=(), defines=None, wide=False):
        import hashlib
        self.dim = int(dim)
        self.wide = bool(wide)
        if self.wide:
            if not 1 <= self.dim <= 1024:
                raise ValueError("CudaField(wide=True): 1 <= dim <= 1024")
            if diffusion is not None or noise is not None or events:
                raise ValueError("CudaField(wide=True) is an ODE functor: no diffusion / noise / events")
        elif not 1 <= self.dim <= 8:
            raise ValueError("CudaField: 1 <= dim <= 8 with one trajectory per thread; pass wide=True (one trajectory per warp, "
                             "dim <= 1024) and write the drift per component: `fi = ...` from `i`, `y[...]`, `p[...]`")
        if diffusion is not None and noise is not None:
            raise ValueError("CudaField: give `diffusion` (additive g(t)) or `noise` (general g(t, y) . x), not both")
        self.p = [float(v) for v in params]
        self.drift_src, self.diffusion_src, self.noise_src, self.preamble = str(drift), diffusion, noise, str(preamble)
        self.is_sde = diffusion is not None or noise is not None
        self.noise_dim = 1 if not self.is_sde else (int(noise_dim) if noise_dim else (self.dim if diffusion is not None else 1))
        if diffusion is not None and self.noise_dim not in (1, self.dim):
            raise ValueError("CudaField: additive `diffusion` is scalar (dim 1) or diagonal (noise_dim == dim)")
        if self.is_sde and not 1 <= self.noise_dim <= 8:
            raise ValueError("CudaField: 1 <= noise_dim <= 8")
        self.min_blocks = None if min_blocks_per_sm is None else int(min_blocks_per_sm)
        if self.min_blocks is not None and not 1 <= self.min_blocks <= 16:
            raise ValueError("CudaField: 1 <= min_blocks_per_sm <= 16")
        # `defines`: preprocessor switches of csrc/ensemble_kernel.cuh for THIS functor's kernels, e.g. the reference's plain
        # operation order {"DFX_OPT_CHAIN_Y0": 0, "DFX_OPT_LAST_STAGE_F": 0, "DFX_OPT_FAST_PID": 0, "DFX_OPT_ABSMAX_FP64": 0}
        self.defines = dict(defines or {})
        self.event_srcs = [str(e) for e in events]
        if len(self.event_srcs) > _lib.MAX_EVENTS:
            raise ValueError(f"CudaField: at most {_lib.MAX_EVENTS} condition functions")
        # everything that shapes the generated source identifies the functor (parameter VALUES do not: they are run-time data)
        key = "\0".join(self.event_srcs + ["wide" if self.wide else "thread", repr(sorted(self.defines.items())), str(self.dim),
                                           self.drift_src, str(diffusion), str(noise), str(self.noise_dim), self.preamble,
                                           str(len(self.p)), str(self.min_blocks)])
        self._hash = hashlib.sha256(key.encode()).hexdigest()[:16]
        self._id = _lib.FIELD_USER + int(self._hash[:7], 16)
        self.name = name or f"user_{self._hash}"
        self._ready = set()

    @property
    def field_id(self):
        return self._id

    def params(self):
        return self.p

    def event(self, index=0):
        """The `index`-th expression of `events=[...]` as a condition function: `Event(field.event(0), Newton(1e-10, 1e-10))`."""
        if not 0 <= index < len(self.event_srcs):
            raise IndexError(f"this CudaField defines {len(self.event_srcs)} condition function(s)")
        return UserEvent(self, index)

    @property
    def field_id_args(self):
        """The id of the kernel variant that reads per-trajectory parameters (`diffeqsolve(..., args=[N, n_params])`)."""
        return self._id + (1 << 29)

    def source(self, solver_id, dtype_id, levy, per_traj=False):
        inner = solver_id & ~_HALF
        solver = _SOLVER_CPP[inner]
        if solver_id & _HALF:
            solver = f"::dfx::HalfOf<{solver}>"
        np_ = len(self.p)
        if self.wide:
            rep = {"@DEFINES@": "".join(f"#define {k} {v}\n" for k, v in sorted(self.defines.items())), "@PREAMBLE@": self.preamble,
                   "@ARGS_TRAIT@": "  static constexpr bool kPerTrajArgs = true;" if per_traj else "",
                   "@ID@": str(self.field_id_args if per_traj else self._id), "@DIM@": str(self.dim), "@NP@": str(np_), "@NP1@": str(max(np_, 1)), "@DRIFT@": self.drift_src,
                   "@REAL@": "double" if dtype_id == _lib.F64 else "float", "@SOLVER@": solver}
            src = _WIDE_TU
            for k, v in rep.items():
                src = src.replace(k, v)
            return src
        traits, fns = "", ""
        if self.is_sde:
            traits = f"  static constexpr int kNoise = {self.noise_dim};\n"
            if self.noise_src is not None:
                traits += "  static constexpr bool kStateNoise = true;\n"
                fns = (f"  template <class R> static __device__ __forceinline__ R noise_prod(const P<R> &P_, R t, const R (&y)[{self.dim}], "
                       f"const R (&x)[{self.noise_dim}], int c) {{\n    [[maybe_unused]] const R *p = P_.p;\n    (void)t;\n    R gx[{self.dim}];\n"
                       f"{self.noise_src}\n    return gx[c];\n  }}\n")
            else:
                fns = ("  template <class R> static __device__ __forceinline__ R diffusion(const P<R> &P_, R t) {\n"
                       f"    [[maybe_unused]] const R *p = P_.p;\n    (void)t;\n    return (R)({self.diffusion_src});\n  }}\n")
        if self.event_srcs:
            traits += f"  static constexpr int kUserEvents = {len(self.event_srcs)};\n"
            cases = "".join(f"      case {i}: return (R)({e});\n" for i, e in enumerate(self.event_srcs))
            fns += (f"  template <class R> static __device__ __forceinline__ R event(const P<R> &P_, int i, R t, const R (&y)[{self.dim}]) {{\n"
                    f"    [[maybe_unused]] const R *p = P_.p;\n    (void)t;\n    switch (i) {{\n{cases}    }}\n    return R(0);\n  }}\n")
        defs = "".join(f"#define {k} {v}\n" for k, v in sorted(self.defines.items()))
        rep = {"@DEFINES@": defs + ("" if self.min_blocks is None else f"#define DFX_MIN_BLOCKS {self.min_blocks}"),
               "@ARGS_TRAIT@": "  static constexpr bool kPerTrajArgs = true;" if per_traj else "",
               "@PREAMBLE@": self.preamble, "@ID@": str(self.field_id_args if per_traj else self._id), "@DIM@": str(self.dim),
               "@SDE@": "true" if self.is_sde else "false",
               "@NP@": str(np_), "@NP1@": str(max(np_, 1)), "@NOISE_TRAITS@": traits, "@DRIFT@": self.drift_src, "@NOISE_FNS@": fns,
               "@REAL@": "double" if dtype_id == _lib.F64 else "float", "@SOLVER@": solver, "@LEVY@": str(int(levy))}
        src = _USER_TU
        for k, v in rep.items():
            src = src.replace(k, v)
        return src

    def ensure_kernel(self, dim, solver_id, dtype_id, levy, per_traj=False):
        """Compile (once) and load the kernel for this (solver, dtype, Levy area); called by `prepare`.  `per_traj`: the variant
        whose parameters `p[]` are read per trajectory from `args[N, n_params]` (registered under `field_id_args`)."""
        if dim != self.dim:
            raise ValueError(f"CudaField has state dimension {self.dim}, got y0 with d={dim}")
        if per_traj and not self.p:
            raise ValueError("CudaField: per-trajectory args need a functor with parameters (`params=[...]` gives their number)")
        key = (int(solver_id), int(dtype_id), int(levy), bool(per_traj))
        if key in self._ready:
            return
        inner = solver_id & ~_HALF
        if inner not in _SOLVER_CPP:
            raise ValueError(f"unknown solver id {solver_id}")
        if bool(levy) != self.is_sde:
            raise ValueError("CudaField: SDE solves need `diffusion=` or `noise=`; ODE solves must not have them")
        if self.wide and ((solver_id & _HALF) or inner in (7, 8)):
            raise ValueError("CudaField(wide=True): the warp-per-trajectory kernel runs the explicit RK tableaux "
                             "(Tsit5, Dopri5, Dopri8, Bosh3, Heun, Midpoint, Ralston)")
        if self.noise_src is not None and inner == 8:
            raise ValueError("ShARK is an additive-noise SRK (shark.py:10-30): the diffusion of this field depends on y")
        L = _lib.lib()
        fid = self.field_id_args if per_traj else self._id
        if not L.dfx_has_kernel(fid, self.dim, int(solver_id), int(dtype_id), int(levy)):
            from . import build
            path = build.build_user_field(f"{self._hash}_{solver_id:x}_{dtype_id}_{levy}" + ("_a" if per_traj else ""),
                                          self.source(solver_id, dtype_id, levy, per_traj))
            _lib.load_plugin(path)
            if not L.dfx_has_kernel(fid, self.dim, int(solver_id), int(dtype_id), int(levy)):
                raise RuntimeError(f"{path} was loaded but registered no launcher for this combination")
        self._ready.add(key)


_WIDE_TU = r"""// generated by diffrax_b200.fields.CudaField(wide=True) - do not edit
@DEFINES@
#include "wide_kernel.cuh"
namespace {
@PREAMBLE@
struct UserField {
  static constexpr int kId = @ID@;
  static constexpr int kDim = @DIM@;
  static constexpr bool kSde = false;
  static constexpr int kNumParams = @NP@;
@ARGS_TRAIT@
  template <class R> struct P { R p[@NP1@]; };
  template <class R> static P<R> make(const double *q, int n, const void *) {
    P<R> o;
    for (int i = 0; i < @NP1@; ++i) o.p[i] = (i < n && i < @NP@) ? (R)q[i] : R(0);
    return o;
  }
  // component i of vector_field(t, y, args); y is the whole state (shared memory, read-only)
  template <class R>
  static __device__ __forceinline__ R component(const P<R> &P_, R t, int i, const R *y) {
    [[maybe_unused]] const R *p = P_.p;
    [[maybe_unused]] constexpr int D = @DIM@;
    (void)t;
    R fi = R(0);
@DRIFT@
    return fi;
  }
};
[[maybe_unused]] constexpr bool kUseSde = UserField::kSde;
DFX_REGISTER_WIDE(@REAL@, UserField, @SOLVER@)
}  // namespace
"""


_BUILTIN_TU = r"""// generated by diffrax_b200.fields.ensure_builtin_kernel - do not edit
#include "launch.cuh"
namespace {
struct @FUNCTOR_DECL@;
using F = @FUNCTOR_NAME@;
DFX_REGISTER(@REAL@, F, @SOLVER@, @LEVY@)
}  // namespace
"""


def ensure_builtin_kernel(field, dim, solver_id, dtype_id, levy, bm_dim):
    """The shipped library prebuilds the (field, solver, dtype, Levy area) combinations of csrc/inst_*.cu; any other
    combination of a built-in functor with a solver - HalfSolver(Midpoint()), Euler on a vector OU process, LinearDecay with
    d = 5, ... - is instantiated here the first time it is asked for: one generated translation unit with ONE DFX_REGISTER,
    nvcc for sm_100a (~6 s), cached under lib/user/, dlopen.  `DFX_JIT=0` disables it (the solve then fails with "no kernel
    registered").  Returns True when a kernel is available afterwards."""
    import os
    L = _lib.lib()
    if L.dfx_has_kernel(field.field_id, dim, int(solver_id), int(dtype_id), int(levy)):
        return True
    functor = field.cpp_functor(dim, bm_dim)
    inner = solver_id & ~_HALF
    if functor is None or inner not in _SOLVER_CPP or os.environ.get("DFX_JIT", "1") == "0":
        return False
    if bool(levy) and not field.is_sde:
        return False
    if inner == 8 and (levy != _lib.LEVY_STLA or field.name == "gbm"):
        return False   # refused by the argument checks of the C ABI with the reference's messages (srk.py:391-395, shark.py:10-30)
    solver = _SOLVER_CPP[inner]
    if solver_id & _HALF:
        solver = f"::dfx::HalfOf<{solver}>"
    # cpp_functor: a type name, or "Name; struct Name : Base { ... }" for a derived functor (MLP sizes with ids of their own)
    name, _, decl = functor.partition("; struct ")
    src = _BUILTIN_TU.replace("@FUNCTOR_DECL@", decl if decl else "DfxUnused {}").replace("@FUNCTOR_NAME@", name)
    src = src.replace("@REAL@", "double" if dtype_id == _lib.F64 else "float")
    src = src.replace("@SOLVER@", solver).replace("@LEVY@", str(int(levy)))
    from . import build
    tag = f"builtin_{field.field_id}_{dim}_{int(solver_id):x}_{int(dtype_id)}_{int(levy)}"
    _lib.load_plugin(build.build_user_field(tag, src))
    return bool(L.dfx_has_kernel(field.field_id, dim, int(solver_id), int(dtype_id), int(levy)))
