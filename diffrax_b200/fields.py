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


class LinearDecay(Field):
    """dy/dt = -lam * y  (test/test_integrate.py:60, test_saveat_solution.py:29-41); d in {1,2,3}."""
    name = "decay"

    def __init__(self, lam=1.0):
        self.lam = float(lam)

    def params(self):
        return [self.lam]


class LotkaVolterra(Field):
    """benchmarks/lotka_volterra.py:13-20: [a x + b x y, c y + d x y]."""
    name, dim = "lotka_volterra", 2

    def __init__(self, a=1.5, b=-1.0, c=-3.0, d=1.0):
        self.p = [float(a), float(b), float(c), float(d)]

    def params(self):
        return self.p


class Lorenz(Field):
    """Lorenz-63: [sigma (y-x), x (rho - z) - y, x y - beta z]."""
    name, dim = "lorenz", 3

    def __init__(self, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
        self.p = [float(sigma), float(rho), float(beta)]

    def params(self):
        return self.p


class CR3BP(Field):
    """Planar circular restricted three-body problem, rotating frame, state (x, y, vx, vy)."""
    name, dim = "cr3bp", 4

    def __init__(self, mu=0.012277471):
        self.mu = float(mu)

    def params(self):
        return [self.mu]


class ForcedOscillator(Field):
    """y0' = y1, y1' = -w0^2 y0 + A sin(w t)."""
    name, dim = "forced_osc", 2

    def __init__(self, w0sq=1.0, amp=1.0, w=2.0):
        self.p = [float(w0sq), float(amp), float(w)]

    def params(self):
        return self.p


class VanDerPol(Field):
    name, dim = "vdp", 2

    def __init__(self, mu=1.0):
        self.mu = float(mu)

    def params(self):
        return [self.mu]


class OrnsteinUhlenbeck(Field):
    """dy = theta (mu - y) dt + (sigma + sigma_t t) dW (additive noise, optionally growing linearly in time).  With a state of
    dimension m > 1 the components are independent and driven by VirtualBrownianTree(shape=(m,)) (diagonal diffusion)."""
    name, dim, is_sde = "ou", 0, True

    def __init__(self, theta=1.0, mu=0.0, sigma=0.5, sigma_t=0.0):
        self.p = [float(theta), float(mu), float(sigma)] + ([float(sigma_t)] if sigma_t else [])

    def params(self):
        return self.p


class GeometricBrownianMotion(Field):
    """dy = mu y dt + sigma y dW: state-dependent (multiplicative) diagonal diffusion with a scalar Brownian motion -
    ``ControlTerm(lambda t, y, args: sigma * y, VirtualBrownianTree(..., shape=(), ...))``.  Euler / Heun (Heun converges to the
    Stratonovich solution); kernels for state dimension 1 and 2."""
    name, dim, is_sde = "gbm", 0, True

    def __init__(self, mu=0.1, sigma=0.2):
        self.p = [float(mu), float(sigma)]

    def params(self):
        return self.p


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


class MLP(Field):
    """Neural-ODE vector field: ``eqx.nn.MLP(in=d, out=d, width, depth=2, activation=softplus, final_activation=tanh)``
    (docs/examples/neural_ode.ipynb cell 5).  ``layers`` is ``[(W1, b1), (W2, b2), (W3, b3)]`` with eqx ``Linear``
    weight shapes ``(out, in)``.  Built-in kernels exist for d=4, width=128, fp32."""
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
