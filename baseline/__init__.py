"""Live-reference arm: run the UNMODIFIED Diffrax (``/root/reference`` or ``baseline/_ref``) on the same inputs.

TEST / BENCH INFRASTRUCTURE ONLY - imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py --impl reference``
(and by ``baseline/gen_golden.py``); the product (``diffrax_b200``) never imports it.

Diffrax is pure Python on jax + equinox + lineax + optimistix (+ jaxtyping, wadler_lindig).  None of those is installed
in the authoring container and ``/opt/wheelhouse`` holds no wheel for them, so ``pip install --no-index --find-links
/opt/wheelhouse --target baseline/_ref /root/reference`` fails at dependency resolution and ``--no-deps`` yields a package
that cannot be imported (recorded in DESIGN.md section 7).  ``probe()`` therefore answers "is a live Diffrax importable
right now?" at run time, here and on the GPU box:

    ok, why = baseline.probe()

When it is, ``baseline.solve(case)`` drives ``diffrax.diffeqsolve`` (``_integrate.py:888``) under ``jax.vmap`` exactly the
way ``test/helpers.py:136-186`` and ``test/test_vmap.py:28-39`` do, on the JAX CPU backend with x64 enabled, and
  * ``tests/test_live_reference.py`` pins the oracle (and, on a GPU box, the CUDA path) against it at the north-star
    tolerances,
  * ``bench.py --impl reference`` times it (``cpu_baseline.kind == "reference"``),
  * ``python baseline/gen_golden.py`` writes ``tests/golden/diffrax_golden.npz`` + the versions it came from.
When it is not, callers fall back to the C oracle (``kind == "port"``) and say so.
"""
from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INSTALL = os.path.join(HERE, "_ref")          # pip --target directory (git-ignored, travels with gpurun)
REF_SOURCE = "/root/reference"                   # read-only checkout; absent on the GPU box

_state = {"probed": False, "ok": False, "why": "", "jax": None, "diffrax": None}


def probe():
    """(ok, reason).  Imports jax and diffrax at most once; never raises."""
    if _state["probed"]:
        return _state["ok"], _state["why"]
    _state["probed"] = True
    os.environ.setdefault("JAX_PLATFORMS", "cpu")          # the reference arm is the JAX-CPU vmap path (BASELINE.json)
    try:
        jax = importlib.import_module("jax")
    except Exception as e:  # noqa: BLE001
        _state["why"] = f"jax not importable ({type(e).__name__}: {e})"
        return False, _state["why"]
    added = []
    for p in (REF_INSTALL, REF_SOURCE):
        if os.path.isdir(os.path.join(p, "diffrax")) and p not in sys.path:
            sys.path.append(p)
            added.append(p)
    try:
        diffrax = importlib.import_module("diffrax")
    except Exception as e:  # noqa: BLE001
        for p in added:
            sys.path.remove(p)
        _state["why"] = f"diffrax not importable ({type(e).__name__}: {e})"
        return False, _state["why"]
    jax.config.update("jax_enable_x64", True)               # fp64 configs; fp32 cases pass float32 arrays explicitly
    _state.update(ok=True, jax=jax, diffrax=diffrax,
                  why=f"diffrax {getattr(diffrax, '__version__', '?')} on jax {jax.__version__} "
                      f"({jax.default_backend()}), threefry_partitionable={jax.config.jax_threefry_partitionable}")
    return True, _state["why"]


def versions():
    ok, why = probe()
    if not ok:
        return {"available": False, "why": why}
    out = {"available": True, "jax": _state["jax"].__version__,
           "diffrax": getattr(_state["diffrax"], "__version__", "?"),
           "jax_threefry_partitionable": bool(_state["jax"].config.jax_threefry_partitionable),
           "backend": _state["jax"].default_backend()}
    for m in ("jaxlib", "equinox", "lineax", "optimistix"):
        try:
            out[m] = importlib.import_module(m).__version__
        except Exception:  # noqa: BLE001
            out[m] = None
    return out


def solve(case, *, jit=True):
    """Run one oracle-style case (the dicts of tests/golden/make_golden.py and bench.workload) through live Diffrax."""
    ok, why = probe()
    if not ok:
        raise RuntimeError(f"live Diffrax unavailable: {why}")
    from . import ref_cases
    return ref_cases.solve(_state["jax"], _state["diffrax"], case, jit=jit)


def compiled(case):
    """(fn, args): the jitted vmapped solve and its device-resident arguments, for timing (`fn(*args)` then block)."""
    ok, why = probe()
    if not ok:
        raise RuntimeError(f"live Diffrax unavailable: {why}")
    from . import ref_cases
    return ref_cases.compiled(_state["jax"], _state["diffrax"], case)
