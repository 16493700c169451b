#!/usr/bin/env python
"""Extract the Butcher-tableau / interpolant coefficient VALUES from the reference.

Test infrastructure, run in the authoring container only (the reference tree does
not exist on the GPU box).  The reference writes its coefficients as Python
expressions (`35 / 384 - 1951 / 21600`, 88-digit decimal literals, ...) that are
evaluated in IEEE double by CPython; the integrator must use *those* doubles bit
for bit (SURVEY.md §7 "do not difference in fp32", App. C).  This script parses
the reference files with `ast` (nothing is imported or executed from the
reference - jax/equinox are absent), evaluates only the numeric literal
expressions with CPython float arithmetic, and writes the resulting values as
C99 hex-float strings to `tests/golden/tableaux.json`.

`tools/gen_tableaux.py` turns that JSON into the generated headers used by the
oracle and by the CUDA kernels; `tests/test_tableaux.py` re-checks the values
against the Butcher order conditions, so a transcription slip cannot hide.

Sources (file:line in /root/reference/diffrax):
  _solver/tsit5.py:18-100   _solver/dopri5.py:11-47   _solver/dopri8.py:18-292
  _solver/heun.py:12-17     _solver/shark.py:10-30    _solver/bosh3.py, midpoint.py,
  ralston.py, euler.py (free once the kernel is tableau-generic)
"""
import ast
import json
import math
import operator
import pathlib
import sys

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/diffrax")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "tableaux.json"

_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul,
        ast.Div: operator.truediv, ast.Pow: operator.pow}


def ev(node):
    """Evaluate a numeric literal expression exactly as CPython would."""
    if isinstance(node, ast.Constant):
        if node.value is None:
            return None
        assert isinstance(node.value, (int, float)), node.value
        return node.value
    if isinstance(node, ast.UnaryOp):
        v = ev(node.operand)
        return -v if isinstance(node.op, ast.USub) else +v
    if isinstance(node, ast.BinOp):
        return _BIN[type(node.op)](ev(node.left), ev(node.right))
    if isinstance(node, (ast.List, ast.Tuple)):
        return [ev(e) for e in node.elts]
    if isinstance(node, ast.Call):
        fn = ast.unparse(node.func)
        if fn in ("np.array", "jnp.array", "np.asarray"):
            return ev(node.args[0])
        if fn == "np.sqrt" or fn == "math.sqrt":
            return math.sqrt(ev(node.args[0]))
    raise ValueError(f"unsupported node: {ast.dump(node)[:200]}")


def find_assign(tree, name):
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(
            isinstance(t, ast.Name) and t.id == name for t in node.targets
        ):
            return node.value
        if isinstance(node, ast.AnnAssign) and isinstance(node.target, ast.Name) \
                and node.target.id == name and node.value is not None:
            return node.value
    raise KeyError(name)


def kwargs(call):
    return {k.arg: ev(k.value) for k in call.keywords}


def hx(x):
    if isinstance(x, list):
        return [hx(e) for e in x]
    return float(x).hex()


def main():
    out = {}
    for fname, var in [("tsit5", "_tsit5_tableau"), ("dopri5", "_dopri5_tableau"),
                       ("dopri8", "_dopri8_tableau"), ("heun", "_heun_tableau"),
                       ("bosh3", "_bosh3_tableau"), ("midpoint", "_midpoint_tableau"),
                       ("ralston", "_ralston_tableau"), ("euler", None)]:
        if var is None:
            continue
        tree = ast.parse((REF / "_solver" / f"{fname}.py").read_text())
        kw = kwargs(find_assign(tree, var))
        out[fname] = {k: hx(v) for k, v in kw.items() if v is not None}
    # Dopri5 quartic-interpolant mid-point weights, Dopri8 dense polynomial.
    tree = ast.parse((REF / "_solver" / "dopri5.py").read_text())
    out["dopri5"]["c_mid"] = hx(ev(find_assign(tree, "c_mid")))
    tree = ast.parse((REF / "_solver" / "dopri8.py").read_text())
    out["dopri8"]["eval_coeffs"] = hx(ev(find_assign(tree, "eval_coeffs")))
    # ShARK additive-noise SRK tableau.
    tree = ast.parse((REF / "_solver" / "shark.py").read_text())
    tab = kwargs_partial(find_assign(tree, "_tab"))
    out["shark"] = {k: hx(v) for k, v in tab.items()}
    for nm in ("_coeffs_w", "_coeffs_hh"):
        kw = kwargs(find_assign(tree, nm))
        out["shark"][nm.strip("_")] = {k: hx(v) for k, v in kw.items()}
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


def kwargs_partial(call):
    res = {}
    for k in call.keywords:
        try:
            v = ev(k.value)
        except ValueError:
            continue  # names (coeffs_w=_coeffs_w) are extracted separately
        if v is not None:
            res[k.arg] = v
    return res


if __name__ == "__main__":
    main()
