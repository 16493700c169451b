#!/usr/bin/env python
"""Golden vectors from the LIVE reference.

    python baseline/gen_golden.py            # needs `import jax, diffrax` to work (baseline.probe())

Runs every case of tests/golden/make_golden.py (the same seeded inputs the oracle's goldens use) through
``diffrax.diffeqsolve`` under ``jax.vmap`` on the JAX CPU backend and writes

    tests/golden/diffrax_golden.npz      <case>/ys, ts, stats, result  +  prng/* and vbt/* words from jax.random / the tree
    tests/golden/diffrax_golden.json     the versions (jax, jaxlib, diffrax, equinox, ...) and jax_threefry_partitionable

``tests/test_live_reference.py`` compares the oracle (CPU) and the CUDA path (GPU box) with that file whenever it
exists; until it has been generated the repository's parity status stays "unpinned against a live Diffrax"
(oracle/oracle.h).  This script cannot run in the authoring container (no jax wheel, no network).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import baseline
    ok, why = baseline.probe()
    if not ok:
        print(f"live Diffrax unavailable: {why}\nnothing written.")
        return 2
    import make_golden
    import jax
    import jax.numpy as jnp
    import jax.random as jr
    import diffrax

    out = {}
    skipped = []
    for name, kw in list(make_golden.CASES.items()) + list(make_golden.CASES_EXTRA.items()):
        if kw.get("event") is not None or kw.get("store_rejected_steps"):
            skipped.append(name)        # registered-functor events have no generic JAX twin here
            continue
        try:
            r = baseline.solve(kw)
        except Exception as e:  # noqa: BLE001
            skipped.append(f"{name}: {type(e).__name__}: {e}")
            continue
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, r["stats"][:3].tolist(), "failed:", int((r["result"] != 0).sum()))

    # PRNG / Brownian words straight from jax.random and diffrax.VirtualBrownianTree
    keys = np.asarray(make_golden.keys, np.uint32)
    out["prng/keys"] = keys
    part0 = bool(jax.config.jax_threefry_partitionable)
    for part in (1, 0):
        jax.config.update("jax_threefry_partitionable", bool(part))
        jk = jr.wrap_key_data(jnp.asarray(keys))
        out[f"prng/split3_part{part}"] = np.asarray(jr.key_data(jax.vmap(lambda k: jr.split(k, 3))(jk)))[:16]
        for dt_, tag in ((jnp.float64, "f64"), (jnp.float32, "f32")):
            out[f"prng/normal_{tag}_part{part}"] = np.asarray(jax.vmap(lambda k: jr.normal(k, (), dt_))(jk))
            out[f"prng/normal3_{tag}_part{part}"] = np.asarray(jax.vmap(lambda k: jr.normal(k, (3,), dt_))(jk))
    jax.config.update("jax_threefry_partitionable", True)
    jk = jr.wrap_key_data(jnp.asarray(keys))
    for lv, la in (("bi", diffrax.BrownianIncrement), ("stla", diffrax.SpaceTimeLevyArea)):
        for dt_, tag in ((jnp.float64, "f64"), (jnp.float32, "f32")):
            for shape, stag in (((), ""), ((3,), "_m3")):
                def ev(k):
                    bm = diffrax.VirtualBrownianTree(0.0, 1.0, 2.0 ** -8, jax.ShapeDtypeStruct(shape, dt_), k, levy_area=la)
                    x = bm.evaluate(jnp.asarray(0.3, dt_), jnp.asarray(0.7, dt_), use_levy=True)
                    return x.W, (x.H if lv == "stla" else jnp.zeros_like(x.W))
                W, H = jax.vmap(ev)(jk)
                out[f"vbt/{lv}_{tag}{stag}_W"] = np.asarray(W)
                out[f"vbt/{lv}_{tag}{stag}_H"] = np.asarray(H)
    jax.config.update("jax_threefry_partitionable", part0)

    dst = os.path.join(ROOT, "tests", "golden", "diffrax_golden.npz")
    np.savez_compressed(dst, **out)
    meta = baseline.versions()
    meta["skipped"] = skipped
    with open(os.path.join(ROOT, "tests", "golden", "diffrax_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", dst, os.path.getsize(dst), "bytes;", meta)
    return 0


if __name__ == "__main__":
    sys.exit(main())
