"""Generates tests/golden/vectors.npz: small input/output vectors of the hot path.

The reference (Cascade + Arb) cannot be built in this image, so the outputs are produced by the
CPU oracle (oracle/), whose semantics are pinned by tests/test_oracle_golden.py; the one literal
answer the reference publishes (README.md:72-77) is stored as `readme_mid` and checked as such.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys
from fractions import Fraction as F

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cascade_b200.format import Balls, from_complex, random_balls  # noqa: E402
from oracle import cbo  # noqa: E402
from tests.cases import bessel_shifted  # noqa: E402


def main():
    rng = np.random.default_rng(0x601D)
    out = {}

    def put(name, b: Balls):
        out[name + "_m"], out[name + "_t"] = b.mant, b.meta

    L = 32
    ode = cbo.example("legendre", L, 4)
    init = from_complex([F(3, 8), 0], L)
    put("kat_ode", ode)
    put("kat_init", init)
    put("kat_res", cbo.fuchs_batch(2, 2, 5, ode, init, 1, True)[1])
    out["readme_mid"] = np.array([0.375, 0.0, -3.75, 0.0, 4.375])

    ode = bessel_shifted(cbo, rng, L)
    init = random_balls(rng, 6 * 4, L, -1, 0, "zero")
    put("bessel_ode", ode)
    put("bessel_init", init)
    put("bessel_res", cbo.fuchs_batch(2, 2, 24, ode, init, 6, True)[1])

    L4 = 4
    odes = random_balls(rng, 3 * 20 * 2, L4, -4, 4, "ulp", short_frac=0.2)  # order 3, degree 4
    init = random_balls(rng, 3 * 3 * 2, L4, -3, 3, "ulp")
    put("gen_ode", odes)
    put("gen_init", init)
    put("gen_res", cbo.fuchs_batch(3, 4, 20, odes, init, 3, False)[1])

    par = random_balls(rng, 6, L, -1, 1, "zero")
    hyp = cbo.example("hypgeom", L, 0, par)
    rho = random_balls(rng, 2, L, -1, 1, "ulp")
    put("frob_ode", hyp)
    put("frob_rho", rho)
    put("frob_res", cbo.frobenius1_batch(2, 2, 16, hyp, rho, 1, True))

    f = random_balls(rng, 2 * 30, L, -3, 3, "ulp")
    pts = random_balls(rng, 2 * 5, L, -2, 0, "zero")
    put("horner_f", f)
    put("horner_pts", pts)
    put("horner_out", cbo.horner_batch(f, pts, 2))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
