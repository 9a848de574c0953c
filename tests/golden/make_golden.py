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
    # logarithmic Frobenius (M = 2), right-hand side, bookkeeping steps, exp/log, full evaluate, residual
    gens = cbo.frobenius_general_batch(2, 2, 10, hyp, rho, 1, True, True, 2, 2)
    put("frobg_gens", gens)
    rhs = random_balls(rng, 2 * 5, L, -2, 2, "ulp")
    put("frob_rhs", rhs)
    put("frob_rhs_res", cbo.frobenius1_rhs_batch(2, 2, 12, hyp, rho, rhs, 1, True))
    fpoly = random_balls(rng, 2 * 7, L, -3, 3, "ulp")
    put("sol_f", fpoly)
    put("sol_updated", cbo.solution_op("update", rho, gens, 1, 2, 2, fpoly)[0])
    put("indicial", cbo.indicial_polynomial(2, 2, hyp, 1, 3))
    x = random_balls(rng, 2 * 6, L, -3, 3, "ulp")
    put("trans_x", x)
    put("trans_exp", cbo.ew_trans("exp", x))
    put("trans_log", cbo.ew_trans("log", x))
    put("eval_out", cbo.solution_evaluate(gens, 2, rho, 2, 0, pts))
    series = cbo.fuchs_batch(2, 2, 12, hyp, from_complex([1], L), 1, True)[1]
    put("apply_in", series)
    put("apply_out", cbo.ode_apply(2, 2, hyp, series))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
