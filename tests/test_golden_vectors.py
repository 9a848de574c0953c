"""Committed golden vectors (tests/golden/vectors.npz, made by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), the host-simulated device code must reproduce them
(CPU), and the GPU must reproduce them bit for bit (gpu marker)."""
import os

import numpy as np
import pytest

from cascade_b200 import api
from cascade_b200.format import Balls

V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.npz"))


def g(name):
    m = V[name + "_m"]
    return Balls(m.shape[1], m, V[name + "_t"])


def run_all(drv):
    r = api.acb_ode_solve_fuchs_batch(g("kat_ode"), g("kat_init"), 2, 2, 5, 1, True, driver=drv)
    assert r.same_as(g("kat_res"))
    assert [float(r.mid(2 * k)) for k in range(5)] == list(V["readme_mid"])  # reference README.md:72-77
    assert api.acb_ode_solve_fuchs_batch(g("bessel_ode"), g("bessel_init"), 2, 2, 24, 6, True, driver=drv).same_as(g("bessel_res"))
    assert api.acb_ode_solve_fuchs_batch(g("gen_ode"), g("gen_init"), 3, 4, 20, 3, False, driver=drv).same_as(g("gen_res"))
    assert api.acb_ode_solve_frobenius1_batch(g("frob_ode"), g("frob_rho"), 2, 2, 16, 1, True, True, driver=drv).same_as(g("frob_res"))
    assert api.acb_poly_evaluate_batch(g("horner_f"), g("horner_pts"), 2, driver=drv).same_as(g("horner_out"))
    gens = api.acb_ode_solve_frobenius_batch(g("frob_ode"), g("frob_rho"), 2, 0, 2, 2, 10, 1, True, True, driver=drv)
    assert gens.same_as(g("frobg_gens"))
    assert api.acb_ode_solve_frobenius1_batch(g("frob_ode"), g("frob_rho"), 2, 2, 12, 1, True, True, driver=drv,
                                              rhs=g("frob_rhs")).same_as(g("frob_rhs_res"))
    assert api.acb_ode_solution_op_batch(api.SOL_UPDATE, g("frob_rho"), gens, 1, 2, 2, g("sol_f"), driver=drv)[0].same_as(g("sol_updated"))
    assert api.indicial_polynomial_batch(g("frob_ode"), 2, 2, 1, 3, 1, True, driver=drv).same_as(g("indicial"))
    assert api.acb_exp_batch(g("trans_x"), driver=drv).same_as(g("trans_exp"))
    assert api.acb_log_batch(g("trans_x"), driver=drv).same_as(g("trans_log"))
    assert api.acb_ode_solution_evaluate_batch(gens, 2, g("frob_rho"), g("horner_pts"), driver=drv).same_as(g("eval_out"))
    assert api.acb_ode_apply_batch(g("frob_ode"), g("apply_in"), 2, 2, 1, True, driver=drv)[0].same_as(g("apply_out"))


def test_oracle_reproduces_golden(oracle):
    assert oracle.fuchs_batch(2, 2, 5, g("kat_ode"), g("kat_init"), 1, True)[1].same_as(g("kat_res"))
    assert oracle.fuchs_batch(2, 2, 24, g("bessel_ode"), g("bessel_init"), 6, True)[1].same_as(g("bessel_res"))
    assert oracle.fuchs_batch(3, 4, 20, g("gen_ode"), g("gen_init"), 3, False)[1].same_as(g("gen_res"))
    assert oracle.frobenius1_batch(2, 2, 16, g("frob_ode"), g("frob_rho"), 1, True).same_as(g("frob_res"))
    assert oracle.horner_batch(g("horner_f"), g("horner_pts"), 2).same_as(g("horner_out"))
    gens = oracle.frobenius_general_batch(2, 2, 10, g("frob_ode"), g("frob_rho"), 1, True, True, 2, 2)
    assert gens.same_as(g("frobg_gens"))
    assert oracle.ew_trans("exp", g("trans_x")).same_as(g("trans_exp")) and oracle.ew_trans("log", g("trans_x")).same_as(g("trans_log"))
    assert oracle.solution_evaluate(gens, 2, g("frob_rho"), 2, 0, g("horner_pts")).same_as(g("eval_out"))
    assert oracle.ode_apply(2, 2, g("frob_ode"), g("apply_in")).same_as(g("apply_out"))


def test_device_logic_reproduces_golden(sim_driver):
    run_all(sim_driver)


@pytest.mark.gpu
def test_gpu_reproduces_golden(gpu_driver):
    run_all(gpu_driver)
