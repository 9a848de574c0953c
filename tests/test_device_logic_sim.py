"""CPU checks of the DEVICE code's logic: the per-thread kernel bodies compiled for the host
(tests/host_sim) against the oracle, bit for bit.  No GPU needed; the GPU runs the same cases
through the C ABI in tests/test_gpu_parity.py."""
import pytest

from tests import cases


@pytest.mark.parametrize("L,n", [(4, 300), (32, 200), (64, 40), (128, 24)])
def test_elementwise(oracle, sim_driver, L, n):
    cases.case_elementwise(oracle, sim_driver, L, n, seed=100 + L)


def test_readme_kat(oracle, sim_driver):
    cases.case_readme_kat(oracle, sim_driver)


@pytest.mark.parametrize("L", [4, 32])
def test_fuchs_random(oracle, sim_driver, L):
    cases.case_fuchs_random(oracle, sim_driver, L, seed=7 + L)


@pytest.mark.parametrize("L", [4, 32])
def test_fuchs_random_shared(oracle, sim_driver, L):
    cases.case_fuchs_random_shared(oracle, sim_driver, L, seed=31 + L, trials=6)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_fuchs_shared_adversarial(oracle, sim_driver, L):
    cases.case_fuchs_shared_adversarial(oracle, sim_driver, L, seed=77 + L, batch=12)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_fuchs_integer_operators(oracle, sim_driver, L):
    cases.case_fuchs_integer_operators(oracle, sim_driver, L, seed=5 + L)


def test_fuchs_bessel_shared(oracle, sim_driver):
    cases.case_fuchs_bessel_shared(oracle, sim_driver, 32, seed=5, n=40, batch=3)


def test_fuchs_hypgeom(oracle, sim_driver):
    cases.case_fuchs_hypgeom(oracle, sim_driver, 32, seed=1)


@pytest.mark.parametrize("L", [4, 32, 128])
def test_frobenius(oracle, sim_driver, L):
    cases.case_frobenius(oracle, sim_driver, L, seed=9, n=12 if L == 128 else 20, batch=2)


@pytest.mark.parametrize("L", [4, 128])
def test_frobenius_in_coefficient_ranges(monkeypatch, oracle, sim_driver, L):
    """the persistent 4096-bit Frobenius kernel cuts a homogeneous solve into work units of a few coefficients
    (frobenius1_thread with nu_from .. nu_to, earlier coefficients read back from the result array)"""
    monkeypatch.setenv("CBSIM_FROB_RANGE", "3")
    cases.case_frobenius(oracle, sim_driver, L, seed=39, n=10, batch=2)


@pytest.mark.parametrize("L", [64, 128])
def test_integer_like_operands_take_the_two_row_product(oracle, sim_driver, L):
    """2048 / 4096 bits: a factor whose limbs below the top two are zero (rho + k with rho = 0, g_0 = 1, an integer
    exponent) goes through product_short_y; the series must still agree with the oracle bit for bit, also on the
    cooperating lanes."""
    from tests.host_sim import driver as simdrv
    before = simdrv.path_count(6)
    cases.case_frobenius(oracle, sim_driver, L, seed=19, n=8, batch=2)
    cases.case_frobenius_singleton(oracle, sim_driver, L, seed=23, trials=1)
    assert simdrv.path_count(6) > before, "the two-row product was never taken"
    coop = simdrv.SimDriver(coop=True)
    mid = simdrv.path_count(6)
    cases.case_frobenius(oracle, coop, L, seed=29, n=6, batch=2)
    assert simdrv.path_count(6) > mid


def test_frobenius_singleton(oracle, sim_driver):
    cases.case_frobenius_singleton(oracle, sim_driver, 4, seed=21)


@pytest.mark.parametrize("L", [4, 64])
def test_horner(oracle, sim_driver, L):
    cases.case_horner(oracle, sim_driver, L, seed=13)


@pytest.mark.parametrize("L", [32, 64, 128])
def test_horner_cancellation(oracle, sim_driver, L):
    cases.case_horner_cancellation(oracle, sim_driver, L, seed=113 + L, length=12 if L == 128 else 24, npts=12)


@pytest.mark.parametrize("L", [4, 32])
def test_continuation(oracle, sim_driver, L):
    cases.case_continuation(oracle, sim_driver, L, seed=3 + L)


@pytest.mark.parametrize("L", [4, 32])
def test_continuation_series_and_singular_corners(oracle, sim_driver, L):
    cases.case_continuation_series(oracle, sim_driver, L, seed=5 + L)


@pytest.mark.parametrize("L", [4, 32])
def test_radius_and_truncation(oracle, sim_driver, L):
    cases.case_radius_and_truncation(oracle, sim_driver, L, seed=11 + L, batch=4)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_ode_shift(oracle, sim_driver, L):
    cases.case_ode_shift(oracle, sim_driver, L, seed=7 + L, batch=4)


@pytest.mark.parametrize("L", [4, 32])
def test_taylor_shift(oracle, sim_driver, L):
    cases.case_taylor_shift(oracle, sim_driver, L, seed=9 + L, length=20)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_exponent_overflow(oracle, sim_driver, L):
    cases.case_exponent_overflow(oracle, sim_driver, L)


@pytest.mark.parametrize("L", [4, 32])
def test_frobenius_general(oracle, sim_driver, L):
    cases.case_frobenius_general(oracle, sim_driver, L, seed=41 + L, trials=4 if L == 4 else 2)


@pytest.mark.parametrize("L", [4, 32])
def test_frobenius_general_sweep(oracle, sim_driver, L):
    cases.case_frobenius_general_sweep(oracle, sim_driver, L, seed=43 + L, n=8 if L == 4 else 6, batch=3)


@pytest.mark.parametrize("L", [4, 32])
def test_solution_ops(oracle, sim_driver, L):
    cases.case_solution_ops(oracle, sim_driver, L, seed=47 + L, batch=2)


@pytest.mark.parametrize("L", [4, 32])
def test_indicial_polynomial(oracle, sim_driver, L):
    cases.case_indicial_polynomial(oracle, sim_driver, L, seed=53 + L, trials=3 if L == 4 else 1)


@pytest.mark.parametrize("L", [4, 32])
def test_frobenius_rhs(oracle, sim_driver, L):
    cases.case_frobenius_rhs(oracle, sim_driver, L, seed=59 + L, n=10)


def test_reduced_radix_products(oracle, sim_driver, monkeypatch):
    """The reduced-radix fused dot product (cb_rr.cuh, opt-in with CB200_RR=1) must give the same balls:
    dense, shared-operator and adversarial cases through the recurrence kernel's logic."""
    monkeypatch.setenv("CB200_RR", "1")
    cases.case_fuchs_bessel_shared(oracle, sim_driver, 32, seed=5, n=40, batch=3)
    cases.case_fuchs_random_shared(oracle, sim_driver, 32, seed=63, trials=6)
    cases.case_fuchs_shared_adversarial(oracle, sim_driver, 32, seed=109, batch=12)


@pytest.mark.parametrize("L", [4, 32])
def test_exp_log(oracle, sim_driver, L):
    cases.case_exp_log(oracle, sim_driver, L, seed=61 + L, n=20 if L == 4 else 16)


@pytest.mark.parametrize("L", [4, 32])
def test_solution_evaluate(oracle, sim_driver, L):
    cases.case_solution_evaluate(oracle, sim_driver, L, seed=67 + L, npts=4, cap=8 if L == 32 else 12)


@pytest.mark.parametrize("L", [4, 32])
def test_ode_apply(oracle, sim_driver, L):
    cases.case_ode_apply(oracle, sim_driver, L, seed=71 + L)


def test_fuchs_coefficient_ranges(oracle, sim_driver):
    """Solving in coefficient ranges (cb200_fuchs_batch_range_dev: the recurrence resumes from res) gives the
    same balls as one call."""
    import numpy as np
    from cascade_b200 import api
    from cascade_b200.format import entry_major_to_instance_major, instance_major_to_entry_major, pack_entries, random_balls, unpack_entries
    rng = np.random.default_rng(5)
    L, n, batch = 32, 45, 3
    ode = cases.bessel_shifted(oracle, rng, L)
    init = random_balls(rng, batch * 4, L, -1, 0, "zero")
    rc, want = oracle.fuchs_batch(2, 2, n, ode, init, batch, True)
    res = oracle._series_with_init(init, batch, n, 2)
    res_m, res_t = pack_entries(instance_major_to_entry_major(res, batch, n + 1), n + 1, 2 * batch)
    ode_m, ode_t = api._ode_to_device(ode, 2, 2, batch, True)
    for lo, hi in ((0, 7), (8, 8), (9, 30), (31, 45)):
        assert sim_driver.fuchs_range(L, 2, 2, -2, lo, hi, batch, ode_m, ode_t, True, res_m, res_t) == 0
    got = entry_major_to_instance_major(unpack_entries(L, res_m, res_t), batch, n + 1)
    cases.assert_same(want, got, "fuchs in coefficient ranges")


@pytest.mark.parametrize("L", [4, 32])
def test_frobenius_general_edge(oracle, sim_driver, L):
    cases.case_frobenius_general_edge(oracle, sim_driver, L)


@pytest.mark.parametrize("L", [64, 128])
def test_cooperating_lanes(monkeypatch, oracle, L):
    """cb_coop.cuh: four lanes per instance (the products of a complex multiplication on four lanes, real and imaginary
    parts of everything else on two), simulated lane by lane and phase by phase: Frobenius M = 1 and Horner, including
    the cancellation case that needs the exact full-width path on one lane's pair of hybrid columns."""
    from tests.host_sim.driver import SimDriver
    drv = SimDriver(coop=True)
    cases.case_frobenius(oracle, drv, L, seed=29, n=12 if L == 64 else 9, batch=5)
    cases.case_horner(oracle, drv, L, seed=23, length=14, npts=5)
    cases.case_horner_cancellation(oracle, drv, L, seed=213 + L, length=10, npts=12)
    # the shared-operator Fuchs recurrence on four lanes per instance (CBSIM_COOP routes the simulation through it)
    monkeypatch.setenv("CBSIM_COOP", "1")
    cases.case_fuchs_bessel_shared(oracle, drv, L, seed=5, n=40 if L == 64 else 24, batch=5)
    cases.case_fuchs_random_shared(oracle, drv, L, seed=41 + L, trials=3, batch=4)
    cases.case_fuchs_shared_adversarial(oracle, drv, L, seed=87 + L, batch=6)
    cases.case_continuation(oracle, drv, L, seed=23 + L, steps=3, n=20, batch=2)
