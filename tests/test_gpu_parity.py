"""GPU parity tests proper: libcascade_b200.so through its C ABI (host buffers in and out)
against the oracle, bit for bit, plus the reference's own acceptance properties."""
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,n", [(4, 1000), (32, 600), (64, 200), (128, 100)])
def test_elementwise(oracle, gpu_driver, L, n):
    cases.case_elementwise(oracle, gpu_driver, L, n, seed=200 + L)


def test_readme_kat(oracle, gpu_driver):
    cases.case_readme_kat(oracle, gpu_driver)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_fuchs_random(oracle, gpu_driver, L):
    cases.case_fuchs_random(oracle, gpu_driver, L, seed=17 + L, trials=5 if L < 128 else 2, batch=37)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_fuchs_random_shared(oracle, gpu_driver, L):
    cases.case_fuchs_random_shared(oracle, gpu_driver, L, seed=41 + L, trials=6 if L < 128 else 3, batch=45)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_fuchs_shared_adversarial(oracle, gpu_driver, L):
    cases.case_fuchs_shared_adversarial(oracle, gpu_driver, L, seed=87 + L, batch=70)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_fuchs_integer_operators(oracle, gpu_driver, L):
    cases.case_fuchs_integer_operators(oracle, gpu_driver, L, seed=15 + L)


@pytest.mark.parametrize("L", [32, 64, 128])
def test_fuchs_bessel_shared(oracle, gpu_driver, L):
    cases.case_fuchs_bessel_shared(oracle, gpu_driver, L, seed=5, n=60 if L < 128 else 30, batch=300)


def test_fuchs_hypgeom(oracle, gpu_driver):
    cases.case_fuchs_hypgeom(oracle, gpu_driver, 32, seed=1)


@pytest.mark.parametrize("L", [4, 32, 128])
def test_frobenius(oracle, gpu_driver, L):
    cases.case_frobenius(oracle, gpu_driver, L, seed=9, n=16 if L == 128 else 24, batch=70)


def test_frobenius_singleton(oracle, gpu_driver):
    cases.case_frobenius_singleton(oracle, gpu_driver, 32, seed=21)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_horner(oracle, gpu_driver, L):
    cases.case_horner(oracle, gpu_driver, L, seed=13, length=60 if L < 128 else 24, npts=333)


@pytest.mark.parametrize("L", [32, 64, 128])
def test_horner_cancellation(oracle, gpu_driver, L):
    cases.case_horner_cancellation(oracle, gpu_driver, L, seed=113 + L, length=24, npts=300)


def test_frobenius_2048(oracle, gpu_driver):
    cases.case_frobenius(oracle, gpu_driver, 64, seed=19, n=20, batch=300)


@pytest.mark.parametrize("hybrid", ["0", "1"])
@pytest.mark.parametrize("L", [64, 128])
def test_hybrid_scratch_columns(monkeypatch, oracle, gpu_driver, L, hybrid):
    """2048 / 4096 bits: the Horner and Frobenius (M = 1) kernels exist with all-shared scratch columns and with
    hybrid ones (L + 5 words in shared memory, the tail of the exact paths in global memory); the launcher picks by
    batch size, so both are forced here -- including the cancellation case that needs the full-width path and the
    long divisions of the Frobenius recurrence."""
    monkeypatch.setenv("CB200_HYBRID", hybrid)
    cases.case_horner(oracle, gpu_driver, L, seed=23, length=30, npts=300)
    cases.case_horner_cancellation(oracle, gpu_driver, L, seed=213 + L, length=20, npts=300)
    cases.case_frobenius(oracle, gpu_driver, L, seed=29, n=14, batch=300)


@pytest.mark.parametrize("tma", ["0", "1"])
@pytest.mark.parametrize("L", [64, 128])
def test_horner_series_staged_by_tma(monkeypatch, oracle, gpu_driver, L, tma):
    """The 2048 / 4096-bit Horner kernel with the series streamed into shared memory by cp.async.bulk (tiles of 16 / 8
    coefficients, double buffered, mbarrier completion) against the kernel that loads every coefficient itself: series
    lengths that are and are not a whole number of tiles, one tile, ragged last blocks, several derivatives."""
    monkeypatch.setenv("CB200_TMA", tma)
    monkeypatch.setenv("CB200_HYBRID", "1")
    monkeypatch.setenv("CB200_COOP", "0")
    for length in (5, 16, 33, 70):
        cases.case_horner(oracle, gpu_driver, L, seed=41 + length, length=length, npts=300 if length < 70 else 520)
    cases.case_horner_cancellation(oracle, gpu_driver, L, seed=413 + L, length=20, npts=300)


@pytest.mark.parametrize("L", [64, 128])
def test_two_lanes_per_instance(monkeypatch, oracle, gpu_driver, L):
    """The Frobenius (M = 1) recurrence with TWO lanes per instance (each lane plays two of the four roles of
    cb_coop.cuh; CTAs of 120 / 104 instances with the [word][role][instance] column layout): ragged last CTAs, one
    instance, several work-unit ranges."""
    monkeypatch.setenv("CB200_COOP", "2")
    cases.case_horner(oracle, gpu_driver, L, seed=47, length=25, npts=301)  # (Horner on two lanes per point)
    cases.case_horner_cancellation(oracle, gpu_driver, L, seed=317 + L, length=20, npts=250)
    cases.case_frobenius(oracle, gpu_driver, L, seed=31, n=19 if L == 64 else 12, batch=333)
    cases.case_frobenius(oracle, gpu_driver, L, seed=32, n=9, batch=1)
    cases.case_frobenius(oracle, gpu_driver, L, seed=33, n=17, batch=121)


@pytest.mark.parametrize("coop", ["0", "1"])
@pytest.mark.parametrize("L", [64, 128])
def test_cooperating_lanes(monkeypatch, oracle, gpu_driver, L, coop):
    """2048 / 4096 bits: the Frobenius (M = 1) and Horner kernels with FOUR lanes per instance (cb_coop.cuh; the
    launcher picks them for batches below two waves of the one-thread kernels) forced on and off: ragged batches
    (a last CTA with idle quads, a last warp with idle lanes), several work-unit ranges, the cancellation case that
    needs the exact full-width path on one lane, and the continuation pipeline on top."""
    monkeypatch.setenv("CB200_COOP", coop)
    cases.case_frobenius(oracle, gpu_driver, L, seed=31, n=19 if L == 64 else 12, batch=333)
    cases.case_horner(oracle, gpu_driver, L, seed=37, length=30, npts=301)
    cases.case_horner_cancellation(oracle, gpu_driver, L, seed=313 + L, length=20, npts=300)
    cases.case_continuation(oracle, gpu_driver, L, seed=123 + L, steps=3, n=24, batch=7)
    # the shared-operator Fuchs recurrence on four lanes per instance (what a continuation of few solutions runs)
    cases.case_fuchs_bessel_shared(oracle, gpu_driver, L, seed=5, n=70 if L == 64 else 40, batch=2)
    cases.case_fuchs_bessel_shared(oracle, gpu_driver, L, seed=6, n=40, batch=131)
    cases.case_fuchs_random_shared(oracle, gpu_driver, L, seed=141 + L, trials=3, batch=45)
    cases.case_fuchs_shared_adversarial(oracle, gpu_driver, L, seed=187 + L, batch=70)


@pytest.mark.parametrize("L", [32, 64, 128])
def test_continuation(oracle, gpu_driver, L):
    cases.case_continuation(oracle, gpu_driver, L, seed=23 + L, steps=6 if L < 128 else 3, n=40 if L < 128 else 24, batch=5)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_continuation_series_and_singular_corners(oracle, gpu_driver, L):
    cases.case_continuation_series(oracle, gpu_driver, L, seed=5 + L, n=24 if L < 64 else 16)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_radius_and_truncation(oracle, gpu_driver, L):
    cases.case_radius_and_truncation(oracle, gpu_driver, L, seed=11 + L, batch=6 if L < 128 else 3)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_ode_shift(oracle, gpu_driver, L):
    cases.case_ode_shift(oracle, gpu_driver, L, seed=7 + L, batch=37)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_taylor_shift(oracle, gpu_driver, L):
    cases.case_taylor_shift(oracle, gpu_driver, L, seed=9 + L, length=300 if L == 32 else 40, batch=3)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_exponent_overflow(oracle, gpu_driver, L):
    cases.case_exponent_overflow(oracle, gpu_driver, L)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_frobenius_general(oracle, gpu_driver, L):
    cases.case_frobenius_general(oracle, gpu_driver, L, seed=41 + L, trials=4 if L < 64 else (2 if L < 128 else 1))


@pytest.mark.parametrize("L", [4, 32, 128])
def test_frobenius_general_sweep(oracle, gpu_driver, L):
    cases.case_frobenius_general_sweep(oracle, gpu_driver, L, seed=43 + L, n=10 if L < 128 else 5, batch=40)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_solution_ops(oracle, gpu_driver, L):
    cases.case_solution_ops(oracle, gpu_driver, L, seed=47 + L, batch=35)


@pytest.mark.parametrize("L", [4, 32])
def test_indicial_polynomial(oracle, gpu_driver, L):
    cases.case_indicial_polynomial(oracle, gpu_driver, L, seed=53 + L, trials=3)


@pytest.mark.parametrize("L", [4, 32, 128])
def test_frobenius_rhs(oracle, gpu_driver, L):
    cases.case_frobenius_rhs(oracle, gpu_driver, L, seed=59 + L, n=14 if L < 128 else 8, batch=33)


def test_reduced_radix_products(oracle, gpu_driver, monkeypatch):
    """cb_rr.cuh (opt-in, CB200_RR=1): same balls as the oracle on the GPU."""
    monkeypatch.setenv("CB200_RR", "1")
    cases.case_fuchs_bessel_shared(oracle, gpu_driver, 32, seed=5, n=60, batch=300)
    cases.case_fuchs_random_shared(oracle, gpu_driver, 32, seed=63, trials=6, batch=45)
    cases.case_fuchs_shared_adversarial(oracle, gpu_driver, 32, seed=109, batch=70)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_exp_log(oracle, gpu_driver, L):
    cases.case_exp_log(oracle, gpu_driver, L, seed=61 + L, n=40 if L < 128 else 24)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_solution_evaluate(oracle, gpu_driver, L):
    cases.case_solution_evaluate(oracle, gpu_driver, L, seed=67 + L, npts=40 if L < 128 else 12, cap=16 if L < 128 else 10)


@pytest.mark.parametrize("L", [4, 32, 64, 128])
def test_ode_apply(oracle, gpu_driver, L):
    cases.case_ode_apply(oracle, gpu_driver, L, seed=71 + L, batch=40)


@pytest.mark.parametrize("L", [4, 32, 64])
def test_frobenius_general_edge(oracle, gpu_driver, L):
    cases.case_frobenius_general_edge(oracle, gpu_driver, L)


@pytest.mark.parametrize("batch", [385, 1000])
def test_persistent_recurrence_kernel(oracle, gpu_driver, monkeypatch, batch):
    """The persistent form of the recurrence kernel (work units = instance block x coefficient range, one CTA per
    SM; chosen by itself for batches beyond one wave of 384-thread blocks) forced onto small ragged batches:
    several ranges, idle lanes in the last block, bit-identical to the oracle."""
    monkeypatch.setenv("CB200_PERSISTENT", "1")
    cases.case_fuchs_bessel_shared(oracle, gpu_driver, 32, seed=5, n=75, batch=batch)
    cases.case_fuchs_shared_adversarial(oracle, gpu_driver, 32, seed=109, batch=70)


@pytest.mark.parametrize("L", [32, 64])
def test_persistent_integer_operator_kernel(oracle, gpu_driver, monkeypatch, L):
    """The persistent form of the integer-operator (Legendre sweep) kernel, forced onto a small batch."""
    monkeypatch.setenv("CB200_PERSISTENT", "1")
    cases.case_fuchs_integer_operators(oracle, gpu_driver, L, seed=15 + L)
