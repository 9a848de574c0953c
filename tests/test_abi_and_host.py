"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the headers
declare (no compute calls without a GPU), refuses to run without a device, and the host logic
(valuation, layout packing, partitioning) behaves as the reference's."""
import os
import re

import numpy as np
import pytest

from cascade_b200 import _lib, api, sharding
from cascade_b200.format import (FLAG_ZERO, Balls, entry_major_to_instance_major, from_complex,
                                 instance_major_to_entry_major, pack_entries, random_balls, unpack_entries)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_\w+|acb_ode_\w+|_acb_ode_\w+|analytic_continuation|indicial_polynomial(?:_evaluate)?|find_monodromy_matrix|radius_of_convergence|truncation_order|flint_rand\w+|n_randint)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    lib = _lib.lib()
    names = declared("cascade_b200.h")
    assert set(_lib.SYMBOLS) <= set(names)
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/cascade_b200.h but not exported"
    assert b"sm_100a" in lib.cb200_version()
    # the reference's own names live in the SEPARATE compat library (for machines without Arb)
    compat = ctypes.CDLL(os.path.join(ROOT, "cascade_b200", "libcascade_b200_compat.so"))
    for name in [n for n in declared("cascade_compat.h") if not n.endswith("_struct") and not n.startswith("cb200_")]:
        if name in ("acb_ode_poly", "acb_ode_coeff", "acb_ode_t", "acb_ode_solution_t"):
            continue  # macros / types
        assert hasattr(compat, name), f"{name} is declared in include/cascade_compat.h but not exported"


def test_batch_library_exports_only_its_own_prefix_and_links_next_to_arb_names():
    """libcascade_b200.so must be linkable into a Cascade that is built against the REAL Arb (INTEGRATION.md
    section 2): it exports cb200_* and nothing else, so a program that defines (or links a library defining)
    acb_mul, acb_init, acb_poly_* ... -- Arb's names -- does not clash with it."""
    import subprocess
    so = os.path.join(ROOT, "cascade_b200", "libcascade_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], check=True, capture_output=True, text=True).stdout
    syms = [line.split()[-1] for line in out.splitlines() if line.strip()]
    assert syms and all(x.startswith("cb200_") for x in syms), [x for x in syms if not x.startswith("cb200_")][:5]
    src = os.path.join(ROOT, "tests", "c", "link_next_to_arb.c")
    exe = os.path.join(ROOT, "tests", "c", "link_next_to_arb")
    subprocess.run(["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                    "-L", os.path.join(ROOT, "cascade_b200"), "-lcascade_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "cascade_b200")], check=True)
    r = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "own acb_mul: 42" in r and "cascade_b200" in r


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly, never compute on the CPU."""
    lib = _lib.lib()
    if lib.cb200_device_count() > 0:
        pytest.skip("a GPU is present")
    x = from_complex([1, 2], 32)
    with pytest.raises(_lib.CascadeB200Error):
        api.ew(api.OP_MUL, x, x)


def test_valuation_matches_reference_rule():
    """reference src/acb_ode.c:159-181 on the three shipped examples (src/examples.c)."""
    L = 32
    assert api.acb_ode_valuation(api.acb_ode_legendre_batch([4], L), 2, 2) == -2
    # hypgeom: z(1-z) y'' + ... : valuation -1;  Bessel at the origin: valuation 0
    hyp = Balls.zeros(18, L)
    one = from_complex([1], L)
    for e in (2 * 3 + 1, 2 * 3 + 2, 1 * 3 + 0, 1 * 3 + 1, 0):
        hyp.mant[2 * e:2 * e + 2] = one.mant
        hyp.meta[2 * e:2 * e + 2] = one.meta
    assert api.acb_ode_valuation(hyp, 2, 2) == -1
    bes = Balls.zeros(18, L)
    for e in (2 * 3 + 2, 1 * 3 + 1, 0 * 3 + 2, 0):
        bes.mant[2 * e:2 * e + 2] = one.mant
        bes.meta[2 * e:2 * e + 2] = one.meta
    assert api.acb_ode_valuation(bes, 2, 2) == 0
    with pytest.raises(_lib.CascadeB200Error):
        api.acb_ode_solve_fuchs_batch(bes, Balls.zeros(0, L), 2, 2, 5, 1, True)


def test_mixed_valuations_are_rejected():
    """A batched launch runs with one valuation: per-instance operators whose valuations differ (exact-zero
    leading coefficients, as acb_ode_random can produce) must be rejected, not solved with instance 0's window."""
    L = 32
    good = api.acb_ode_legendre_batch([3, 4, 5], L)            # valuation -2 each
    assert api.acb_ode_valuations(good, 2, 2).tolist() == [-2, -2, -2]
    assert api.batch_valuation(good, 2, 2, False, "t") == -2
    bad = good.copy()
    e = 2 * (1 * 9 + 2 * 3 + 0)                                 # instance 1: zero P[2][0] -> (1 - z^2) becomes -z^2: valuation 0
    bad.mant[e:e + 2] = 0
    bad.meta[e:e + 2] = 0
    bad.meta[e:e + 2, 1] = FLAG_ZERO
    assert api.acb_ode_valuations(bad, 2, 2).tolist() == [-2, 0, -2]
    for inst in range(3):
        assert api.acb_ode_valuations(bad, 2, 2)[inst] == api.acb_ode_valuation(bad, 2, 2, inst)
    with pytest.raises(_lib.CascadeB200Error, match="different valuations"):
        api.acb_ode_solve_fuchs_batch(bad, Balls.zeros(3 * 2 * 2, L), 2, 2, 5, 3, False)
    with pytest.raises(_lib.CascadeB200Error, match="different valuations"):
        api.acb_ode_solve_frobenius1_batch(bad, Balls.zeros(3 * 2, L), 2, 2, 5, 3, False)


def test_layout_roundtrip():
    rng = np.random.default_rng(0)
    b = random_balls(rng, 5 * 7 * 2, 4, -3, 3, "ulp")
    e = instance_major_to_entry_major(b, 5, 7)
    m, t = pack_entries(e, 7, 10)
    assert m.shape == (7, 4, 10) and t.shape == (7, 10, 4)
    # limb-major: limb k of slot s of entry e
    assert m[3, 2, 5] == e.mant[3 * 10 + 5, 2]
    back = entry_major_to_instance_major(unpack_entries(4, m, t), 5, 7)
    assert back.same_as(b)


def test_partition_covers_batch():
    for batch in (1, 7, 64, 65536, 1000003):
        for world in (1, 2, 4, 8):
            spans = [sharding.partition(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
