"""The reference's own unit/property tests, re-run against the oracle.

These are the only tests the reference holds for the hot path; they pin
invariants, not numbers.  Each test names the reference test it restates.
"""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle_binding as ob

SEGS = [ob.SEG_DETERMINISTIC, ob.SEG_BINOMIAL_NO_UNEVEN, ob.SEG_BINOMIAL_NO_NMINUS, ob.SEG_BINOMIAL]


@st.composite
def distributions(draw):
    """lib.rs:61-75 NonEmptyDistribtionWithNPlusCells: up to 500 draws of an even copy number in
    [2,254] with 1..255 cells each (later draws overwrite), plus 1..255 cells without ecDNA."""
    n = draw(st.integers(1, 40))
    h = np.zeros(512, dtype=np.uint64)
    for _ in range(n):
        k = draw(st.integers(1, 255))
        if k == 1 or k % 2 == 1:  # lib.rs:83-86
            k = 2
        h[k] = draw(st.integers(1, 255))
    h[0] = draw(st.integers(1, 255))
    return h


def _apply(h, event, seg, seed):
    h = h.copy()
    k, k1, k2, un = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = ob.lib().orc_apply_event(h.ctypes.data, len(h), event, seg, seed, C.byref(k), C.byref(k1), C.byref(k2),
                                  C.byref(un))
    return rc, h, k.value, k1.value, k2.value, un.value


@settings(max_examples=150, deadline=None)
@given(seed=st.integers(0, 2 ** 64 - 1), h=distributions(), seg=st.sampled_from(SEGS))
def test_increase_nplus(seed, h, seg):
    """proliferation.rs:159-242 increase_nplus_test."""
    nplus, nminus = int(h[1:].sum()), int(h[0])
    rc, h2, k, k1, k2, un = _apply(h, ob.EV_BIRTH_NPLUS, seg, seed)
    assert rc == 0 and k1 + k2 == 2 * k
    nplus2, nminus2 = int(h2[1:].sum()), int(h2[0])
    if seg == ob.SEG_DETERMINISTIC:
        assert un == 0 and nplus2 == nplus + 1 and nminus2 == nminus
    if seg == ob.SEG_BINOMIAL_NO_UNEVEN:
        assert un == 0
    if un == 0:  # IsUneven::False
        assert (nplus2, nminus2) == (nplus + 1, nminus)
    elif un == 1:  # IsUneven::True
        assert (nplus2, nminus2) == (nplus, nminus + 1)
    else:  # TrueWithoutNMinusIncrease
        assert (nplus2, nminus2) == (nplus, nminus)
    # copies are conserved up to the doubling: total' = total + k
    tot = lambda x: int((np.arange(len(x), dtype=np.uint64) * x).sum())
    assert tot(h2) == tot(h) + k


@settings(max_examples=50, deadline=None)
@given(h=distributions())
def test_increase_nminus(h):
    """proliferation.rs:244-256."""
    rc, h2, *_ = _apply(h, ob.EV_BIRTH_NMINUS, ob.SEG_BINOMIAL, 0)
    assert rc == 0 and h2[0] == h[0] + 1 and np.array_equal(h2[1:], h[1:])


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 2 ** 64 - 1), h=distributions())
def test_decrease_nplus(seed, h):
    """proliferation.rs:258-272."""
    rc, h2, k, *_ = _apply(h, ob.EV_DEATH_NPLUS, ob.SEG_BINOMIAL, seed)
    assert rc == 0 and h2[0] == h[0] and h2[1:].sum() == h[1:].sum() - 1 and h2[k] == h[k] - 1


@settings(max_examples=50, deadline=None)
@given(h=distributions())
def test_decrease_nminus(h):
    """proliferation.rs:274-286."""
    rc, h2, *_ = _apply(h, ob.EV_DEATH_NMINUS, ob.SEG_BINOMIAL, 0)
    assert rc == 0 and h2[0] == h[0] - 1 and np.array_equal(h2[1:], h[1:])


def test_division_without_nplus_cells_is_an_error():
    """proliferation.rs:55-57: pick_remove_random_nplus fails on an ecDNA-free population."""
    h = np.zeros(16, dtype=np.uint64)
    h[0] = 5
    assert _apply(h, ob.EV_BIRTH_NPLUS, ob.SEG_BINOMIAL, 1)[0] == -1


def _seg(rule, copies, seed):
    k1, k2, un = C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = ob.lib().orc_segregate(rule, copies, seed, C.byref(k1), C.byref(k2), C.byref(un))
    return rc, k1.value, k2.value, un.value


@pytest.mark.parametrize("copies", [0, 1, 3, 255])
def test_dna_copy_segregating_rejects(copies):
    """segregation.rs:223-239: zero, one and odd values are not DNACopySegregating."""
    assert _seg(ob.SEG_BINOMIAL, copies, 0)[0] == -1


@settings(max_examples=100, deadline=None)
@given(half=st.integers(1, 32767), seed=st.integers(0, 2 ** 64 - 1))
def test_segregation_rules(half, seed):
    """segregation.rs:248-291: Deterministic halves; Binomial conserves copies and flags uneven iff
    a daughter got nothing; NoUneven never returns an empty daughter."""
    copies = 2 * half
    rc, k1, k2, un = _seg(ob.SEG_DETERMINISTIC, copies, seed)
    assert rc == 0 and k1 == k2 == half and un == 0
    rc, k1, k2, un = _seg(ob.SEG_BINOMIAL, copies, seed)
    assert k1 + k2 == copies and (un == 1) == (k1 == 0 or k2 == 0)
    rc, k1, k2, un = _seg(ob.SEG_BINOMIAL_NO_UNEVEN, copies, seed)
    assert k1 + k2 == copies and un == 0 and k1 > 0 and k2 > 0
    rc, k1, k2, un = _seg(ob.SEG_BINOMIAL_NO_NMINUS, copies, seed)
    assert k1 + k2 == copies and (un == 2) == (k1 == 0 or k2 == 0) and un != 1


@settings(max_examples=30, deadline=None)
@given(h=distributions(), t=st.integers(0, 255))
def test_process_construction_preserves_state(h, t):
    """process.rs:356-384 create_birth_death_process_test: a run that stops at once (time cap
    already reached) returns the initial distribution, counts and clock untouched."""
    init = {int(k): int(c) for k, c in enumerate(h) if c}
    o = ob.make_opts(b1=1.2, d0=0.1, d1=0.1, max_cells=10 ** 6, max_time=0.0, initial=init, state=ob.STATE_VECTOR,
                     rng=ob.RNG_RAND)
    r = ob.run(o, hist_cap=512)
    assert r.stop_reason == ob.STOP_MAX_TIME and r.n_events == 0 and r.time == 0.0
    assert np.array_equal(r.hist, h) and r.nminus == h[0] and r.nplus == h[1:].sum()
    assert ob.stats(r.hist)[0] == ob.stats(h)[0]
