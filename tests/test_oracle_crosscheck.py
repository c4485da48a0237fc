"""Independent cross-checks of the oracle's restated third-party algorithms at high
statistics: 10^7 draws per case, tails included, against exact distributions (scipy) and against an
independent implementation of the same published algorithm (numpy's Generator.binomial is BTPE/inversion as
well; numpy's standard_exponential is a 256-layer ziggurat), and the PCG32 reference implementation's own
demo vectors for the generator behind rand_core's seed_from_u64."""
import numpy as np
import pytest
from scipy import stats as sps

import oracle_binding as ob

N = 10_000_000
M64 = (1 << 64) - 1


def _pmf_check(draws, n, label):
    """Every outcome whose expected count is >= 10 individually (z-score), both tails pooled beyond that,
    a chi-square over all cells, and the first two moments."""
    obs = np.bincount(draws.astype(np.int64), minlength=n + 1).astype(np.float64)
    exp = len(draws) * sps.binom.pmf(np.arange(n + 1), n, 0.5)
    body = exp >= 10
    lo, hi = np.argmax(body), n - np.argmax(body[::-1])
    cells_o = np.concatenate([[obs[:lo].sum()], obs[lo:hi + 1], [obs[hi + 1:].sum()]])
    cells_e = np.concatenate([[exp[:lo].sum()], exp[lo:hi + 1], [exp[hi + 1:].sum()]])
    keep = cells_e >= 5
    z = (cells_o[keep] - cells_e[keep]) / np.sqrt(cells_e[keep])
    assert np.abs(z).max() < 5.5, (label, float(np.abs(z).max()))
    chi2 = float((z ** 2).sum())
    assert sps.chi2.sf(chi2, keep.sum() - 1) > 1e-5, (label, chi2, int(keep.sum()))
    # outcomes the exact pmf all but forbids must not appear
    assert obs[exp < 1e-6].sum() == 0, label
    assert abs(draws.mean() - n / 2) < 5.5 * np.sqrt(n / 4 / len(draws)), label
    assert abs(draws.var() - n / 4) < 6 * (n / 4) * np.sqrt(2.0 / len(draws)) + 1e-9, label


@pytest.mark.parametrize("k", [4, 9, 10, 11, 50, 1000, 32767])
def test_restated_rand_binomial_pmf_and_tails_at_1e7(k):
    """BINV (2k * 0.5 < 10) and BTPE with the switch at k = 10, incl. the top of u16 (segregation.rs:122-123)."""
    n = 2 * k
    out = np.zeros(N, dtype=np.uint64)
    ob.lib().orc_rand_binomial(1234 + k, k, n, 0.5, N, out.ctypes.data)
    _pmf_check(out, n, f"restated rand_distr binomial n={n}")
    # the independent implementation of the same algorithm family agrees cell by cell (two-sample)
    ref = np.random.default_rng(99 + k).binomial(n, 0.5, N)
    a = np.bincount(out.astype(np.int64), minlength=n + 1).astype(np.float64)
    b = np.bincount(ref, minlength=n + 1).astype(np.float64)
    keep = (a + b) >= 40
    z = (a[keep] - b[keep]) / np.sqrt(a[keep] + b[keep])
    assert np.abs(z).max() < 5.5 and sps.chi2.sf(float((z ** 2).sum()), keep.sum() - 1) > 1e-5


@pytest.mark.parametrize("k", [5, 64, 65, 2000])
def test_native_popcount_binomial_pmf_and_tails(k):
    """The native stream's segregation draw (popcount of 2k Philox bits, 128 per slot) at 2e6 draws."""
    n, cnt = 2 * k, 2_000_000 if k <= 65 else 300_000
    f = ob.lib().orc_binomial_half_philox
    draws = np.fromiter((f(5, 77, e, 0, n) for e in range(cnt)), dtype=np.int64, count=cnt)
    _pmf_check(draws, n, f"popcount binomial n={n}")


def test_restated_ziggurat_exp1_against_exact_cdf_at_1e7():
    out = np.zeros(N, dtype=np.float32)
    ob.lib().orc_rand_exp1_f32(3, 1, N, out.ctypes.data)
    x = out.astype(np.float64)
    # 200 equal-probability cells of Exp(1) plus the far tail in its own cells
    edges = np.concatenate([-np.log1p(-np.linspace(0, 1, 201)[:-1]), [8.0, 10.0, 12.0, 14.0, np.inf]])
    edges = np.unique(edges)
    obs = np.histogram(x, bins=edges)[0].astype(np.float64)
    exp = N * np.diff(-np.expm1(-edges))
    keep = exp >= 5
    z = (obs[keep] - exp[keep]) / np.sqrt(exp[keep])
    assert np.abs(z).max() < 5.5, float(np.abs(z).max())
    assert sps.chi2.sf(float((z ** 2).sum()), keep.sum() - 1) > 1e-5
    assert abs(x.mean() - 1.0) < 5.5 / np.sqrt(N) and abs(x.var() - 1.0) < 6 * np.sqrt(8.0 / N)
    # and against numpy's own 256-layer ziggurat (two-sample KS)
    ref = np.random.default_rng(17).standard_exponential(2_000_000)
    assert sps.ks_2samp(x[:2_000_000], ref).pvalue > 1e-3


def _pcg32(state, inc):
    """The PCG32 (XSH-RR 64/32) reference implementation, pcg_basic.c: output from the OLD state."""
    old = state
    state = (old * 6364136223846793005 + inc) & M64
    xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
    rot = old >> 59
    return state, ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF


def test_seed_from_u64_is_pcg32_with_the_reference_vectors():
    """rand_core 0.6.4 seed_from_u64 fills the seed with PCG32 outputs (fixed increment, output taken from the
    NEW state).  The permutation and the multiplier are pinned by pcg_basic's published demo output
    (pcg32_srandom(42, 54)); the oracle's key derivation is then that generator run from `seed`."""
    inc = (54 << 1) | 1
    state = 0
    state, _ = _pcg32(state, inc)
    state = (state + 42) & M64
    state, _ = _pcg32(state, inc)
    got = []
    for _ in range(6):
        state, v = _pcg32(state, inc)
        got.append(v)
    assert got == [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    for seed in (0, 1, 26, 2 ** 63, M64):
        s, want = seed, []
        for _ in range(8):
            s = (s * 6364136223846793005 + 11634580027462260723) & M64
            # output of the state just produced = what _pcg32 returns for it as its "old" state
            want.append(_pcg32(s, 0)[1])
        key = np.zeros(8, dtype=np.uint32)
        ob.lib().orc_seed_from_u64(seed, key.ctypes.data)
        assert [int(v) for v in key] == want, seed
