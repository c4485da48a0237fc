"""Known-answer and distribution tests that pin the oracle's building blocks.

The reference's own tests hold no numeric golden vector for this path, so the
anchors are the published vectors of the primitives (Random123 kat_vectors for Philox4x32-10, the
ChaCha reference keystreams for 8 and 20 rounds) and exact distributions (binomial pmf, Exp(1) cdf).
"""
import ctypes as C

import numpy as np
import pytest
from scipy import stats as sps

import oracle_binding as ob


def test_philox4x32_10_random123_vectors():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert list(ob.philox(ctr, key)) == want


def _chacha(key, counter, stream, rounds):
    k = np.array(key, dtype=np.uint32)
    out = np.zeros(16, dtype=np.uint32)
    ob.lib().orc_chacha_block(k.ctypes.data, counter, stream, rounds, out.ctypes.data)
    return out.tobytes().hex()


def test_chacha_zero_key_keystreams():
    # ChaCha20 and ChaCha8, all-zero key and nonce, block 0 (Bernstein's reference vectors)
    assert _chacha([0] * 8, 0, 0, 20).startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    assert _chacha([0] * 8, 0, 0, 8).startswith("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e")


def test_chacha8_stream_layout():
    """BlockRng semantics: next_u64 = two consecutive words, low first; streams are independent."""
    key = np.zeros(8, dtype=np.uint32)
    ob.lib().orc_seed_from_u64(26, key.ctypes.data)
    out = np.zeros(40, dtype=np.uint64)
    ob.lib().orc_chacha8_u64(26, 260, 40, out.ctypes.data)
    blk = np.zeros(16, dtype=np.uint32)
    ob.lib().orc_chacha_block(key.ctypes.data, 0, 260, 8, blk.ctypes.data)
    assert int(out[0]) == (int(blk[1]) << 32) | int(blk[0])
    assert int(out[7]) == (int(blk[15]) << 32) | int(blk[14])
    ob.lib().orc_chacha_block(key.ctypes.data, 4, 260, 8, blk.ctypes.data)  # 5th block = second refill
    assert int(out[32]) == (int(blk[1]) << 32) | int(blk[0])
    other = np.zeros(40, dtype=np.uint64)
    ob.lib().orc_chacha8_u64(26, 261, 40, other.ctypes.data)
    assert not np.any(out == other)


def test_deterministic_log_accuracy_and_edges():
    L = ob.lib()
    assert L.orc_neg_log_u24((1 << 24) - 1) == 0.0  # u = 1
    assert abs(L.orc_neg_log_u24(0) - 24 * np.log(2)) < 2e-6
    rng = np.random.default_rng(1)
    for m in rng.integers(0, (1 << 24) - 1, 20000):
        got = L.orc_neg_log_u24(int(m))
        want = -np.log((int(m) + 1) / 2.0 ** 24)
        assert got > 0 and abs(got - want) <= 2.5e-7 * max(want, 1e-3) + 1e-9, (m, got, want)


@pytest.mark.parametrize("k", [1, 2, 5, 9, 10, 11, 31, 32, 33, 128, 129, 500, 1024, 1025, 5000, 32767])
def test_popcount_binomial_matches_exact_pmf(k):
    """Philox segregation draw: chi-square against Binomial(2k, 1/2), incl. the slot boundaries
    (64 bits per slot)."""
    n = 2 * k
    N = 20000
    draws = np.array([ob.lib().orc_binomial_half_philox(7, 3, e, 0, n) for e in range(N)])
    assert draws.min() >= 0 and draws.max() <= n
    lo, hi = int(sps.binom.ppf(1e-4, n, 0.5)), int(sps.binom.ppf(1 - 1e-4, n, 0.5))
    edges = np.unique(np.linspace(lo, hi + 1, min(hi - lo + 2, 30)).astype(int))
    obs = np.histogram(draws, bins=np.concatenate([[-1], edges, [n + 1]]))[0]
    cdf = sps.binom.cdf(np.concatenate([edges - 1, [n]]), n, 0.5)
    exp = N * np.diff(np.concatenate([[0], cdf]))
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert sps.chi2.sf(chi2, keep.sum() - 1) > 1e-4, (k, chi2)
    assert abs(draws.mean() - k) < 5 * np.sqrt(k / 2 / N)


@pytest.mark.parametrize("k", [1, 4, 9, 10, 50, 500, 5000, 32767])
def test_rand_binomial_matches_exact_pmf(k):
    """The restated rand_distr Binomial (BINV for 2k*0.5 < 10, BTPE otherwise)."""
    n, N = 2 * k, 40000
    out = np.zeros(N, dtype=np.uint64)
    ob.lib().orc_rand_binomial(11, k, n, 0.5, N, out.ctypes.data)
    assert out.max() <= n
    lo, hi = int(sps.binom.ppf(1e-4, n, 0.5)), int(sps.binom.ppf(1 - 1e-4, n, 0.5))
    edges = np.unique(np.linspace(lo, hi + 1, min(hi - lo + 2, 30)).astype(int))
    obs = np.histogram(out.astype(np.int64), bins=np.concatenate([[-1], edges, [n + 1]]))[0]
    cdf = sps.binom.cdf(np.concatenate([edges - 1, [n]]), n, 0.5)
    exp = N * np.diff(np.concatenate([[0], cdf]))
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert sps.chi2.sf(chi2, keep.sum() - 1) > 1e-4, (k, chi2)


def test_rand_exp1_is_exponential():
    N = 200000
    out = np.zeros(N, dtype=np.float32)
    ob.lib().orc_rand_exp1_f32(5, 9, N, out.ctypes.data)
    assert out.min() >= 0
    assert sps.kstest(out.astype(np.float64), "expon").pvalue > 1e-3
    assert out.max() > 9.0  # the ziggurat tail branch is exercised


def test_philox_waiting_time_is_exponential():
    e = np.array([ob.lib().orc_neg_log_u24(int(ob.philox([i, 0, 1, 0], [3, 0])[0]) >> 8) for i in range(100000)])
    assert sps.kstest(e.astype(np.float64), "expon").pvalue > 1e-3


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 12345678, (1 << 32) - 1])
def test_pick_is_uniform(n):
    N = 30000
    a = np.zeros(N, dtype=np.uint64)
    ob.lib().orc_rand_gen_range(1, 2, n, N, a.ctypes.data)
    b = np.array([ob.lib().orc_pick_philox(1, 2, e, n) for e in range(N)], dtype=np.uint64)
    for x in (a, b):
        assert x.max() < n
        if 1 < n <= 7:
            obs = np.bincount(x.astype(np.int64), minlength=n)
            assert sps.chisquare(obs).pvalue > 1e-4
        elif n > 1000:
            assert sps.kstest(x.astype(np.float64) / n, "uniform").pvalue > 1e-3
