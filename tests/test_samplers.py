"""Pins the RNG contract (hibayes_b200/csrc/hb_rng.h) that replaces libR's stream:
Philox4x32-10 against the Random123 known-answer vectors, the AS241 normal quantile against
scipy, and each sampler of stats.cpp:3-28,55-76 against the distribution R documents for it."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats


def _philox(oracle, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    oracle.lib().hbo_philox(c, k, out)
    return [int(v) for v in out]


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _philox(oracle, [0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox(oracle, [0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox(oracle, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_qnorm_matches_scipy(oracle):
    L = oracle.lib()
    ps = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 20001), 10.0 ** -np.arange(3, 16), 1 - 10.0 ** -np.arange(3, 15)])
    got = np.array([L.hbo_qnorm(float(p)) for p in ps])
    ref = stats.norm.ppf(ps)
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-14)


def _draws(oracle, n, slot=1):
    L = oracle.lib()
    u, z = C.c_double(), C.c_double()
    us, zs = np.empty(n), np.empty(n)
    for i in range(n):
        L.hbo_draw_uz(12345, 1, 7, i, slot, 0, C.byref(u), C.byref(z))
        us[i], zs[i] = u.value, z.value
    return us, zs


def test_uniform_and_normal_distribution(oracle):
    us, zs = _draws(oracle, 40000)
    assert us.min() > 0.0 and us.max() < 1.0
    assert stats.kstest(us, "uniform").pvalue > 1e-3
    assert stats.kstest(zs, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(us, zs)[0, 1]) < 0.02


@pytest.mark.parametrize("shape", [0.3, 1.0, 2.5, 52.0, 5000.5])
def test_gamma_distribution(oracle, shape):
    L = oracle.lib()
    x = np.array([L.hbo_draw_gamma(99, 0, 3, i, 0, shape) for i in range(20000)])
    assert stats.kstest(x, "gamma", args=(shape,)).pvalue > 1e-3
    assert abs(x.mean() / shape - 1) < 0.05


@pytest.mark.parametrize("df", [5.0, 298.0])
def test_chisq_distribution(oracle, df):
    L = oracle.lib()
    x = np.array([L.hbo_draw_chisq(7, 1, 0, i, 0, df) for i in range(20000)])
    assert stats.kstest(x, "chi2", args=(df,)).pvalue > 1e-3


def test_inverse_gaussian_distribution(oracle):
    L = oracle.lib()
    us, zs = _draws(oracle, 20000, slot=2)
    mu, lam = 1.7, 3.2
    x = np.array([L.hbo_invgauss(mu, lam, u, z) for u, z in zip(us, zs)])
    # scipy: invgauss(mu/lam, scale=lam) is IG(mean=mu, shape=lam)
    assert stats.kstest(x, "invgauss", args=(mu / lam, 0, lam)).pvalue > 1e-3


def test_inverse_gaussian_stable_form_equals_reference_expression(oracle):
    """hb_rng.h evaluates the smaller root of stats.cpp:57-59 as mu/(1+w+sqrt(w(w+2))); where the
    reference's own expression is well conditioned (w = mu*z^2/(2*lambda) <~ 1) the two agree to
    rounding, and the stable form stays accurate where the literal one cancels."""
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(2000):
        mu, lam, z = rng.uniform(0.1, 3), rng.uniform(1, 50), rng.normal()
        lit = L.hbo_invgauss_literal_root(mu, lam, z)
        got = L.hbo_invgauss(mu, lam, 0.0, z)      # u = 0 always takes the root itself
        assert abs(got / lit - 1) < 1e-12
    # ill-conditioned corner reached through the |g| >= 1e-6 clamp of Bayes.cpp:728
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    mu, lam, z = 2.8e8, 800.0, 1.3
    w = Decimal(mu) * Decimal(z) * Decimal(z) / (2 * Decimal(lam))
    exact = Decimal(mu) / (1 + w + (w * (w + 2)).sqrt())
    assert abs(Decimal(L.hbo_invgauss(mu, lam, 0.0, z)) / exact - 1) < Decimal(1e-14)
    assert abs(Decimal(L.hbo_invgauss_literal_root(mu, lam, z)) / exact - 1) > Decimal(1e-9)


def test_draws_are_position_addressed(oracle):
    L = oracle.lib()
    a = L.hbo_draw_chisq(5, 1, 2, 33, 0, 5.0)
    [L.hbo_draw_chisq(5, 1, 2, k, 0, 5.0) for k in range(10)]
    assert a == L.hbo_draw_chisq(5, 1, 2, 33, 0, 5.0)
    assert a != L.hbo_draw_chisq(6, 1, 2, 33, 0, 5.0)
    assert a != L.hbo_draw_chisq(5, 1, 3, 33, 0, 5.0)
