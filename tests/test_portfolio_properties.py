"""CPU: property tests (hypothesis) of the oracle's Portfolio restatement, and — when the reference is
mounted — equality with the reference's own `Portfolio` class on random states (utils/portfolio.py)."""
import math
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as orc

price_s = st.floats(min_value=0.5, max_value=5e4, allow_nan=False)
target_s = st.sampled_from([-3, -2, -1, -0.5, 0, 0.5, 1, 1.5, 2, 3])
fee_s = st.sampled_from([0.0, 1e-4, 1e-3, 0.01])


def _state(position, value, price, rate):
    """A reachable state: fresh TargetPortfolio + one interest accrual."""
    return orc.update_interest(orc.target_portfolio(position, value, price), rate)


@settings(max_examples=300, deadline=None)
@given(p0=target_s, tgt=target_s, price=price_s, move=st.floats(0.8, 1.25), fee=fee_s)
def test_trade_reaches_the_target_position(p0, tgt, price, move, fee):
    s = _state(p0, 1000.0, price, 3e-6)
    p = price * move
    if orc.valorisation(s, p) <= 50.0:
        return
    s2 = orc.trade_to_position(s, tgt, p, fee)
    assert abs(orc.position_of(s2, p) - tgt) <= 1e-11          # ~4e-15 in the common case


@settings(max_examples=300, deadline=None)
@given(p0=target_s, tgt=target_s, price=price_s, move=st.floats(0.8, 1.25))
def test_trade_conserves_valuation_without_fees(p0, tgt, price, move):
    s = _state(p0, 1000.0, price, 0.0)
    p = price * move
    v = orc.valorisation(s, p)
    if v <= 50.0:
        return
    s2 = orc.trade_to_position(s, tgt, p, 0.0)
    assert math.isclose(orc.valorisation(s2, p), v, rel_tol=1e-12)


@settings(max_examples=200, deadline=None)
@given(p0=target_s, tgt=target_s, price=price_s, fee=st.sampled_from([1e-4, 1e-3, 0.01]))
def test_fees_never_increase_valuation(p0, tgt, price, fee):
    s = _state(p0, 1000.0, price, 3e-6)
    v = orc.valorisation(s, price)
    s2 = orc.trade_to_position(s, tgt, price, fee)
    assert orc.valorisation(s2, price) <= v * (1 + 1e-12)


def test_update_interest_overwrites_and_keeps_positive_zero():
    s = orc.update_interest([2.0, -500.0, 7.0, 9.0], 1e-3)
    assert s.tolist() == [2.0, -500.0, 0.0, 0.5]                    # overwrite, not accumulate (portfolio.py:44-46)
    s = orc.update_interest([-3.0, 10.0, 0.0, 0.0], 1e-3)
    assert s[2] == 0.003 and s[3] == 0.0 and not np.signbit(s[3])  # max(0, -x) returns +0, never -0
    s = orc.update_interest([0.0, 0.0, 1.0, 1.0], 1e-3)
    assert not np.signbit(s[2]) and not np.signbit(s[3])


def test_all_five_trade_branches_are_exercised():
    """reduce-short repay, reduce-leverage repay, no repay, buy, sell."""
    hit = set()
    rng = np.random.default_rng(0)
    for _ in range(4000):
        p0, tgt = rng.choice([-3, -1, 0, 0.5, 1, 2, 3], 2)
        price = float(rng.uniform(50, 150))
        s = _state(p0, 1000.0, price, 1e-4)
        p = price * float(rng.uniform(0.9, 1.1))
        v = orc.valorisation(s, p)
        if v <= 10:
            continue
        cur = orc.position_of(s, p)
        if tgt <= 0 and cur < 0 and tgt / cur < 1:
            hit.add("repay_short")
        elif tgt >= 1 and cur > 1 and (tgt - 1) / (cur - 1) < 1:
            hit.add("repay_leverage")
        else:
            hit.add("no_repay")
        s2 = orc.trade_to_position(s, tgt, p, 1e-4)
        hit.add("buy" if s2[0] > s[0] else "sell")
    assert hit == {"repay_short", "repay_leverage", "no_repay", "buy", "sell"}


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gym_trading_env"), reason="reference not mounted")
def test_oracle_portfolio_is_bit_identical_to_the_reference_class():
    sys.path.insert(0, "/root/reference/src/gym_trading_env/utils")
    import portfolio as refp                                     # utils/portfolio.py has no imports
    rng = np.random.default_rng(1)
    n_checked = 0
    for _ in range(3000):
        p0, tgt = rng.choice([-3, -2, -1, -0.5, 0, 0.5, 1, 1.5, 2, 3], 2)
        price = np.float64(rng.uniform(10, 1000))
        pf = refp.TargetPortfolio(position=float(p0), value=1000.0, price=price)
        pf.update_interest(3e-6)
        s = _state(float(p0), 1000.0, float(price), 3e-6)
        p2 = np.float64(price * rng.uniform(0.9, 1.1))
        if pf.valorisation(p2) <= 10:
            continue
        pf.trade_to_position(float(tgt), p2, 1e-4)
        s = orc.trade_to_position(s, float(tgt), float(p2), 1e-4)
        pf.update_interest(3e-6)
        s = orc.update_interest(s, 3e-6)
        ref = np.array([pf.asset, pf.fiat, pf.interest_asset, pf.interest_fiat], dtype=np.float64)
        assert ref.tobytes() == s.tobytes()
        assert np.float64(pf.valorisation(p2)).tobytes() == np.float64(orc.valorisation(s, float(p2))).tobytes()
        n_checked += 1
    assert n_checked > 2000


def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10; counter = (tick_lo, tick_hi, env_lo, env_hi), key = seed."""
    def ph(c, k):
        return [int(x) for x in orc.philox(k[0] | (k[1] << 32), c[0] | (c[1] << 32), c[2] | (c[3] << 32))]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
