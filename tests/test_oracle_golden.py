"""CPU: the C restatement (oracle/gte_oracle.c) against the golden vectors recorded from the
UNMODIFIED reference (tests/golden/*.npz, written by oracle/make_golden.py)."""
import numpy as np
import pytest

import helpers as H


@pytest.mark.parametrize("name", H.GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name):
    g = H.load_golden(name)
    stats = H.replay_golden(H.OracleAdapter(H.make_oracle(g)), g, exact_money=True, check_final_obs=True)
    assert stats["steps"] == g["actions"].size
    # np.log (numpy SIMD) vs libm log: <= 1 ulp, only a few percent of steps
    assert stats["reward_bit_mismatch"] <= 0.25 * stats["steps"]


def test_normalised_oracle_differs_from_raw_reference_only_in_stale_rows():
    """Hazard H3: the raw reference leaks dynamic-feature rows of earlier episodes into later
    windows.  The zero-on-reset model (what the device implements) must differ from the raw golden
    ONLY in dynamic columns of rows before the current episode's start, and only after an env
    starts re-visiting rows."""
    g = H.load_golden("raw_stale_dynamic_rows")
    env = H.make_oracle(g, dyn_mode=1)
    env.reset()
    ns = g["features"].shape[2]
    W = g["params"]["windows"]
    n_diff = 0
    for k in range(g["actions"].shape[0]):
        env.step(g["actions"][k])
        assert H.bits_equal(env.obs[..., :ns], g["obs"][k][..., :ns])          # static part always identical
        diff = env.obs[..., ns:] != g["obs"][k][..., ns:]
        if diff.any():
            n_diff += 1
            for i, w, _ in np.argwhere(diff):
                row = int(env.idx[i]) - W + 1 + int(w)
                assert row < int(env.ep_start[i])                                # only pre-episode rows
                assert env.obs[i, w, ns:].tolist() == [0.0, 0.0]
        assert H.bits_equal(env.valuation, g["valuation"][k])                   # money never affected
    assert n_diff > 0   # the fixture really exercises the leak


def test_golden_covers_the_edge_cases():
    names = set(H.GOLDEN_NAMES)
    assert {"c1_single_nowindow", "c3_windows_leveraged", "termination_stop", "multi_dataset_k1",
            "multi_dataset_k3", "btc_luckymodel_config", "no_dynamic_features"} <= names
    assert H.load_golden("termination_stop")["terminated"].sum() > 10
    assert (H.load_golden("w64_fixed_start_holds")["actions"] < 0).any()
    g = H.load_golden("c1_single_nowindow")
    assert g["truncated"].sum() == 2 and (g["idx"][g["truncated"].astype(bool)] == g["lengths"][0] - 1).all()
    g = H.load_golden("multi_dataset_k1")
    assert len(set(g["post_dataset"].ravel().tolist())) == 4 and len(set(g["lengths"].tolist())) == 4
