"""Property tests (hypothesis) of the CPU oracle itself -- the checker must obey the physics it encodes:
closed-form free flight, translation invariance, Newton's third law for contacts, landmark-permutation
invariance of the spread reward, the fp32 build staying inside the north-star tolerance of the fp64 build."""
import numpy as np
from hypothesis import given, settings, strategies as st

from helpers import descriptor, random_actions, random_states
from oracle import Oracle

SEEDS = st.integers(min_value=0, max_value=2**31 - 1)


@settings(max_examples=25, deadline=None)
@given(seed=SEEDS)
def test_free_flight_closed_form(seed):
    """no contact: v' = (1 - damping) v + u dt / m, p' = p + v' dt (core.py:158-169)"""
    rng = np.random.RandomState(seed)
    d = descriptor("simple_spread_n3")
    o = Oracle(d, "f64")
    n = 64
    pv = np.zeros((n, 3, 4))
    pv[:, :, 0:2] = rng.uniform(-1, 1, (n, 3, 2)) + np.array([[-10, 0], [0, 0], [10, 0]])[None]   # far apart
    pv[:, :, 2:4] = rng.uniform(-2, 2, (n, 3, 2))
    lm = rng.uniform(-1, 1, (n, 3, 2))
    act = random_actions(o.act_dims, n, rng)
    npv, _, obs, rew, done, _ = o.step(pv, lm, np.zeros((n, 3, 2)), act, 1)
    a = act.reshape(n, 3, 5)
    u = 5.0 * np.stack([a[:, :, 1] - a[:, :, 2], a[:, :, 3] - a[:, :, 4]], -1)
    v = pv[:, :, 2:4] * 0.75 + u * 0.1
    np.testing.assert_allclose(npv[:, :, 2:4], v, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(npv[:, :, 0:2], pv[:, :, 0:2] + v * 0.1, rtol=1e-12, atol=1e-12)
    assert not done.any()


@settings(max_examples=25, deadline=None)
@given(seed=SEEDS, tx=st.floats(-3, 3), ty=st.floats(-3, 3))
def test_translation_invariance_and_momentum(seed, tx, ty):
    rng = np.random.RandomState(seed)
    d = descriptor("simple_tag")
    o = Oracle(d, "f64")
    n = 64
    pv, lm, comm = random_states(d, n, rng)
    act = random_actions(o.act_dims, n, rng)
    a_pv, _, a_obs, _, _, _ = o.step(pv, lm, comm, act, 0)
    t = np.array([tx, ty])
    pv2, lm2 = pv.copy(), lm + t
    pv2[:, :, 0:2] += t
    b_pv, _, b_obs, _, _, _ = o.step(pv2, lm2, comm, act, 0)
    np.testing.assert_allclose(b_pv[:, :, 2:4], a_pv[:, :, 2:4], rtol=1e-7, atol=1e-7)      # same velocities
    np.testing.assert_allclose(b_pv[:, :, 0:2] - t, a_pv[:, :, 0:2], rtol=1e-7, atol=1e-7)
    # relative parts of the observation (everything but own position, columns 2:4) do not move
    od = o.obs_dims
    c0 = 0
    for k in od:
        keep = [c for c in range(k) if c not in (2, 3)]
        np.testing.assert_allclose(b_obs[:, c0:c0 + k][:, keep], a_obs[:, c0:c0 + k][:, keep], rtol=1e-6, atol=1e-6)
        c0 += k


@settings(max_examples=20, deadline=None)
@given(seed=SEEDS)
def test_contact_forces_are_equal_and_opposite(seed):
    """two spread agents, zero action, unclamped: total momentum change is zero (core.py:193-196)"""
    rng = np.random.RandomState(seed)
    d = descriptor("simple_spread_n3")
    o = Oracle(d, "f64")
    n = 128
    pv = np.zeros((n, 3, 4))
    pv[:, :, 0:2] = rng.uniform(-0.2, 0.2, (n, 3, 2))
    act = np.zeros((n, 15))
    act[:, [0, 5, 10]] = 1.0                                # no-op action for every agent
    npv, *_ = o.step(pv, rng.uniform(-1, 1, (n, 3, 2)), np.zeros((n, 3, 2)), act, 1)
    np.testing.assert_allclose(npv[:, :, 2:4].sum(axis=1), 0.0, atol=1e-9)
    assert np.abs(npv[:, :, 2:4]).max() > 0.1             # and there really were contacts


@settings(max_examples=20, deadline=None)
@given(seed=SEEDS)
def test_spread_reward_is_landmark_permutation_invariant_and_shared(seed):
    rng = np.random.RandomState(seed)
    d = descriptor("simple_spread_n6")
    o = Oracle(d, "f64")
    n = 32
    pv, lm, comm = random_states(d, n, rng)
    _, rew, _, info = o.observe(pv, lm, comm, flags=0)
    perm = rng.permutation(6)
    _, rew2, _, _ = o.observe(pv, lm[:, perm], comm, flags=0)
    np.testing.assert_allclose(rew2, rew, rtol=1e-12, atol=1e-12)
    _, shared, _, _ = o.observe(pv, lm, comm, flags=1)
    np.testing.assert_allclose(shared, np.repeat(rew.sum(1, keepdims=True), 6, 1), rtol=1e-12)
    assert (info[:, :, 1] >= 1).all()                      # the agent always "collides" with itself


@settings(max_examples=15, deadline=None)
@given(seed=SEEDS, tag=st.sampled_from(["simple_spread_n3", "simple_tag", "simple_world_comm", "simple_push"]))
def test_fp32_build_stays_within_north_star_tolerance(seed, tag):
    rng = np.random.RandomState(seed)
    d = descriptor(tag)
    o64, o32 = Oracle(d, "f64"), Oracle(d, "f32")
    n = 256
    pv, lm, comm = random_states(d, n, rng)
    pv, lm, comm = pv.astype(np.float32), lm.astype(np.float32), comm.astype(np.float32)
    movable = [bool(d.agent_movable[i]) for i in range(d.n_agents)]
    act = random_actions(o64.act_dims, n, rng, movable=movable).astype(np.float32)
    goal = rng.randint(0, d.n_landmarks, (n, 2)).astype(np.int32)[:, :1] if tag == "simple_push" else None
    a = o64.step(pv, lm, comm, act, 0, goal=goal)
    b = o32.step(pv, lm, comm, act, 0, goal=goal)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-5, atol=1e-6)     # state
    np.testing.assert_allclose(b[2], a[2], rtol=1e-5, atol=1e-6)     # observations


def test_flag_mismatch_accounting_bites():
    """helpers.explain_flag_mismatches (the replacement of the percentage budgets): a reward difference is accepted only
    if it is a whole number of contact quanta AND some pair of that world sits within 2e-6 of a threshold"""
    import pytest
    from helpers import explain_flag_mismatches
    a_size, l_size = [0.15, 0.15, 0.15], [0.05, 0.05, 0.05]
    pv = np.zeros((4, 3, 4))
    pv[:, 0, 0:2] = (-0.8, -0.8)
    pv[:, 1, 0:2] = (0.8, 0.8)
    pv[:, 2, 0:2] = (0.8, -0.8)
    lm = np.full((4, 3, 2), 5.0)                      # far away: no landmark threshold nearby
    pv[1, 1, 0:2] = (-0.8 + 0.3 + 3e-7, -0.8)         # world 1: agents 0 and 1 a hair outside contact (0.15 + 0.15)
    ref = np.zeros((4, 3))
    same = ref.copy()
    assert explain_flag_mismatches("simple_spread_n3", same, ref, None, None, pv, lm, a_size, l_size) == 0
    flipped = ref.copy()
    flipped[1] -= 2.0                                  # fp32 saw the contact: both agents lose 1, shared sum -2: explained
    assert explain_flag_mismatches("simple_spread_n3", flipped, ref, None, None, pv, lm, a_size, l_size) == 1
    wrong_world = ref.copy()
    wrong_world[2] -= 1.0                              # a whole quantum, but no pair of world 2 is near a threshold
    with pytest.raises(AssertionError):
        explain_flag_mismatches("simple_spread_n3", wrong_world, ref, None, None, pv, lm, a_size, l_size)
    not_quantum = ref.copy()
    not_quantum[1] -= 0.37                             # near a threshold, but not a multiple of the contact quantum
    with pytest.raises(AssertionError):
        explain_flag_mismatches("simple_spread_n3", not_quantum, ref, None, None, pv, lm, a_size, l_size)
    with pytest.raises(AssertionError):                # scenarios without indicator terms tolerate nothing
        explain_flag_mismatches("simple_push", flipped, ref, None, None, pv, lm, a_size, l_size)
