"""CPU tests that PIN the oracle (test infrastructure): hand-derived known answers from the reference's source constants,
the independent NumPy restatement bit for bit, structural invariants, and the committed golden rollout.
The reference ships no tests/golden vectors and cannot be run here (no .NET): parity is unpinned beyond these."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_appendix_d.json")))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_timestep_constants(O):
    assert float.hex(O.DT_FRAME) == "0x1.1111340000000p-6"
    dt_sub = np.float32(O.DT_FRAME) / np.float32(50)
    assert float.hex(float(dt_sub)) == "0x1.5d86a80000000p-12"


def test_initial_geometry_known_answers(O):
    env = O.EnvBatch(1)
    f, iv = env.get_state()
    f = f[0]
    assert np.array_equal(f[0:12].reshape(6, 2), np.array(KAT["lower_leg"], np.float32))
    assert np.array_equal(f[12:24].reshape(6, 2), np.array(KAT["upper_leg"], np.float32))
    assert np.array_equal(f[24:34].reshape(5, 2), np.array(KAT["body"], np.float32))
    assert np.array_equal(f[34:46], f[0:12]) and np.array_equal(f[46:58], f[12:24])  # legs start coincident
    cent = f[58:68].reshape(5, 2)
    assert np.array_equal(cent[0], KAT["lower_leg_centroid"]) and np.array_equal(cent[1], KAT["upper_leg_centroid"])
    assert np.array_equal(cent[2], KAT["body_centroid"])
    assert not f[68:92].any() and list(iv[0]) == [0, 0]
    np.testing.assert_array_equal(env.get_obs()[0], np.array(KAT["initial_obs"], np.float32))


def test_pole_from_size_and_reciprocal_centroid(O):
    v = np.zeros(12, np.float32)
    c = np.zeros(2, np.float32)
    O.lib().wo_pole_from_size(125.0, 830.0, 75.0, O._fp(v), O._fp(c))
    assert np.array_equal(v.reshape(6, 2), np.array(KAT["upper_leg"], np.float32))
    # FindCentroid multiplies by (1f/6f) instead of dividing (MonoGame Vector2 / float): 750 * (1/6f) rounds to 125 exactly
    assert c[0] == np.float32(750.0) * (np.float32(1.0) / np.float32(6.0)) == 125.0


def test_first_joint_step_known_answers(O):
    env = O.EnvBatch(1)
    env.take_actions(np.zeros((1, 4), np.float32))
    pt, jt = env.step_objects(O.DT_FRAME, 1, trace=True)
    assert jt[0, 0, 0]["active"] == 1 and jt[0, 0, 0]["depth"] == np.float32(KAT["first_joint_gap"])
    # J1 sees the Body already moved by half of J0's gap
    assert jt[0, 0, 1]["depth"] == np.float32(8.125)


def test_materials_table(O):
    for i, name in enumerate(O.MATERIALS):
        m = O.material(name)
        want = KAT["materials"][name]
        assert (np.float32(m.inverse_mass), np.float32(m.restitution), np.float32(m.friction)) == tuple(np.float32(x) for x in want)


def test_rotation_uses_double_trig_rounded_to_float(O):
    c = np.zeros(1, np.float32)
    s = np.zeros(1, np.float32)
    for ang in [0.0, 1e-4, -3.3e-3, 0.7, 3.0]:
        O.lib().wo_rotz(np.float32(ang), O._fp(c), O._fp(s))
        assert c[0] == np.float32(np.cos(np.float64(np.float32(ang)))) and s[0] == np.float32(np.sin(np.float64(np.float32(ang))))


@pytest.mark.parametrize("floor", ["Metal", "Wood", "SuperRubber", "Ice"])
def test_c_oracle_equals_numpy_restatement_bit_for_bit(O, floor):
    import np_oracle as P
    fm = O.material(floor)
    env = O.EnvBatch(1, floor=floor)
    ref = P.Environment(floor=(fm.inverse_mass, fm.restitution, fm.friction))
    rng = np.random.default_rng(hash(floor) % 1000)
    for t in range(12):
        a = rng.uniform(-1.3, 1.3, 4).astype(np.float32)
        obs, r, d = env.step(a[None])
        pobs, pr, pd = ref.step(a, O.DT_FRAME, auto_reset=True)
        f, iv = env.get_state()
        pf, piv = ref.flat_state()
        assert np.array_equal(bits(f[0]), bits(pf)), f"state differs at step {t}"
        assert np.array_equal(iv[0], piv) and np.array_equal(bits(obs[0]), bits(pobs))
        assert bits(r)[0] == bits(np.float32(pr))[0] and bool(d[0]) == pd


def test_c_oracle_equals_numpy_restatement_after_reset_order_flip(O):
    """After the first Reset the floor precedes the walker in the body list (Walker.cs:212-223)."""
    import np_oracle as P
    env = O.EnvBatch(1)
    ref = P.Environment()
    env.reset(0)
    ref.reset()
    _, iv = env.get_state()
    assert iv[0, 0] & O.FLAG_FLOOR_FIRST
    rng = np.random.default_rng(9)
    for t in range(8):
        a = rng.uniform(-1, 1, 4).astype(np.float32)
        env.step(a[None])
        ref.step(a, O.DT_FRAME, auto_reset=True)
    f, iv = env.get_state()
    pf, piv = ref.flat_state()
    assert np.array_equal(bits(f[0]), bits(pf)) and np.array_equal(iv[0], piv)


def test_sat_invariants(O):
    rng = np.random.default_rng(0)
    box = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], np.float32)
    hits = 0
    for _ in range(300):
        ang = rng.uniform(0, 6.28)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]], np.float32)
        a = (box * rng.uniform(0.5, 2)) @ R.T + rng.uniform(-2, 2, 2).astype(np.float32)
        b = box * rng.uniform(0.5, 2) + rng.uniform(-2, 2, 2).astype(np.float32)
        a, b = a.astype(np.float32), b.astype(np.float32)
        r1, n1, d1, ax1 = O.sat(a, b)
        r2, n2, d2, ax2 = O.sat(b, a)
        assert r1 == r2
        if r1:
            hits += 1
            assert d1 > 0 and abs(d1 - d2) <= 1e-5 * max(1, d1)
            assert np.allclose(n1, -n2, atol=1e-5)          # symmetric pair: antiparallel normals
            assert abs(np.linalg.norm(n1) - 1) < 1e-5
            assert np.dot(b.mean(0) - a.mean(0), n1) <= 1e-6  # normal points from B towards A
            pts = O.contacts(a, b, n1)
            assert 0 <= len(pts) <= 2
    assert hits > 50


def test_static_floor_never_moves_and_momentum_is_exchanged(O):
    env = O.EnvBatch(4, floor="Rubber")
    rng = np.random.default_rng(1)
    for _ in range(40):
        env.step(rng.uniform(-1, 1, (4, 4)).astype(np.float32), auto_reset=False)
    f, _ = env.get_state()
    assert np.isfinite(f).all()
    # lowest vertex never tunnels far below the floor top (y = 900): de-penetration works every substep
    assert f[:, 1:58:2].max() < 915


def test_golden_rollout_regression(O):
    g = np.load(os.path.join(HERE, "golden", "physics_rollout.npz"))
    env = O.EnvBatch(8, floor=[str(x) for x in g["floors"]])
    for t in range(g["actions"].shape[0]):
        obs, r, d = env.step(g["actions"][t])
        f, iv = env.get_state()
        assert np.array_equal(bits(f), bits(g["states"][t])), f"state differs from the golden fixture at step {t}"
        assert np.array_equal(iv, g["ints"][t]) and np.array_equal(bits(obs), bits(g["obs"][t]))
        assert np.array_equal(bits(r), bits(g["reward"][t])) and np.array_equal(d, g["done"][t])


def test_reward_quirks(O):
    """Appendix E: '+0.1' when the body is too LOW (sign bug), -40 on terminal, terminal is AABB-latched."""
    env = O.EnvBatch(1)
    rng = np.random.default_rng(2)
    seen_low_bonus = seen_terminal = False
    for _ in range(150):
        obs, r, d = env.step(rng.uniform(-1, 1, (1, 4)).astype(np.float32), auto_reset=False)
        if d[0]:
            seen_terminal = True
            assert r[0] < -30
            break
        if obs[0, 1] > 1.65:  # pre-step check uses the post-step h; the bonus shows up in r
            seen_low_bonus = seen_low_bonus or abs(r[0] - 0.1) < 1e-6 or r[0] > 0.09
    assert seen_terminal and seen_low_bonus


# ------------------------------------------------------------------ PPO oracle
def test_log_prob_and_sigma_known_answers(O):
    assert np.float32(np.exp(np.float32(-1.0))) == np.float32(KAT["sigma"])
    lp = O.lib().wo_log_prob(0.0, 1.0, 0.0)
    assert abs(lp + KAT["log_sqrt_2pi"]) < 1e-7
    assert O.Net(12, O.ACTOR_LAYERS).num_params == KAT["actor_params"]
    assert O.Net(12, O.CRITIC_LAYERS).num_params == KAT["critic_params"]


def test_oracle_gradient_matches_torch_autograd_of_the_reference_surrogate(O):
    """fp64 autograd of the reference's objective (per-DIMENSION ratio and clip, sum over dims, /B) vs the
    hand-derived per-sample gradient the oracle restates (PPOAgent.cs:248-326)."""
    import torch
    rng = np.random.default_rng(3)
    actor, critic = O.Net(12, O.ACTOR_LAYERS), O.Net(12, O.CRITIC_LAYERS)
    wa = (rng.normal(size=actor.num_params) * 0.2).astype(np.float32)
    wc = (rng.normal(size=critic.num_params) * 0.2).astype(np.float32)
    actor.set_params(wa)
    critic.set_params(wc)
    hp = O.hyper_defaults()
    n = hp.batch_size
    states = rng.normal(size=(n, 12)).astype(np.float32)
    std = float(np.exp(np.float32(-1)))
    mean = actor.forward(states)
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = -np.log(std) - 0.5 * np.log(2 * np.pi) - 0.5 * ((actions - mean) / std) ** 2
    old = (logp + 0.25 * rng.normal(size=(n, 4))).astype(np.float32)
    adv = rng.normal(size=n).astype(np.float32)
    ret = rng.normal(size=n).astype(np.float32)
    skipped, _, _ = O.ppo_train_batch(actor, critic, hp, states, actions, old, adv, ret, optimise=False)
    assert skipped == 0

    def unpack(flat, shapes):
        out, p = [], 0
        for o, i in shapes:
            W = torch.tensor(flat[p:p + o * i].reshape(o, i), dtype=torch.float64, requires_grad=True)
            p += o * i
            b = torch.tensor(flat[p:p + o], dtype=torch.float64, requires_grad=True)
            p += o
            out += [W, b]
        return out

    A = unpack(wa, [(64, 12), (64, 64), (4, 64)])
    Cc = unpack(wc, [(64, 12), (1, 64)])
    x = torch.tensor(states, dtype=torch.float64)
    lrelu = torch.nn.functional.leaky_relu
    mu = torch.tanh(lrelu(lrelu(x @ A[0].T + A[1], 0.2) @ A[2].T + A[3], 0.2) @ A[4].T + A[5])
    V = (lrelu(x @ Cc[0].T + Cc[1], 0.2) @ Cc[2].T + Cc[3])[:, 0]
    a = torch.tensor(actions, dtype=torch.float64)
    lp = -np.log(std) - 0.5 * np.log(2 * np.pi) - 0.5 * ((a - mu) / std) ** 2
    ratio = torch.exp(lp - torch.tensor(old, dtype=torch.float64))
    Adv = torch.tensor(adv, dtype=torch.float64)[:, None]
    surrogate = torch.minimum(ratio * Adv, torch.clamp(ratio, 1 - 0.3, 1 + 0.3) * Adv)
    loss = (-surrogate).sum() / n + ((V - torch.tensor(ret, dtype=torch.float64)) ** 2).sum() / n
    loss.backward()
    ga = np.concatenate([t.grad.numpy().ravel() for t in A])
    gc = np.concatenate([t.grad.numpy().ravel() for t in Cc])
    assert np.abs(actor.get_grads() - ga).max() / np.abs(ga).max() < 1e-4
    assert np.abs(critic.get_grads() - gc).max() / np.abs(gc).max() < 1e-4


def test_gae_quirk_next_gae_is_never_updated(O):
    r = np.array([1, 2, 3], np.float32)
    v = np.array([0.5, 0.25, 0.125], np.float32)
    G, A = O.gae(r, v, 0.9, 0.95)
    # delta_t = r_t + gamma * V_{t+1} - V_t with V_3 = 0; the lambda term is always zero in the reference (PPOAgent.cs:422-431)
    np.testing.assert_allclose(A, [1 + 0.9 * 0.25 - 0.5, 2 + 0.9 * 0.125 - 0.25, 3 - 0.125], rtol=1e-6)
    np.testing.assert_allclose(G, A + v, rtol=1e-6)


def test_mc_returns_and_normalize(O):
    r = np.array([1, 0, -1, 2], np.float32)
    v = np.zeros(4, np.float32)
    G, A = O.mc_returns(r, v, 0.9)
    np.testing.assert_allclose(G, [1 + 0.9 * (0 + 0.9 * (-1 + 0.9 * 2)), 0.9 * (-1 + 1.8), -1 + 1.8, 2], rtol=1e-6)
    z = O.normalize(A, 0.3)  # divides by (std + 0.3): the PPO clip epsilon doubles as the normaliser's epsilon (PPOAgent.cs:470)
    np.testing.assert_allclose(z, (A - A.mean()) / (A.std() + 0.3), rtol=1e-5)
