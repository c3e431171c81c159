"""GPU tests of the device-resident lockstep PPO loop (rollout glue; BASELINE configs[3] at test size):
segment returns vs the oracle per episode fragment, the minibatch gather, and one full rollout + update replayed step by step
through the CPU oracle (physics bit-exact, weights within the fp32 tolerance)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


@pytest.mark.parametrize("bootstrap", [False, True])
@pytest.mark.parametrize("use_gae", [0, 1])
def test_segment_returns_match_oracle_per_episode_fragment(gpu, O, use_gae, bootstrap):
    """Every episode fragment inside the segment follows the reference's trajectory recurrences bit for bit (PPOAgent.cs:414-498).
    bootstrap=False: the segment end is scored like a trajectory end.  bootstrap=True: a fragment still running at the segment
    end continues from V(s_T) -- identical to the oracle's recurrence over the fragment extended by one virtual step whose reward
    (MC) / value (GAE) is V(s_T), with that step dropped again."""
    torch = _torch()
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    rng = np.random.default_rng(use_gae)
    n, T = 37, 29
    hp = gpu.default_hyperparams()
    hp.use_gae = use_gae
    agent = gpu.PPOAgent(hp=hp, seed=1)
    rew = rng.normal(size=(T, n)).astype(np.float32)
    val = rng.normal(size=(T, n)).astype(np.float32)
    last = rng.normal(size=n).astype(np.float32)
    done = (rng.random((T, n)) < 0.15).astype(np.uint8)
    d = [torch.from_numpy(x).cuda() for x in (rew, val, done, last)]
    G = torch.empty(T, n, device="cuda")
    A = torch.empty(T, n, device="cuda")
    check(lib().wb_segment_returns_dev(agent._h, n, T, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]) if bootstrap else None, ptr(G), ptr(A)))
    torch.cuda.synchronize()
    G, A = G.cpu().numpy(), A.cpu().numpy()
    lam = hp.lambda_ if hasattr(hp, "lambda_") else 0.95
    open_fragments = 0
    for e in range(n):
        start = 0
        ends = list(np.flatnonzero(done[:, e])) + ([T - 1] if not done[T - 1, e] else [])
        for end in ends:  # [start, end] is one episode fragment: the reference's trajectory semantics apply to it alone
            r, v = rew[start:end + 1, e], val[start:end + 1, e]
            running = bootstrap and end == T - 1 and not done[T - 1, e]
            if running:  # extend by the virtual step, drop it afterwards
                open_fragments += 1
                r = np.append(r, last[e] if not use_gae else np.float32(0)).astype(np.float32)
                v = np.append(v, last[e]).astype(np.float32)
            rG, rA = O.gae(r, v, hp.gamma, lam) if use_gae else O.mc_returns(r, v, hp.gamma)
            if running:
                rG, rA = rG[:-1], rA[:-1]
            assert np.array_equal(G[start:end + 1, e].view(np.uint32), rG.view(np.uint32))
            assert np.array_equal(A[start:end + 1, e].view(np.uint32), rA.view(np.uint32))
            start = end + 1
    assert (open_fragments > 0) == bootstrap


def test_segment_advantages_normalised_like_the_reference(gpu, O):
    """hp.normalize_advantages: PPOAgent.Normalize (PPOAgent.cs:461-472) over the pool -- double-accumulated mean and
    standard deviation, divided by (std + clip epsilon).  The oracle sums left to right, the kernels in a fixed tree: the
    double sums differ by ~1e-16 relative before they are rounded to float, hence 2 ulp instead of bit-exact."""
    torch = _torch()
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    rng = np.random.default_rng(5)
    n, T = 700, 33
    hp = gpu.default_hyperparams()
    agent = gpu.PPOAgent(hp=hp, seed=1)
    rew = rng.normal(size=(T, n)).astype(np.float32)
    val = rng.normal(size=(T, n)).astype(np.float32)
    done = (rng.random((T, n)) < 0.1).astype(np.uint8)
    d = [torch.from_numpy(x).cuda() for x in (rew, val, done)]
    G, A, An = (torch.empty(T, n, device="cuda") for _ in range(3))
    check(lib().wb_segment_returns_dev(agent._h, n, T, ptr(d[0]), ptr(d[1]), ptr(d[2]), None, ptr(G), ptr(A)))
    hp.normalize_advantages = 1
    agent.set_hyperparams(hp)
    check(lib().wb_segment_returns_dev(agent._h, n, T, ptr(d[0]), ptr(d[1]), ptr(d[2]), None, ptr(G), ptr(An)))
    torch.cuda.synchronize()
    ref = O.normalize(A.cpu().numpy().reshape(-1).copy(), hp.epsilon)
    got = An.cpu().numpy().reshape(-1)
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-7)
    assert abs(float(got.mean())) < 1e-5 and abs(float(got.std()) - 1.0 / (1.0 + hp.epsilon / A.cpu().numpy().std())) < 1e-3
    # staged form (what a data-parallel caller runs, here with one rank): identical result
    A2 = A.clone()
    for stage in (0, 1, 2):
        check(lib().wb_normalize_advantages_dev(agent._h, stage, n * T, n * T, ptr(A2)))
    torch.cuda.synchronize()
    assert torch.equal(A2, An)


def test_gather_minibatch(gpu):
    torch = _torch()
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    agent = gpu.PPOAgent(seed=2)
    P, B = 1000, 333
    g = torch.Generator(device="cuda").manual_seed(0)
    pools = [torch.randn(P, 12, device="cuda", generator=g), torch.randn(P, 4, device="cuda", generator=g), torch.randn(P, 4, device="cuda", generator=g),
             torch.randn(P, device="cuda", generator=g), torch.randn(P, device="cuda", generator=g)]
    idx = torch.randperm(P, device="cuda", dtype=torch.int32, generator=g)[:B].contiguous()
    outs = [torch.empty(B, 12, device="cuda"), torch.empty(B, 4, device="cuda"), torch.empty(B, 4, device="cuda"), torch.empty(B, device="cuda"),
            torch.empty(B, device="cuda")]
    check(lib().wb_gather_minibatch_dev(agent._h, B, ptr(idx), *[ptr(p) for p in pools], *[ptr(o) for o in outs]))
    torch.cuda.synchronize()
    for p, o in zip(pools, outs):
        assert torch.equal(o, p[idx.long()])


def test_full_loop_replayed_through_the_oracle(gpu, O):
    """One rollout + update of VectorPPO at test size; the oracle replays it: physics from the recorded actions (bit-exact),
    log-probabilities / values from the recorded states (fp32 tolerance), then the same mini-batches through the per-sample
    reference update (weights within 1e-5)."""
    torch = _torch()
    n, T, B = 96, 12, 128
    vp = gpu.VectorPPO(n, horizon=T, minibatch_global=B, epochs=1, floor="Wood", seed=3, policy_variant=1)
    w_actor0, w_critic0 = vp.agent.actor.get_flat().copy(), vp.agent.critic.get_flat().copy()
    gen_state = vp.gen.get_state()
    stats = vp.iterate()
    assert stats["minibatches"] == (n * T) // B and stats["samples_trained"] == stats["minibatches"] * B
    S, A, LP = vp.states.cpu().numpy(), vp.actions.cpu().numpy(), vp.logp.cpu().numpy()
    R, D, V = vp.rewards.cpu().numpy(), vp.dones.cpu().numpy(), vp.values.cpu().numpy()
    # -- physics replay
    ref = O.EnvBatch(n, floor="Wood")
    obs = ref.get_obs()
    for t in range(T):
        assert np.array_equal(S[t].view(np.uint32), obs.view(np.uint32)), f"recorded state differs at t={t}"
        obs, rew, done = ref.step(A[t])
        assert np.array_equal(R[t].view(np.uint32), rew.view(np.uint32)) and np.array_equal(D[t], done)
    # -- policy replay: log-probabilities of the recorded (unclipped) actions and the value estimates
    actor, critic = O.Net(12, O.ACTOR_LAYERS), O.Net(12, O.CRITIC_LAYERS)
    actor.set_params(w_actor0)
    critic.set_params(w_critic0)
    mean = actor.forward(S.reshape(-1, 12))
    std = np.exp(np.float32(-1.0))
    lp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((A.reshape(-1, 4) - mean) / std) ** 2).astype(np.float32)
    np.testing.assert_allclose(LP.reshape(-1, 4), lp, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(V.reshape(-1), critic.forward(S.reshape(-1, 12))[:, 0], rtol=1e-5, atol=1e-5)
    # the trajectory stores the SAMPLED actions, not what Matrix.Clip handed to TakeActions (Environment.cs:78,87-88): some of the
    # 4608 draws leave [-1, 1], their log-probabilities above are those of the unclipped values, and the physics replay above
    # (the oracle clips inside its step, like Environment.Update) reproduced the rollout from them
    assert np.abs(A).max() > 1.0 and (np.abs(A) > 1.0).sum() >= 3
    # -- update replay with the same permutation
    G, ADV = vp.returns.cpu().numpy().reshape(-1), vp.adv.cpu().numpy().reshape(-1)
    g2 = torch.Generator(device="cuda")
    g2.set_state(gen_state)
    perm = torch.randperm(n * T, generator=g2, device="cuda", dtype=torch.int32).cpu().numpy()
    ohp = O.hyper_defaults()
    ohp.batch_size = B
    Sf, Af, LPf = S.reshape(-1, 12), A.reshape(-1, 4), LP.reshape(-1, 4)
    for j in range((n * T) // B):
        idx = perm[j * B:(j + 1) * B]
        O.ppo_train_batch(actor, critic, ohp, Sf[idx], Af[idx], LPf[idx], ADV[idx], G[idx], optimise=True)
    np.testing.assert_allclose(vp.agent.actor.get_flat(), actor.get_params(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(vp.agent.critic.get_flat(), critic.get_params(), rtol=0, atol=2e-5)
    assert not np.array_equal(vp.agent.actor.get_flat(), w_actor0)
