"""GPU parity tests for the PPO kernels through the C ABI vs the CPU oracle (per-sample Matrix-library restatement).
Tolerances (fp32, BASELINE.json north_star): gradients <= 1e-4 relative; forward / weights <= 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


VARIANTS = [0, 1]  # 0: tcgen05 tensor-core kernel (3xTF32), 1: fp32 CUDA-core kernel


def make_pair(gpu, O, seed=0, batch_size=64, variant=0):
    hp = gpu.default_hyperparams()
    hp.batch_size = batch_size
    agent = gpu.PPOAgent(hp=hp, seed=seed)
    agent.set_variant(variant)
    actor = O.Net(12, O.ACTOR_LAYERS)
    critic = O.Net(12, O.CRITIC_LAYERS)
    actor.set_params(agent.actor.get_flat())
    critic.set_params(agent.critic.get_flat())
    ohp = O.hyper_defaults()
    ohp.batch_size = batch_size
    return agent, actor, critic, ohp


def keep_clear_of_leaky_kinks(states, actor_flat, critic_flat, margin=1e-5):
    """LeakyReLU' jumps from 0.2 to 1 at a pre-activation of exactly 0: a sample with a hidden pre-activation within rounding
    distance of 0 may legitimately take either slope depending on the summation order (with B = 65536 one such sample moves the
    gradient by ~1e-3 relative).  Like the clip boundaries, such samples are replaced by a generic one (fp64 pre-activations)."""
    s = states.astype(np.float64)
    a = np.asarray(actor_flat, np.float64)
    c = np.asarray(critic_flat, np.float64)
    w1, b1 = a[:768].reshape(64, 12), a[768:832]
    w2, b2 = a[832:832 + 4096].reshape(64, 64), a[832 + 4096:832 + 4160]
    wc1, bc1 = c[:768].reshape(64, 12), c[768:832]
    z1 = s @ w1.T + b1
    z2 = np.maximum(0.2 * z1, z1) @ w2.T + b2
    zc = s @ wc1.T + bc1
    tight = np.minimum(np.minimum(np.abs(z1).min(1), np.abs(z2).min(1)), np.abs(zc).min(1)) < margin
    if tight.any():
        states[tight] = states[np.flatnonzero(~tight)[0]]
    return states


def synth_batch(rng, actor, n, O=None, std=np.exp(np.float32(-1.0)), critic_flat=None):
    """cfg3-style synthetic minibatch: scaled observations, actions near the mean, perturbed old log-probs."""
    scale = np.array([1, 1, 1, 1, 1, 1, 0.1, 0.1, 0.5, 0.5, 0.5, 0.5], np.float32)
    shift = np.array([0.14, 1.6, 0.13, 1.7, 0.13, 1.7, 0, 0, 0, 0, 0, 0], np.float32)
    states = (rng.normal(size=(n, 12)).astype(np.float32) * scale * 0.3 + shift).astype(np.float32)
    if critic_flat is not None:
        states = keep_clear_of_leaky_kinks(states, actor.get_params(), critic_flat)
    mean = actor.forward(states)
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old_logp = (logp + 0.25 * rng.normal(size=(n, 4))).astype(np.float32)  # ratios straddle 1 +- 0.3
    ratio = np.exp(logp.astype(np.float64) - old_logp)  # keep clear of the clip discontinuities (see the cross-kernel test)
    old_logp[(np.abs(ratio - 1.3) < 1e-3) | (np.abs(ratio - 0.7) < 1e-3)] += np.float32(0.01)
    adv = rng.normal(size=n).astype(np.float32)
    ret = (5 * rng.normal(size=n)).astype(np.float32)
    return states, actions, old_logp, adv, ret


def rel_err(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("variant", VARIANTS)
def test_forward_matches_oracle(gpu, O, variant):
    agent, actor, critic, _ = make_pair(gpu, O, 1, variant=variant)
    rng = np.random.default_rng(1)
    for n in (1, 63, 64, 65, 1000):
        s = rng.normal(size=(n, 12)).astype(np.float32)
        mean, value = agent.FeedForward(s)
        np.testing.assert_allclose(mean, actor.forward(s), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(value, critic.forward(s)[:, 0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", VARIANTS)
def test_sample_actions_with_injected_uniforms(gpu, O, variant):
    agent, actor, critic, ohp = make_pair(gpu, O, 2, variant=variant)
    rng = np.random.default_rng(2)
    n = 300
    s = rng.normal(size=(n, 12)).astype(np.float32)
    u = rng.random((n, 4, 2)).astype(np.float32)
    u[0, 0, 0] = 0.0  # the reference maps u1 == 0 to 1 (NormalDistribution.cs:16)
    a, lp, mu, std = agent.SampleActions(s, u)
    for i in range(0, n, 7):
        ra, rlp, rmu = O.sample_actions(actor, ohp, s[i], u[i].ravel())
        np.testing.assert_allclose(mu[i], rmu, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(a[i], ra, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(lp[i], rlp, rtol=1e-4, atol=2e-5)
    assert np.allclose(std, np.exp(np.float32(-1)))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n,batch_size", [(64, 64), (100, 64), (4096, 4096), (20000, 20000)])
def test_ppo_gradient_matches_oracle(gpu, O, n, batch_size, variant):
    agent, actor, critic, ohp = make_pair(gpu, O, 3, batch_size, variant=variant)
    rng = np.random.default_rng(n)
    batch = synth_batch(rng, actor, n, critic_flat=critic.get_params())
    closs, aloss, skipped = agent.Gradients(*batch)
    rskip, rcl, ral = O.ppo_train_batch(actor, critic, ohp, *batch, optimise=False)
    assert skipped == rskip == 0
    assert rel_err(agent.actor.get_grads(), actor.get_grads()) < 1e-4
    assert rel_err(agent.critic.get_grads(), critic.get_grads()) < 1e-4
    errs = per_layer_rel_err(agent.actor.get_grads(), actor.get_grads(), ACTOR_BLOCKS)
    errs.update(per_layer_rel_err(agent.critic.get_grads(), critic.get_grads(), CRITIC_BLOCKS))
    assert max(errs.values()) < 1e-4, errs
    assert abs(closs - rcl) <= 1e-4 * max(1.0, abs(rcl)) and abs(aloss - ral) <= 1e-4 * max(1.0, abs(ral))


# dense layers inside the flat parameter vectors: (name, offset, count)
ACTOR_BLOCKS = [("actor W1", 0, 768), ("actor b1", 768, 64), ("actor W2", 832, 4096), ("actor b2", 4928, 64), ("actor W3", 4992, 256),
                ("actor b3", 5248, 4)]
CRITIC_BLOCKS = [("critic W1", 0, 768), ("critic b1", 768, 64), ("critic W2", 832, 64), ("critic b2", 896, 1)]


def per_layer_rel_err(got, ref, blocks):
    """max |got - ref| of every dense layer's dW / db, normalised by THAT block's own max |ref| (a small-magnitude layer is held
    to the same relative tolerance as the largest one)."""
    out = {}
    for name, off, cnt in blocks:
        g, r = got[off:off + cnt], ref[off:off + cnt]
        out[name] = float(np.abs(g - r).max() / (np.abs(r).max() + 1e-30))
    return out


@pytest.mark.parametrize("variant", VARIANTS)
def test_ppo_gradient_matches_oracle_on_the_bench_minibatch(gpu, O, variant):
    """BASELINE configs[2] as bench.py times it: ONE 65 536-sample minibatch drawn by the bench's own generator
    (workloads.ppo_minibatch: full-spread observations, old log-probabilities perturbed by N(0, 0.1)), batch_size = 65 536,
    through wb_ppo_grad against the per-sample oracle (PPOAgent.cs:218-346) -- every dense layer's dW and db within 1e-4 of the
    oracle relative to that layer's own largest entry.  Samples that sit within rounding distance of a discontinuity of the
    loss (clip boundary, LeakyReLU kink) are replaced first: there either branch is legitimate (see keep_clear_of_leaky_kinks)."""
    import workloads
    n = 65536
    agent, actor, critic, ohp = make_pair(gpu, O, 42, n, variant=variant)
    rng = np.random.default_rng(1234 + 1)
    states = keep_clear_of_leaky_kinks(workloads.ppo_states(rng, n), actor.get_params(), critic.get_params())
    states, actions, old_logp, adv, ret = workloads.ppo_minibatch(rng, n, actor.forward, states=states)
    mean = actor.forward(states)
    std = np.exp(np.float32(-1.0))
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float64)
    ratio = np.exp(logp - old_logp)
    old_logp[(np.abs(ratio - 1.3) < 1e-3) | (np.abs(ratio - 0.7) < 1e-3)] += np.float32(0.01)
    closs, aloss, skipped = agent.Gradients(states, actions, old_logp, adv, ret)
    rskip, rcl, ral = O.ppo_train_batch(actor, critic, ohp, states, actions, old_logp, adv, ret, optimise=False)
    assert skipped == rskip == 0
    errs = per_layer_rel_err(agent.actor.get_grads(), actor.get_grads(), ACTOR_BLOCKS)
    errs.update(per_layer_rel_err(agent.critic.get_grads(), critic.get_grads(), CRITIC_BLOCKS))
    worst = max(errs, key=errs.get)
    assert errs[worst] < 1e-4, f"{worst}: {errs[worst]:.3e} (all layers: {errs})"
    assert abs(closs - rcl) <= 1e-4 * max(1.0, abs(rcl)) and abs(aloss - ral) <= 1e-4 * max(1.0, abs(ral))
    # clipped and unclipped samples are both present in numbers: the minibatch exercises both branches of the surrogate
    clipped = ((ratio > 1.3) | (ratio < 0.7)).mean()
    assert 0.001 < clipped < 0.5


@pytest.mark.parametrize("variant", VARIANTS)
def test_skipped_sample_when_old_probability_underflows(gpu, O, variant):
    agent, actor, critic, ohp = make_pair(gpu, O, 4, variant=variant)
    rng = np.random.default_rng(4)
    states, actions, old_logp, adv, ret = synth_batch(rng, actor, 64)
    old_logp[5, 2] = -200.0  # exp underflows to 0 -> HadamardDivision throws -> sample skipped (PPOAgent.cs:286-290)
    closs, aloss, skipped = agent.Gradients(states, actions, old_logp, adv, ret)
    rskip, rcl, ral = O.ppo_train_batch(actor, critic, ohp, states, actions, old_logp, adv, ret, optimise=False)
    assert skipped == rskip == 1
    assert rel_err(agent.actor.get_grads(), actor.get_grads()) < 1e-4
    assert rel_err(agent.critic.get_grads(), critic.get_grads()) < 1e-4


@pytest.mark.parametrize("variant", VARIANTS)
def test_adam_steps_track_oracle(gpu, O, variant):
    agent, actor, critic, ohp = make_pair(gpu, O, 5, variant=variant)
    rng = np.random.default_rng(5)
    for it in range(5):
        batch = synth_batch(rng, actor, 64)
        agent.TrainBatch(*batch)
        O.ppo_train_batch(actor, critic, ohp, *batch, optimise=True)
        assert np.abs(agent.actor.get_flat() - actor.get_params()).max() < 2e-5, f"actor weights diverged at Adam step {it}"
        assert np.abs(agent.critic.get_flat() - critic.get_params()).max() < 2e-5
    m, v, iters = agent.actor.get_adam()
    rm, rv, riters = actor.get_adam()
    assert list(iters) == list(riters) == [5, 5, 5]
    assert rel_err(m, rm) < 1e-3 and rel_err(v, rv) < 1e-3


def test_adam_is_bit_exact_given_identical_gradients(gpu, O):
    """Adam itself is restated op for op: from identical (m, v, grads) one step matches the oracle to the last bit."""
    agent, actor, critic, ohp = make_pair(gpu, O, 6)
    rng = np.random.default_rng(6)
    batch = synth_batch(rng, actor, 64)
    agent.Gradients(*batch)
    g_actor, g_critic = agent.actor.get_grads(), agent.critic.get_grads()
    # put the GPU's own gradients into the oracle (dW/db are exposed through a zero-LR trick: run feedback-free Adam)
    # one Adam step restated in numpy fp32 with the reference's operation order (DenseLayer.cs:125-159)
    def adam_np(w, g, m, v, t, hp):
        f = np.float32
        m2 = (f(1) - f(hp.beta1)) * g + f(hp.beta1) * m
        v2 = f(hp.beta2) * v + (f(1) - f(hp.beta2)) * (g * g)
        c1 = f(1.0 - float(np.float32(hp.beta1)) ** t)
        c2 = f(1.0 - float(np.float32(hp.beta2)) ** t)
        return w - f(hp.alpha) * ((m2 / c1) / (np.sqrt(v2 / c2) + f(hp.adam_epsilon)))
    w0a, w0c = agent.actor.get_flat(), agent.critic.get_flat()
    agent.Optimise()
    z = np.zeros_like(w0a)
    exp_a = adam_np(w0a, g_actor, z, z, 1, ohp)
    zc = np.zeros_like(w0c)
    exp_c = adam_np(w0c, g_critic, zc, zc, 1, ohp)
    assert np.array_equal(agent.actor.get_flat().view(np.uint32), exp_a.astype(np.float32).view(np.uint32))
    assert np.array_equal(agent.critic.get_flat().view(np.uint32), exp_c.astype(np.float32).view(np.uint32))


def test_returns_and_advantages(gpu, O):
    rng = np.random.default_rng(7)
    r = rng.normal(size=777).astype(np.float32)
    v = rng.normal(size=777).astype(np.float32)
    import ctypes as C
    for use_gae, norm in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        hp = gpu.default_hyperparams()
        hp.use_gae, hp.normalize_advantages = use_gae, norm
        agent = gpu.PPOAgent(hp=hp, seed=0)
        G = np.empty_like(r)
        A = np.empty_like(r)
        from ppo_bipedalwalker_b200._lib import check, lib, ptr
        check(lib().wb_returns_advantages(agent._h, r.size, ptr(r), ptr(v), ptr(G), ptr(A)))
        rG, rA = (O.gae(r, v, 0.9, 0.95) if use_gae else O.mc_returns(r, v, 0.9))
        if norm:
            rA = O.normalize(rA, 0.3)
        assert np.array_equal(G.view(np.uint32), rG.view(np.uint32))
        assert np.array_equal(A.view(np.uint32), rA.view(np.uint32))


def test_weights_file_roundtrip(gpu, O):
    agent = gpu.PPOAgent(seed=8)
    critic_lines, actor_lines = agent.Save()
    assert actor_lines[0] == gpu.DEFAULT_ACTOR and critic_lines[0] == gpu.DEFAULT_CRITIC
    assert len(actor_lines) == 4 and actor_lines[1].startswith("W ") and " B " in actor_lines[1]
    other = gpu.PPOAgent(seed=9)
    assert other.actor.Load(actor_lines) and other.critic.Load(critic_lines)
    assert np.array_equal(other.actor.get_flat(), agent.actor.get_flat())
    assert not other.actor.Load(critic_lines)  # structure line mismatch -> ignored like NeuralNetwork.Load
    # a file with fewer dense lines than layers updates the leading layers only (NeuralNetwork.cs:110-113)
    third = gpu.PPOAgent(seed=10)
    before = third.actor.get_flat().copy()
    assert third.actor.Load(actor_lines[:2])
    after = third.actor.get_flat()
    assert np.array_equal(after[:832], agent.actor.get_flat()[:832]) and np.array_equal(after[832:], before[832:])
    # ValidateWeights: one unparsable token anywhere -> nothing is loaded (NeuralNetwork.cs:118-156)
    bad = list(actor_lines)
    bad[3] = bad[3].replace(" B ", " B x", 1)
    fourth = gpu.PPOAgent(seed=11)
    before = fourth.actor.get_flat().copy()
    assert not fourth.actor.Load(bad) and np.array_equal(fourth.actor.get_flat(), before)
    assert not fourth.actor.Load(actor_lines + [actor_lines[1]])  # more lines than layers (the reference throws)


# user-edited Hyperparameters.ActorNeuralNetwork / CriticNeuralNetwork strings (PPOAgent.cs:96-143): (state, action, actor, critic)
TOPOLOGIES = [
    (12, 4, "Input |32| (ReLU) |4| (TanH) Output", "Input |16| (ReLU) |16| (TanH) |1| Output"),
    (12, 4, "Input |128| (LeakyReLU) |48| (ReLU) |80| (TanH) |4| Output", "Input |96| (ReLU) |1| Output"),
    (7, 3, "Input |20| (TanH) |3| Output", "Input |5| (LeakyReLU) |9| (LeakyReLU) |1| (ReLU) Output"),
    (12, 4, "Input |4| Output", "Input |1| Output"),
]


def make_topology_pair(gpu, O, topo, seed, batch_size):
    sdim, adim, actor_dsl, critic_dsl = topo
    hp = gpu.default_hyperparams()
    hp.batch_size = batch_size
    agent = gpu.PPOAgent(sdim, adim, hp=hp, actor=actor_dsl, critic=critic_dsl, seed=seed)
    assert agent.actor.structure == actor_dsl and agent.critic.structure == critic_dsl  # parsed, not replaced by the defaults
    actor = O.Net(sdim, gpu.ParseLayers(actor_dsl))
    critic = O.Net(sdim, gpu.ParseLayers(critic_dsl))
    actor.set_params(agent.actor.get_flat())
    critic.set_params(agent.critic.get_flat())
    ohp = O.hyper_defaults()
    ohp.batch_size = batch_size
    return agent, actor, critic, ohp


def topology_batch(rng, actor, n, sdim, adim):
    states = rng.normal(size=(n, sdim)).astype(np.float32)
    mean = actor.forward(states)
    std = np.exp(np.float32(-1.0))
    actions = (mean + std * rng.normal(size=(n, adim))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old_logp = (logp + 0.2 * rng.normal(size=(n, adim))).astype(np.float32)
    ratio = np.exp(logp.astype(np.float64) - old_logp)
    old_logp[(np.abs(ratio - 1.3) < 1e-3) | (np.abs(ratio - 0.7) < 1e-3)] += np.float32(0.01)
    return states, actions, old_logp, rng.normal(size=n).astype(np.float32), (3 * rng.normal(size=n)).astype(np.float32)


@pytest.mark.parametrize("topo", TOPOLOGIES)
def test_any_dsl_topology_runs_on_the_gpu_and_matches_the_oracle(gpu, O, topo):
    """Networks other than the defaults -- ReLU, mixed widths up to 128, other state / action sizes, activation-free and
    activation-terminated stacks -- run on the any-topology kernel: forward, sampling, PPOAgent.Train(Batch) gradients per dense
    layer, losses, and five Adam steps against the per-sample oracle (NeuralNetwork.cs:52-82, DenseLayer.cs:82-159,
    ActivationLayer.cs:32-73, PPOAgent.cs:218-346)."""
    sdim, adim = topo[0], topo[1]
    n = 333
    agent, actor, critic, ohp = make_topology_pair(gpu, O, topo, 21, n)
    rng = np.random.default_rng(21)
    s = rng.normal(size=(n, sdim)).astype(np.float32)
    mean, value = agent.FeedForward(s)
    np.testing.assert_allclose(mean, actor.forward(s), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(value, critic.forward(s)[:, 0], rtol=1e-5, atol=1e-6)
    u = rng.random((n, adim, 2)).astype(np.float32)
    a, lp, mu, std = agent.SampleActions(s, u)
    for i in range(0, n, 37):
        ra, rlp, rmu = O.sample_actions(actor, ohp, s[i], u[i].ravel())
        np.testing.assert_allclose(a[i], ra, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(lp[i], rlp, rtol=1e-4, atol=2e-5)
    batch = topology_batch(rng, actor, n, sdim, adim)
    closs, aloss, skipped = agent.Gradients(*batch)
    rskip, rcl, ral = O.ppo_train_batch(actor, critic, ohp, *batch, optimise=False)
    assert skipped == rskip == 0
    for got, ref, net in ((agent.actor.get_grads(), actor.get_grads(), agent.actor), (agent.critic.get_grads(), critic.get_grads(), agent.critic)):
        p = 0
        for out, inp in net.shapes:  # every dense layer's dW and db against that layer's own largest entry
            for cnt in (out * inp, out):
                g, r = got[p:p + cnt], ref[p:p + cnt]
                p += cnt
                assert np.abs(g - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-12), (net.structure, out, inp)
    assert abs(closs - rcl) <= 1e-4 * max(1.0, abs(rcl)) and abs(aloss - ral) <= 1e-4 * max(1.0, abs(ral))
    for it in range(5):
        b = topology_batch(rng, actor, n, sdim, adim)
        agent.TrainBatch(*b)
        O.ppo_train_batch(actor, critic, ohp, *b, optimise=True)
        assert np.abs(agent.actor.get_flat() - actor.get_params()).max() < 2e-5, f"actor weights diverged at Adam step {it}"
        assert np.abs(agent.critic.get_flat() - critic.get_params()).max() < 2e-5
    m, v, iters = agent.actor.get_adam()
    assert list(iters) == [5] * len(agent.actor.shapes)


def test_the_three_policy_kernels_agree_on_the_default_networks(gpu, O):
    """The any-topology kernel (variant 2) is a third implementation of the default networks: same gradients as the tensor-core
    (0) and CUDA-core (1) kernels and as the oracle, on a batch with a skipped sample."""
    n = 2000
    grads = {}
    for variant in (0, 1, 2):
        agent, actor, critic, ohp = make_pair(gpu, O, 7, n, variant=variant)
        rng = np.random.default_rng(7)
        batch = list(synth_batch(rng, actor, n, critic_flat=critic.get_params()))
        batch[2][11, 1] = -200.0  # old probability underflows -> sample skipped (PPOAgent.cs:286-290)
        closs, aloss, skipped = agent.Gradients(*batch)
        assert skipped == 1
        grads[variant] = np.concatenate([agent.actor.get_grads(), agent.critic.get_grads()])
    rskip, rcl, ral = O.ppo_train_batch(actor, critic, ohp, *batch, optimise=False)
    ref = np.concatenate([actor.get_grads(), critic.get_grads()])
    for variant in (0, 1, 2):
        assert rel_err(grads[variant], ref) < 1e-4, variant


def test_single_launch_train_step_matches_gradient_then_adam(gpu, O):
    """wb_ppo_train_dev on one process runs PPOAgent.Train(Batch) as ONE launch (the tensor-core kernel reduces its partials behind
    a grid barrier and applies Adam); the result must equal wb_ppo_grad_dev + wb_adam_step and track the oracle, for a grid of one
    CTA (n = 64), a partial grid (n = 5000) and the full persistent grid (n = 65536), three steps each."""
    import torch
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    for n in (64, 5000, 65536):
        fused, actor, critic, ohp = make_pair(gpu, O, 31, n, variant=0)
        split, _, _, _ = make_pair(gpu, O, 31, n, variant=0)
        rng = np.random.default_rng(n)
        for step in range(3):
            batch = synth_batch(rng, actor, n, critic_flat=critic.get_params())
            dev = [torch.from_numpy(x).cuda() for x in batch]
            check(lib().wb_ppo_train_dev(fused._h, n, *[ptr(t) for t in dev]))
            check(lib().wb_ppo_grad_dev(split._h, n, *[ptr(t) for t in dev]))
            check(lib().wb_adam_step(split._h))
            fused.sync()
            split.sync()
            O.ppo_train_batch(actor, critic, ohp, *batch, optimise=True)
            for net_f, net_s, net_o in ((fused.actor, split.actor, actor), (fused.critic, split.critic, critic)):
                np.testing.assert_allclose(net_f.get_flat(), net_s.get_flat(), rtol=0, atol=2e-6)   # same gradients, other summation order
                np.testing.assert_allclose(net_f.get_flat(), net_o.get_params(), rtol=0, atol=3e-5)
                assert rel_err(net_f.get_grads(), net_s.get_grads()) < 1e-5
        _, _, iters = fused.actor.get_adam()
        assert list(iters) == [3, 3, 3]


@pytest.mark.parametrize("variant", VARIANTS)
def test_indexed_train_step_equals_gather_then_train(gpu, O, variant):
    """wb_ppo_train_indexed_dev trains on rows index[] of a rollout pool (PPOAgent.CreateBatches, PPOAgent.cs:501-540, fused into
    the tensor-core kernel's prefetch; the other kernels gather first): same weights, bit for bit, as wb_gather_minibatch_dev
    followed by wb_ppo_train_dev, and the oracle's update on the gathered rows."""
    import torch
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    pool_n, n = 3000, 1000
    direct, actor, critic, ohp = make_pair(gpu, O, 41, n, variant=variant)
    gathered, _, _, _ = make_pair(gpu, O, 41, n, variant=variant)
    rng = np.random.default_rng(41)
    pool = synth_batch(rng, actor, pool_n, critic_flat=critic.get_params())
    dev_pool = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in pool]
    mb = [torch.empty((n,) + tuple(t.shape[1:]), dtype=torch.float32, device="cuda") for t in dev_pool]
    for step in range(3):
        idx = rng.permutation(pool_n)[:n].astype(np.int32)
        dev_idx = torch.from_numpy(idx).cuda()
        check(lib().wb_ppo_train_indexed_dev(direct._h, n, ptr(dev_idx), *[ptr(t) for t in dev_pool]))
        check(lib().wb_gather_minibatch_dev(gathered._h, n, ptr(dev_idx), *[ptr(t) for t in dev_pool], *[ptr(t) for t in mb]))
        check(lib().wb_ppo_train_dev(gathered._h, n, *[ptr(t) for t in mb]))
        direct.sync()
        gathered.sync()
        O.ppo_train_batch(actor, critic, ohp, *[x[idx] for x in pool], optimise=True)
        for net_d, net_g, net_o in ((direct.actor, gathered.actor, actor), (direct.critic, gathered.critic, critic)):
            np.testing.assert_array_equal(net_d.get_flat(), net_g.get_flat())
            np.testing.assert_array_equal(net_d.get_grads(), net_g.get_grads())
            np.testing.assert_allclose(net_d.get_flat(), net_o.get_params(), rtol=0, atol=3e-5)


def test_host_train_step_reads_pinned_buffers_in_place(gpu, O):
    """wb_ppo_train (PPOAgent.TrainBatch) from page-locked host arrays (read in place by the kernel), from pageable ones (staged
    copy) and wb_ppo_grad + wb_adam_step must leave identical weights, losses and skip counts; an odd sample count exercises the
    ragged last tile."""
    import torch
    n = 4321
    agents = [make_pair(gpu, O, 51, n, variant=0) for _ in range(3)]
    pinned_agent, actor, critic, ohp = agents[0]
    pageable_agent, split_agent = agents[1][0], agents[2][0]
    rng = np.random.default_rng(51)
    for step in range(2):
        batch = list(synth_batch(rng, actor, n, critic_flat=critic.get_params()))
        batch[2][7, 2] = -200.0  # one skipped sample (PPOAgent.cs:286-290)
        pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy() for x in batch]
        out_pinned = pinned_agent.TrainBatch(*pinned)
        out_pageable = pageable_agent.TrainBatch(*batch)
        out_split = split_agent.Gradients(*batch)
        split_agent.Optimise()
        assert out_pinned == out_pageable and out_pinned[2] == 1
        assert out_split[2] == 1
        np.testing.assert_allclose(out_pinned[:2], out_split[:2], rtol=1e-5)
        O.ppo_train_batch(actor, critic, ohp, *batch, optimise=True)
        for which in ("actor", "critic"):
            a, b, c = (getattr(x, which).get_flat() for x in (pinned_agent, pageable_agent, split_agent))
            np.testing.assert_array_equal(a, b)
            np.testing.assert_allclose(a, c, rtol=0, atol=2e-6)
            np.testing.assert_allclose(a, {"actor": actor, "critic": critic}[which].get_params(), rtol=0, atol=3e-5)


@pytest.mark.timeout(120)
def test_two_policies_on_two_streams_train_concurrently(gpu):
    """The one-launch train step holds a grid barrier, so every CTA must be resident at once: it is a COOPERATIVE launch, and two
    policies that train on different streams are run one after the other by the driver instead of deadlocking with half of the
    SMs each.  Their results equal the same updates issued on one stream."""
    import torch
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    n = 65536
    hp = gpu.default_hyperparams()
    hp.batch_size = n
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a1, a2 = gpu.PPOAgent(hp=hp, seed=1, stream=s1.cuda_stream), gpu.PPOAgent(hp=hp, seed=2, stream=s2.cuda_stream)
    b1, b2 = gpu.PPOAgent(hp=hp, seed=1), gpu.PPOAgent(hp=hp, seed=2)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    dev = [torch.randn(n, 12, device="cuda", generator=g), 0.3 * torch.randn(n, 4, device="cuda", generator=g),
           -0.5 * torch.rand(n, 4, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g),
           torch.randn(n, device="cuda", generator=g)]
    torch.cuda.synchronize()
    for _ in range(40):
        check(lib().wb_ppo_train_dev(a1._h, n, *[ptr(t) for t in dev]))
        check(lib().wb_ppo_train_dev(a2._h, n, *[ptr(t) for t in dev]))
    for _ in range(40):
        check(lib().wb_ppo_train_dev(b1._h, n, *[ptr(t) for t in dev]))
    for _ in range(40):
        check(lib().wb_ppo_train_dev(b2._h, n, *[ptr(t) for t in dev]))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(a1.actor.get_flat(), b1.actor.get_flat())
    np.testing.assert_array_equal(a2.critic.get_flat(), b2.critic.get_flat())


def test_topologies_outside_the_kernels_fail_loudly(gpu):
    with pytest.raises(gpu.WalkerB200Error):  # wider than the 128 the kernels cover: refused, never a silent fallback
        gpu.PPOAgent(actor="Input |256| (ReLU) |4| (TanH) Output")
    # a bad DSL string or a wrong output size falls back to the default network like PPOAgent.CreateNetworks (PPOAgent.cs:41-93)
    agent = gpu.PPOAgent(actor="Input |32| (ReLU) |5| Output", critic="Input 64 Output")
    assert agent.actor.structure == gpu.DEFAULT_ACTOR and agent.critic.structure == gpu.DEFAULT_CRITIC


def test_train_one_episode_end_to_end(gpu, O):
    """Environment + PPOAgent.Train(Trajectory) in the reference's loop shape (Environment.Update / TrainNetworks)."""
    env = gpu.Environment(1)
    agent = gpu.PPOAgent(seed=10)
    traj = gpu.Trajectory()
    state = env.InitialState()
    for _ in range(200):
        a, lp, _, _ = agent.SampleActions(state)
        traj.States.append(state[0].copy())
        state, reward, terminal = env.Update(gpu.DT_FRAME, a, auto_reset=False)
        traj.Actions.append(a[0])
        traj.LogProbabilities.append(lp[0])
        traj.Rewards.append(float(reward[0]))
        if terminal[0]:
            break
    before = agent.actor.get_flat().copy()
    agent.Train(traj)
    if len(traj.States) >= 64:
        assert not np.array_equal(before, agent.actor.get_flat())
    assert np.isfinite(agent.actor.get_flat()).all()


def test_tensor_core_and_cuda_core_kernels_agree(gpu, O):
    """Two independent kernels for the same path (tcgen05 3xTF32 vs fp32 FMA) on a 65536-sample minibatch."""
    rng = np.random.default_rng(11)
    hp = gpu.default_hyperparams()
    hp.batch_size = 65536
    a0 = gpu.PPOAgent(hp=hp, seed=12)
    a1 = gpu.PPOAgent(hp=hp, seed=12)
    a1.set_variant(1)
    actor = O.Net(12, O.ACTOR_LAYERS)
    actor.set_params(a0.actor.get_flat())
    n = 65536
    states = keep_clear_of_leaky_kinks(rng.normal(size=(n, 12)).astype(np.float32), a0.actor.get_flat(), a0.critic.get_flat())
    mean, _ = a1.FeedForward(states)
    std = np.exp(np.float32(-1.0))
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old = (logp + 0.2 * rng.normal(size=(n, 4))).astype(np.float32)
    # the clipped surrogate is discontinuous at ratio = 1 +- epsilon: a sample within rounding distance of a boundary may
    # legitimately take either branch, so keep every ratio at least 1e-3 away from both boundaries
    ratio = np.exp(logp.astype(np.float64) - old)
    near = (np.abs(ratio - 1.3) < 1e-3) | (np.abs(ratio - 0.7) < 1e-3)
    old[near] += np.float32(0.01)
    adv = rng.normal(size=n).astype(np.float32)
    ret = (5 * rng.normal(size=n)).astype(np.float32)
    l0 = a0.Gradients(states, actions, old, adv, ret)
    l1 = a1.Gradients(states, actions, old, adv, ret)
    assert rel_err(a0.actor.get_grads(), a1.actor.get_grads()) < 1e-4
    assert rel_err(a0.critic.get_grads(), a1.critic.get_grads()) < 1e-4
    assert abs(l0[0] - l1[0]) <= 1e-4 * max(1, abs(l1[0])) and abs(l0[1] - l1[1]) <= 1e-4 * max(1, abs(l1[1]))
    m0, v0 = a0.FeedForward(states[:5000])
    m1, v1 = a1.FeedForward(states[:5000])
    np.testing.assert_allclose(m0, m1, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(v0, v1, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("variant", VARIANTS)
def test_fused_train_step_tracks_oracle(gpu, O, variant):
    """wb_ppo_train_dev: gradient kernel + ONE kernel that reduces the per-CTA partials and applies DenseLayer.Adam; five
    consecutive minibatches against the oracle's per-sample update, and against the unfused wb_ppo_grad_dev + wb_adam_step."""
    import torch
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    agent, actor, critic, ohp = make_pair(gpu, O, 21, batch_size=512, variant=variant)
    twin, _, _, _ = make_pair(gpu, O, 21, batch_size=512, variant=variant)
    rng = np.random.default_rng(21)
    for it in range(5):
        batch = synth_batch(rng, actor, 512, critic_flat=critic.get_params())
        dev = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch]
        check(lib().wb_ppo_train_dev(agent._h, 512, *[ptr(t) for t in dev]))
        check(lib().wb_ppo_grad_dev(twin._h, 512, *[ptr(t) for t in dev]))
        check(lib().wb_adam_step(twin._h))
        O.ppo_train_batch(actor, critic, ohp, *batch, optimise=True)
        agent.sync()
        np.testing.assert_allclose(agent.actor.get_flat(), actor.get_params(), rtol=0, atol=2e-5)
        np.testing.assert_allclose(agent.critic.get_flat(), critic.get_params(), rtol=0, atol=2e-5)
        np.testing.assert_allclose(agent.actor.get_flat(), twin.actor.get_flat(), rtol=0, atol=2e-6)
        assert rel_err(agent.actor.get_grads(), twin.actor.get_grads()) < 1e-6
    m, v, its = agent.actor.get_adam()
    assert list(its) == [5, 5, 5]
