"""Synthetic workloads of BASELINE.json's configs, shared by bench.py and tests/ so that what is timed is what is tested.

Nothing here computes with the product or the oracle: the functions only draw seeded inputs (numpy) and call back into
whatever forward function the caller hands in.
"""
from __future__ import annotations

import numpy as np

MATERIALS = ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]  # Materials/*.cs

# cfg3 observation statistics: roughly the reference's scaled ranges (Walker.cs:136-149), SURVEY.md 8d
OBS_SCALE = np.array([1, 1, 1, 1, 1, 1, 0.1, 0.1, 0.5, 0.5, 0.5, 0.5], np.float32)
OBS_SHIFT = np.array([0.14, 1.6, 0.13, 1.7, 0.13, 1.7, 0, 0, 0, 0, 0, 0], np.float32)


def walker_actions(seed, rank, n, steps):
    """cfg2: i.i.d. U(-1, 1) per joint per env-step, generated once (numpy PCG64, seed [seed, rank])."""
    rng = np.random.default_rng([seed, rank])
    return rng.uniform(-1.0, 1.0, (steps, n, 4)).astype(np.float32)


def ppo_states(rng, n, spread=1.0):
    return (rng.normal(size=(n, 12)).astype(np.float32) * OBS_SCALE * np.float32(spread) + OBS_SHIFT).astype(np.float32)


def ppo_minibatch(rng, n, mean_fn, states=None, old_logp_noise=0.1, log_std=-1.0):
    """BASELINE configs[2]: a synthetic PPO minibatch of n samples.  mean_fn(states) -> actor mean [n, 4] (any implementation).
    actions = mean + sigma z, old log-probabilities = the current ones + N(0, old_logp_noise) so the ratios straddle the clip range."""
    if states is None:
        states = ppo_states(rng, n)
    mean = np.asarray(mean_fn(states), np.float32)
    std = np.exp(np.float32(log_std))
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old_logp = (logp + old_logp_noise * rng.normal(size=(n, 4))).astype(np.float32)
    adv = rng.normal(size=n).astype(np.float32)
    ret = (5 * rng.normal(size=n)).astype(np.float32)
    return states, actions, old_logp, adv, ret


def contact_stress_start(n, per_material, state_f, rng):
    """BASELINE configs[4]: n walkers, `per_material` consecutive walkers per floor material (Ice ... SuperRubber), every body
    given an initial angular velocity ~ U(-5, 5) so that floor and knee contacts pile up.  state_f: the [n, 92] record array of a
    freshly constructed batch (modified in place).  Returns the floor material names."""
    floors = [MATERIALS[(i // per_material) % len(MATERIALS)] for i in range(n)]
    state_f[:, 78:83] = rng.uniform(-5, 5, (n, 5)).astype(np.float32)
    return floors
