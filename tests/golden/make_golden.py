"""Regenerates tests/golden/physics_rollout.npz from the C oracle (run from the repo root).
The fixture freezes the oracle's behaviour so that later edits to oracle/ or to the CUDA path are caught; it is NOT a
reference output (the C# reference cannot run here -- parity is unpinned, see oracle/walker_oracle.h)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

MATS = ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]
n, steps = 8, 90
env = O.EnvBatch(n, floor=MATS)
rng = np.random.default_rng(20261018)
actions = rng.uniform(-1.3, 1.3, (steps, n, 4)).astype(np.float32)
states, ints, obs, rew, done = [], [], [], [], []
for t in range(steps):
    o, r, d = env.step(actions[t])
    f, iv = env.get_state()
    states.append(f); ints.append(iv); obs.append(o); rew.append(r); done.append(d)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "physics_rollout.npz"), floors=np.array(MATS), actions=actions,
                    states=np.array(states), ints=np.array(ints), obs=np.array(obs), reward=np.array(rew), done=np.array(done))
print("episodes finished:", int(np.array(done).sum()))
