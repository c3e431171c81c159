"""Quick check of one lanes-per-env variant against the oracle (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
O = ge.load_oracle()
lanes = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 40; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
print("create", flush=True)
env = wb.EnvBatch(n, floor_materials="Wood"); env.set_variant(lanes)
ref = O.EnvBatch(n, floor="Wood")
rng = np.random.default_rng(0)
print("created", flush=True)
for t in range(steps):
    a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    obs, rew, done = env.step(a); robs, rrew, rdone = ref.step(a)
    print(t, "obs equal", np.array_equal(obs.view(np.uint32), robs.view(np.uint32)), "rew", np.array_equal(rew.view(np.uint32), rrew.view(np.uint32)), flush=True)
