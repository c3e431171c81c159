"""Small physics-only driver for profiling: N walkers, Wood floor, fixed random actions, K fused env-steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wb = ge.load_package(); wb.init(0)
env = wb.EnvBatch(n, floor_materials="Wood")
if lanes: env.set_variant(lanes)
rng = np.random.default_rng(0)
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 3
acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() for _ in range(8)]  # fresh actions every step
obs = torch.empty(n, 12, device="cuda"); rew = torch.empty(n, device="cuda"); done = torch.empty(n, dtype=torch.uint8, device="cuda")
env.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(warm): env.step_dev(acts[i % 8], obs, rew, done)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K): env.step_dev(acts[i % 8], obs, rew, done)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"n={n} lanes={lanes or "auto"} {ms:.3f} ms/step  {n/ms*1e3:.3e} env-steps/s  done_frac={done.float().mean().item():.3f}")
