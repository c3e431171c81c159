"""Kernel-variant sweep from ONE rollout state: pre-roll `warm` env-steps with the default variant, snapshot the state, then
for every variant restore the snapshot, time K fused env-steps (CUDA events on the launching stream) and hash the final
state: all variants must end bit-identical (printed as `same=True`).

usage: [WB_ITERATIONS=k] python scripts/sweep_physics.py N K WARM variant [variant ...]
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

n = int(sys.argv[1])
K = int(sys.argv[2])
warm = int(sys.argv[3])
variants = [int(v) for v in sys.argv[4:]]
wb = ge.load_package()
wb.init(0)
hp = wb.default_hyperparams()
hp.iterations = int(os.environ.get("WB_ITERATIONS", hp.iterations))  # Hyperparameters.Iterations (substeps per env-step)
env = wb.EnvBatch(n, floor_materials="Wood", hp=hp)
rng = np.random.default_rng(0)
acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() for _ in range(8)]
obs = torch.empty(n, 12, device="cuda")
rew = torch.empty(n, device="cuda")
done = torch.empty(n, dtype=torch.uint8, device="cuda")
env.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(warm):
    env.step_dev(acts[i % 8], obs, rew, done)
torch.cuda.synchronize()
f0, iv0 = env.get_state()
ref_hash = None
for v in variants:
    env.set_state(f0, iv0)
    env.set_variant(v)
    for i in range(2):  # untimed: code / constant caches
        env.step_dev(acts[i % 8], obs, rew, done)
    env.set_state(f0, iv0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        env.step_dev(acts[i % 8], obs, rew, done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    f, iv = env.get_state()
    h = hashlib.sha1(f.tobytes() + iv.tobytes() + obs.cpu().numpy().tobytes()).hexdigest()[:12]
    if ref_hash is None:
        ref_hash = h
    print(f"n={n} variant={v} {ms:.4f} ms/step  {n / ms * 1e3:.4e} env-steps/s  hash={h} same={h == ref_hash}", flush=True)
