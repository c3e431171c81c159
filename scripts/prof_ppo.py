"""PPO minibatch timing (65536 samples): tensor-core vs CUDA-core kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(sys.argv[3])] if len(sys.argv) > 3 else [0, 1]
rng = np.random.default_rng(0)
hp = wb.default_hyperparams(); hp.batch_size = n
stream = torch.cuda.current_stream().cuda_stream
for variant in variants:
    agent = wb.PPOAgent(hp=hp, seed=1, stream=stream); agent.set_variant(variant)
    dev = [torch.from_numpy(x).cuda() for x in (rng.normal(size=(n, 12)).astype(np.float32), (0.3 * rng.normal(size=(n, 4))).astype(np.float32),
           (-0.5 * rng.random((n, 4))).astype(np.float32), rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32))]
    L = lib()
    def one():
        check(L.wb_ppo_train_dev(agent._h, n, *[ptr(t) for t in dev]))
    for _ in range(3): one()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"variant={variant} n={n}: {ms*1e3:.1f} us/minibatch  {n/ms*1e3:.3e} samples/s  {n*32640/ms/1e9:.1f} TFLOP/s")
