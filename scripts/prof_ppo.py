"""PPO minibatch timing (65536 samples): wb_ppo_train_dev on a contiguous minibatch, wb_ppo_train_indexed_dev on permuted rows of
a 4M-sample pool, and wb_gather_minibatch_dev + wb_ppo_train_dev.  usage: prof_ppo.py [n] [K] [variant] [flush]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(sys.argv[3])] if len(sys.argv) > 3 else [0, 1]
flush = len(sys.argv) > 4 and sys.argv[4] == "flush"
rng = np.random.default_rng(0)
hp = wb.default_hyperparams(); hp.batch_size = n
stream = torch.cuda.current_stream().cuda_stream
L = lib()
scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None


def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    if flush:
        for _ in range(K):
            scratch.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / K
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for variant in variants:
    agent = wb.PPOAgent(hp=hp, seed=1, stream=stream); agent.set_variant(variant)
    pool_n = 64 * n
    pool = [torch.randn(pool_n, 12, device="cuda"), 0.3 * torch.randn(pool_n, 4, device="cuda"), -0.5 * torch.rand(pool_n, 4, device="cuda"),
            torch.randn(pool_n, device="cuda"), torch.randn(pool_n, device="cuda")]
    dev = [t[:n].contiguous() for t in pool]
    mb = [torch.empty_like(t) for t in dev]
    perm = torch.randperm(pool_n, device="cuda", dtype=torch.int32)
    ms = timed(lambda: check(L.wb_ppo_train_dev(agent._h, n, *[ptr(t) for t in dev])))
    print(f"variant={variant} n={n}{' flush' if flush else ''}: {ms*1e3:.1f} us/minibatch  {n/ms*1e3:.3e} samples/s  {n*32640/ms/1e9:.1f} TFLOP/s")
    if hasattr(L, "wb_ppo_train_indexed_dev"):
        state = {"j": 0}
        def indexed():
            j = state["j"]; state["j"] = (j + 1) % 64
            check(L.wb_ppo_train_indexed_dev(agent._h, n, perm.data_ptr() + 4 * j * n, *[ptr(t) for t in pool]))
        ms = timed(indexed)
        print(f"   indexed (rows of a {pool_n}-sample pool): {ms*1e3:.1f} us/minibatch")
        def gathered():
            j = state["j"]; state["j"] = (j + 1) % 64
            check(L.wb_gather_minibatch_dev(agent._h, n, perm.data_ptr() + 4 * j * n, *[ptr(t) for t in pool], *[ptr(t) for t in mb]))
            check(L.wb_ppo_train_dev(agent._h, n, *[ptr(t) for t in mb]))
        ms = timed(gathered)
        print(f"   gather + train: {ms*1e3:.1f} us/minibatch")
