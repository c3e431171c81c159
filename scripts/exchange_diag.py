"""N-GPU diagnostic of the train-step paths (torchrun): per path, does one update move the weights, does it match the single-GPU
update, and how long does a mini-batch take in a stream of 50 back-to-back calls (no host sync in between)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as ge

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
wb = ge.load_package(); wb.init(local_rank)
from ppo_bipedalwalker_b200 import dist as wd
from ppo_bipedalwalker_b200._lib import check, lib, ptr
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
L = lib()
for n in (65536, 9472 * world):
    rng = np.random.default_rng(5)
    hp = wb.default_hyperparams(); hp.batch_size = n
    states = rng.normal(size=(n, 12)).astype(np.float32)
    ref = wb.PPOAgent(hp=hp, seed=9)
    mean, _ = ref.FeedForward(states)
    std = np.exp(np.float32(-1.0))
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old = (logp + 0.1 * rng.normal(size=(n, 4))).astype(np.float32)
    adv = rng.normal(size=n).astype(np.float32); ret = (5 * rng.normal(size=n)).astype(np.float32)
    a, b = wd.shard_range(n, rank, world)
    dev = [torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda() for x in (states, actions, old, adv, ret)]
    w0 = np.concatenate([ref.actor.get_flat(), ref.critic.get_flat()])
    ref.TrainBatch(states, actions, old, adv, ret)
    w1 = np.concatenate([ref.actor.get_flat(), ref.critic.get_flat()])
    for name, variant, nccl in (("tc one launch", 0, False), ("fp32 + exchange kernel", 1, False), ("tc + NCCL", 0, True), ("fp32 + NCCL", 1, True)):
        ag = wb.PPOAgent(hp=hp, seed=9); ag.set_variant(variant)
        gview = None
        if nccl:
            gview = wd.grad_tensor(ag)
        else:
            assert wd.connect_peers(ag)
        def step():
            if nccl:
                check(L.wb_ppo_grad_dev(ag._h, b - a, *[ptr(t) for t in dev])); wd.allreduce_sum_(gview); check(L.wb_adam_step(ag._h))
            else:
                check(L.wb_ppo_train_dev(ag._h, b - a, *[ptr(t) for t in dev]))
        step(); ag.sync()
        w = np.concatenate([ag.actor.get_flat(), ag.critic.get_flat()])
        moved, err = float(np.abs(w - w0).max()), float(np.abs(w - w1).max())
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} n={n} {name}: first update moved the weights by {moved:.2e}, differs from the single-GPU update by {err:.2e}; "
                  f"{1e3 * float(t):.1f} us per mini-batch (50 back to back)", flush=True)
dist.destroy_process_group()
