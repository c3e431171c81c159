"""N-GPU check of the data-parallel gradient path (torchrun, NCCL): every rank computes the clipped-surrogate gradient of ITS
shard of one fixed mini-batch (divided by the GLOBAL batch size), one all-reduce(sum) over NVLink, identical Adam; rank 0
compares with the whole mini-batch computed on a single GPU.  Prints PASS/FAIL lines."""
import os, sys
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
n = 65536
rng = np.random.default_rng(5)
hp = wb.default_hyperparams(); hp.batch_size = n
states = rng.normal(size=(n, 12)).astype(np.float32)
for variant in (0, 1):
    agent = wb.PPOAgent(hp=hp, seed=9); agent.set_variant(variant)
    full = wb.PPOAgent(hp=hp, seed=9); full.set_variant(variant)
    mean, _ = full.FeedForward(states)
    std = np.exp(np.float32(-1.0))
    r2 = np.random.default_rng(6)
    actions = (mean + std * r2.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old = (logp + 0.1 * r2.normal(size=(n, 4))).astype(np.float32)
    adv = r2.normal(size=n).astype(np.float32); ret = (5 * r2.normal(size=n)).astype(np.float32)
    a, b = wd.shard_range(n, rank, world)
    wd.train_minibatch_sharded(agent, states[a:b], actions[a:b], old[a:b], adv[a:b], ret[a:b])
    full.TrainBatch(states, actions, old, adv, ret)
    ga = np.concatenate([agent.actor.get_grads(), agent.critic.get_grads()])
    gf = np.concatenate([full.actor.get_grads(), full.critic.get_grads()])
    wa = np.concatenate([agent.actor.get_flat(), agent.critic.get_flat()])
    wf = np.concatenate([full.actor.get_flat(), full.critic.get_flat()])
    gerr = float(np.abs(ga - gf).max() / np.abs(gf).max()); werr = float(np.abs(wa - wf).max())
    w = torch.from_numpy(wa).cuda(); wmin, wmax = w.clone(), w.clone()
    dist.all_reduce(wmin, op=dist.ReduceOp.MIN); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    spread = float((wmax - wmin).abs().max())
    if rank == 0:
        ok = gerr < 1e-4 and werr < 1e-5 and spread == 0.0
        print(f"{'PASS' if ok else 'FAIL'} world={world} variant={variant}: all-reduced shard gradients vs single-GPU gradient rel err {gerr:.2e}; "
              f"post-Adam weight err {werr:.2e}; weight spread across ranks {spread:.1e}", flush=True)
dist.destroy_process_group()
