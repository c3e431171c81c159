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
    # the same shard through the fused reduce + all-reduce over NVLink peer memory (no NCCL): must give the NCCL result
    fused = wb.PPOAgent(hp=hp, seed=9); fused.set_variant(variant)
    assert wd.connect_peers(fused)
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    dev = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (states[a:b], actions[a:b], old[a:b], adv[a:b], ret[a:b])]
    times = []
    for rep in range(6):  # several exchanges in a row exercise the epoch / parity protocol
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().wb_ppo_grad_allreduce_dev(fused._h, b - a, *[ptr(t) for t in dev]))
        e1.record(); e1.synchronize(); times.append(e0.elapsed_time(e1))
    gfu = np.concatenate([fused.actor.get_grads(), fused.critic.get_grads()])
    import ctypes as C
    cw, failed = C.c_int32(0), C.c_int32(0)
    check(lib().wb_comm_status(fused._h, C.byref(cw), C.byref(failed)))
    ga = np.concatenate([agent.actor.get_grads(), agent.critic.get_grads()])
    gf = np.concatenate([full.actor.get_grads(), full.critic.get_grads()])
    wa = np.concatenate([agent.actor.get_flat(), agent.critic.get_flat()])
    wf = np.concatenate([full.actor.get_flat(), full.critic.get_flat()])
    gerr = float(np.abs(ga - gf).max() / np.abs(gf).max()); werr = float(np.abs(wa - wf).max())
    w = torch.from_numpy(wa).cuda(); wmin, wmax = w.clone(), w.clone()
    dist.all_reduce(wmin, op=dist.ReduceOp.MIN); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    spread = float((wmax - wmin).abs().max())
    gfe = float(np.abs(gfu - ga).max() / np.abs(ga).max())
    gt = torch.from_numpy(gfu).cuda(); gmin, gmax = gt.clone(), gt.clone()
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN); dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    fspread = float((gmax - gmin).abs().max())
    # time the NCCL path the same way (grad kernel + reduce + all-reduce)
    gview = wd.grad_tensor(agent); nt = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().wb_ppo_grad_dev(agent._h, b - a, *[ptr(t) for t in dev])); wd.allreduce_sum_(gview)
        e1.record(); e1.synchronize(); nt.append(e0.elapsed_time(e1))
    if rank == 0:
        okf = gfe < 2e-6 and fspread == 0.0 and failed.value == 0 and cw.value == world
        print(f"{'PASS' if okf else 'FAIL'} world={world} variant={variant}: fused NVLink reduce+all-reduce vs NCCL path rel diff {gfe:.2e}; "
              f"spread across ranks {fspread:.1e}; grad+allreduce {1e3 * min(times):.1f} us fused vs {1e3 * min(nt):.1f} us NCCL", flush=True)
        ok = gerr < 1e-4 and werr < 1e-5 and spread == 0.0
        print(f"{'PASS' if ok else 'FAIL'} world={world} variant={variant}: all-reduced shard gradients vs single-GPU gradient rel err {gerr:.2e}; "
              f"post-Adam weight err {werr:.2e}; weight spread across ranks {spread:.1e}", flush=True)
dist.destroy_process_group()
