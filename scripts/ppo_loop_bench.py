"""BASELINE.json configs[3]: the full PPO loop, 65536 walkers sharded over the GPUs of one box, NCCL all-reduce of the PPO
gradient only.  Launch with torchrun (one rank per GPU) or plain python for one GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
      scripts/ppo_loop_bench.py [--envs 65536] [--horizon 64] [--minibatch 65536] [--iters 3] [--floor Metal]

Prints one JSON line on rank 0: env-steps/s of the rollout phase, samples/s of the update phase, whole-loop env-steps/s
(max over ranks, CUDA events) and the cross-rank weight checksum spread (must be 0: weights stay bit-identical)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=64)
ap.add_argument("--minibatch", type=int, default=65536)
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--floor", default="Metal")
args = ap.parse_args()
rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
wb = ge.load_package()
wb.init(local_rank)
vp = wb.VectorPPO(args.envs, horizon=args.horizon, minibatch_global=args.minibatch, epochs=args.epochs, floor=args.floor, seed=11)
vp.iterate()  # warm-up iteration (also decorrelates the walkers)
tot = {"rollout_ms": 0.0, "update_ms": 0.0, "env_steps": 0, "samples_trained": 0, "minibatches": 0}
last = None
for _ in range(args.iters):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    last = vp.iterate()
    for k in tot:
        tot[k] += last[k]
t = torch.tensor([tot["rollout_ms"], tot["update_ms"]], device="cuda", dtype=torch.float64)
cs = torch.tensor([vp.weights_checksum()], device="cuda", dtype=torch.float64)
cs_min, cs_max = cs.clone(), cs.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cs_min, op=dist.ReduceOp.MIN)
    dist.all_reduce(cs_max, op=dist.ReduceOp.MAX)
if rank == 0:
    ro, up = float(t[0]) * 1e-3, float(t[1]) * 1e-3
    steps_global = args.envs * args.horizon * args.iters
    samples_global = tot["minibatches"] * args.minibatch
    print(json.dumps({
        "config": f"full PPO loop, {args.envs} walkers over {world} GPU(s), horizon {args.horizon}, minibatch {args.minibatch}, epochs {args.epochs} (BASELINE configs[3])",
        "n_gpus": world, "iters": args.iters,
        "rollout_env_steps_per_s": steps_global / ro, "update_samples_per_s": samples_global / up,
        "loop_env_steps_per_s": steps_global / (ro + up), "rollout_s": ro, "update_s": up,
        "minibatches_per_iter": tot["minibatches"] // args.iters, "allreduce_floats": 6152,
        "weights_checksum_spread": float(cs_max - cs_min), "mean_reward_last": last["mean_reward"],
        "episodes_finished_last": last["episodes_finished"], "physics_lanes": vp.env.get_variant()}))
if world > 1:
    dist.destroy_process_group()
