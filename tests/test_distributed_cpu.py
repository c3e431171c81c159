"""world_size-2 gloo test of the N>1 host logic: block sharding of a minibatch, all-reduce(sum) of the gradient buffer,
identical Adam on every rank.  The per-shard gradients come from the CPU oracle (there is no GPU here); the plumbing
under test (dist.shard_range / dist.allreduce_sum_) is exactly what the CUDA path uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    O = ge.load_oracle()
    ge.load_package()
    from ppo_bipedalwalker_b200.dist import allreduce_sum_, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank; each takes its shard
    n = 256
    actor, critic = O.Net(12, O.ACTOR_LAYERS), O.Net(12, O.CRITIC_LAYERS)
    wa = (rng.normal(size=actor.num_params) * 0.2).astype(np.float32)
    wc = (rng.normal(size=critic.num_params) * 0.2).astype(np.float32)
    actor.set_params(wa)
    critic.set_params(wc)
    hp = O.hyper_defaults()
    hp.batch_size = n  # the GLOBAL batch size divides every per-sample gradient
    states = rng.normal(size=(n, 12)).astype(np.float32)
    actions = rng.normal(size=(n, 4)).astype(np.float32) * 0.4
    old = (-0.5 * rng.random((n, 4))).astype(np.float32)
    adv = rng.normal(size=n).astype(np.float32)
    ret = rng.normal(size=n).astype(np.float32)
    a, b = shard_range(n, rank, world)
    O.ppo_train_batch(actor, critic, hp, states[a:b], actions[a:b], old[a:b], adv[a:b], ret[a:b], optimise=False)
    g = torch.from_numpy(np.concatenate([actor.get_grads(), critic.get_grads()]))
    allreduce_sum_(g)
    # identical Adam on every rank, emulated by feeding the reduced gradient back through the oracle's Adam
    full_actor, full_critic = O.Net(12, O.ACTOR_LAYERS), O.Net(12, O.CRITIC_LAYERS)
    full_actor.set_params(wa)
    full_critic.set_params(wc)
    O.ppo_train_batch(full_actor, full_critic, hp, states, actions, old, adv, ret, optimise=False)
    gfull = np.concatenate([full_actor.get_grads(), full_critic.get_grads()])
    q.put((rank, g.numpy().copy(), gfull))
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_equals_full_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda t: t[0])
    g0, g1, gfull = results[0][1], results[1][1], results[0][2]
    assert np.array_equal(g0, g1), "ranks disagree after the all-reduce (Adam would diverge)"
    assert np.abs(g0 - gfull).max() / np.abs(gfull).max() < 1e-5
