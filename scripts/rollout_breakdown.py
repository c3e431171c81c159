"""Where a rollout step's time goes (VectorPPO.rollout: copy obs -> states[t], wb_policy_act_dev, wb_env_step_dev), device-timed
over `T` steps each, after a 300-step warm-up.  usage: rollout_breakdown.py [walkers] [T]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
from ppo_bipedalwalker_b200.env import DT_FRAME
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
vp = wb.VectorPPO(n, horizon=T, minibatch_global=min(65536, n * T), epochs=1, floor="Wood", seed=11)
for _ in range(max(1, 300 // T)): vp.rollout()
torch.cuda.synchronize()
L = lib(); ha, he = vp.agent._h, vp.env._h


def timed(fn, reps=T):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for t in range(reps): fn(t)
    e1.record(); host = (time.perf_counter() - t0) / reps * 1e6; e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, host


def copy(t): vp.states[t].copy_(vp.obs)
def act(t):
    check(L.wb_policy_act_dev(ha, vp.n, ptr(vp.obs), 7, t, ptr(vp.actions[t]), ptr(vp.logp[t]), None, ptr(vp.values[t])))
def env(t):
    check(L.wb_env_step_dev(he, ptr(vp.actions[t]), C.c_float(DT_FRAME), 1, ptr(vp.obs), ptr(vp.rewards[t]), ptr(vp.dones[t])))
def all3(t): copy(t); act(t); env(t)


for name, fn in (("copy obs -> states[t]", copy), ("wb_policy_act_dev", act), ("wb_env_step_dev", env), ("all three", all3)):
    dev, host = timed(fn)
    print(f"{n} walkers  {name:24s} {dev:8.1f} us per step on the device, {host:6.1f} us of host time per step")
torch.cuda.synchronize(); t0 = time.perf_counter(); vp.rollout(); torch.cuda.synchronize()
print(f"{n} walkers  VectorPPO.rollout()      {(time.perf_counter() - t0) / T * 1e6:8.1f} us per step wall clock")
