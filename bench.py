#!/usr/bin/env python
"""bench.py -- walker env-steps/s of the lockstep physics step (BASELINE.json configs[1]) + PPO samples/s (configs[2]).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port; rank 0 only)

Workload (config.workload): 4096 lockstep walkers per GPU, Wood floor (Carpet walker), i.i.d. U(-1,1) actions per joint per
env-step (seeded, generated once), dt = 0x1.111134p-6, Iterations = 50, auto-reset on terminal.  One "step" = one fused
env-step kernel over all walkers of the rank.  Multi-GPU: environments are block-sharded, no data-path collective
(weak scaling, 4096 walkers per GPU).

Both arms first run PREROLL (512) untimed env-steps so that the batch is in the state a long rollout is in: decorrelated (right
after construction all walkers are identical, which flatters a SIMT kernel: no divergence) and past every walker's first reset
(the reference's body list changes order exactly once, at the first Walker.Reset -- Walker.cs:212-223 -- and while a batch still
holds both orders the kernels run extra floor phases: 4096 walkers take 0.404 ms per env-step 68 steps in, 0.376 ms 512 steps
in, 0.370 ms 900 steps in; SURVEY 8d's plan for this config is likewise "timed after 100 warm-up", 1000 steps).

Timing: every timed step is bracketed by CUDA events on the launching stream; between timed steps an L2 flush (256 MiB
memset) runs OUTSIDE the event pairs; ms_per_step = sum of the K event intervals / K, max over ranks.
"e2e" = the same step through the public host-buffer call (EnvBatch.step -> wb_env_step) with pinned host actions in and host
obs/reward/done out every step, wall clock around the call (the bytes cross PCIe inside the timed region: read and written by
the kernel itself through the pinned buffers' device aliases -- the zero-copy path -- instead of staged copies).
Extra keys: "at_scale" = the same step on 262144 walkers per GPU (device-resident, the throughput regime the north-star
target is phrased in); "secondary" = PPO samples/s (configs[2]).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

N_ENVS_PER_GPU = 4096
N_ENVS_AT_SCALE = 262144   # "at_scale": the same step with the GPU full (one lane per walker)
PREROLL = 512              # untimed env-steps before the warm-up: see the module docstring
SEED = 1234
BYTES_PER_ENV_STEP = 2 * 376 + 16 + 56   # read state + actions, write state + obs/reward/done = 824 B
BYTES_PER_SUBSTEP = 2 * 376              # if the state round-tripped HBM every substep (it does not: 50 substeps are fused)
FLOP_PER_SAMPLE = 32640                  # MLP fwd + bwd
PPO_SAMPLES = 65536


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(path))[kernel_key]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_actions(n, steps, rank):
    rng = np.random.default_rng([SEED, rank])
    return rng.uniform(-1.0, 1.0, (steps, n, 4)).astype(np.float32)


def host_threads():
    """All host threads this process may use.  Passed to the oracle EXPLICITLY: torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the "all host threads" CPU arm into a single-threaded one."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(O, budget_s=12.0, max_steps=4000, nthreads=0):
    """The oracle port on the host cores: the same workload, a bounded sample (~budget_s of CPU work)."""
    nthreads = nthreads if nthreads > 0 else host_threads()
    n = N_ENVS_PER_GPU
    env = O.EnvBatch(n, floor="Wood")
    acts = make_actions(n, 8, 0)
    for k in range(PREROLL):
        env.step(acts[k % 8], nthreads=nthreads)  # pre-roll + warm-up
    t0 = time.perf_counter()
    k = 0
    while k < max_steps and (time.perf_counter() - t0) < budget_s:
        env.step(acts[k % 8], nthreads=nthreads)
        k += 1
    dt = time.perf_counter() - t0
    cores = nthreads if nthreads > 0 else (os.cpu_count() or 1)
    return {"value": n * k / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{k} env-steps x {n} walkers (Wood floor, U(-1,1) actions, auto-reset), OpenMP over envs, {dt:.1f} s"}, dt / k


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (no C#/.NET toolchain in the image, so the
    strict-fp32 C port in oracle/ stands in for it -- 'kind': 'port'), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O = ge.load_oracle()
    n = N_ENVS_PER_GPU
    env = O.EnvBatch(n, floor="Wood")
    acts = make_actions(n, PREROLL + args.warmup + args.steps, 0)
    cores = host_threads()
    for w in range(PREROLL + args.warmup):
        env.step(acts[w], nthreads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        env.step(acts[PREROLL + args.warmup + k], nthreads=cores)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "walker env-steps/sec (SAT physics step)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, cpu=True),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} env-steps x {n} walkers, OpenMP over envs on {cores} host threads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus, cpu=False):
    return {"workload": f"{N_ENVS_PER_GPU} lockstep walkers per {'host' if cpu else 'GPU'}, physics step only, Wood ground, random actions "
                        "(BASELINE.json configs[1])",
            "walkers_per_gpu": N_ENVS_PER_GPU, "iterations": 50, "dt": "0x1.111134p-6", "floor": "Wood", "walker": "Carpet",
            "actions": f"U(-1,1) numpy PCG64 seed [{SEED}, rank]", "auto_reset": True, "preroll_env_steps": PREROLL,
            "parallelism": f"env-shard x{n_gpus}, no data-path collective",
            "l2": "256 MiB memset between timed steps, outside the event pairs"}


def bench_ppo(wb, torch, stream, steps, warmup, flush):
    """configs[2]: PPO MLP fwd/bwd + clipped-surrogate gradient + Adam on a 65536-sample minibatch (synthetic obs)."""
    n = PPO_SAMPLES
    rng = np.random.default_rng(SEED + 1)
    hp = wb.default_hyperparams()
    hp.batch_size = n
    agent = wb.PPOAgent(hp=hp, seed=42, stream=stream)
    scale = np.array([1, 1, 1, 1, 1, 1, 0.1, 0.1, 0.5, 0.5, 0.5, 0.5], np.float32)
    shift = np.array([0.14, 1.6, 0.13, 1.7, 0.13, 1.7, 0, 0, 0, 0, 0, 0], np.float32)
    states = (rng.normal(size=(n, 12)).astype(np.float32) * scale + shift).astype(np.float32)
    mean, _ = agent.FeedForward(states)
    std = np.exp(np.float32(-1.0))
    actions = (mean + std * rng.normal(size=(n, 4))).astype(np.float32)
    logp = (-np.log(std) - np.log(np.sqrt(2 * np.pi)) - 0.5 * ((actions - mean) / std) ** 2).astype(np.float32)
    old_logp = (logp + 0.1 * rng.normal(size=(n, 4))).astype(np.float32)
    adv = rng.normal(size=n).astype(np.float32)
    ret = (5 * rng.normal(size=n)).astype(np.float32)
    dev = [torch.from_numpy(x).cuda() for x in (states, actions, old_logp, adv, ret)]
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    L = lib()

    def one():
        check(L.wb_ppo_train_dev(agent._h, n, *[ptr(t) for t in dev]))  # gradient kernel + fused reduce / Adam kernel

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    l0 = agent.launch_count()
    total_ms = 0.0
    for _ in range(steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    ms = total_ms / steps
    hbm, tflops, which = measured_peaks()
    achieved = n * FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12
    return {"metric": "PPO samples/sec (MLP fwd/bwd + clipped-surrogate grad + Adam)", "value": n / (ms * 1e-3), "unit": "samples/s",
            "ms_per_step": ms, "steps": steps, "dtype": "f32", "config": {"workload": "65536-sample minibatch, synthetic obs (BASELINE.json configs[2])"},
            "gpu_launches": agent.launch_count() - l0,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tflops, "unit": "TFLOP/s", "frac": achieved / tflops,
                         "traffic": ncu_traffic("ppo_tc_kernel"),
                         "note": f"32640 flop/sample; peak = {which} dense bf16 (the kernel computes in fp32 accuracy: 3xTF32 on tcgen05)"}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    wb = ge.load_package()
    wb.init(local_rank)  # raises without a CUDA device: there is no CPU fallback
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    n = N_ENVS_PER_GPU
    K, W = args.steps, args.warmup
    env = wb.EnvBatch(n, floor_materials="Wood", stream=stream)
    acts_all = make_actions(n, PREROLL + W + K, rank)
    acts_pre, acts_host = acts_all[:PREROLL], acts_all[PREROLL:]
    acts_pre_dev = torch.from_numpy(acts_pre).cuda()
    acts_dev = torch.from_numpy(acts_host).cuda()
    obs_d = torch.empty(n, 12, device="cuda")
    rew_d = torch.empty(n, device="cuda")
    done_d = torch.empty(n, dtype=torch.uint8, device="cuda")
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    for w in range(PREROLL):
        env.step_dev(acts_pre_dev[w], obs_d, rew_d, done_d)
    for w in range(W):
        env.step_dev(acts_dev[w], obs_d, rew_d, done_d)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = env.launch_count()
    total_ms = 0.0
    ndone = 0
    for k in range(K):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step_dev(acts_dev[W + k], obs_d, rew_d, done_d)
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        ndone += int(done_d.sum().item())
    barrier()
    launches = env.launch_count() - l0
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item()) / K
    kernel_ms_local = total_ms / K

    # ---- end to end through the public API: pinned host actions in, host obs/reward/done out, every step
    env2 = wb.EnvBatch(n, floor_materials="Wood", stream=stream)
    for w in range(PREROLL):
        env2.step_dev(acts_pre_dev[w], obs_d, rew_d, done_d)
    acts_pin = torch.from_numpy(acts_host).pin_memory()
    obs_h = torch.empty(n, 12).pin_memory()
    rew_h = torch.empty(n).pin_memory()
    done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = (obs_h.numpy(), rew_h.numpy(), done_h.numpy())
    for w in range(W):
        env2.step(acts_pin[w], out=out)
    barrier()
    e2e_ms = 0.0
    for k in range(K):
        flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        env2.step(acts_pin[W + k], out=out)  # H2D + kernel + D2H + stream sync inside the call
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    t2 = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t2.item()) / K
    clocks = sampler.stop() if sampler else None

    # ---- the same step with the GPU full: 262144 walkers per GPU, one lane per walker (device-resident)
    del env2
    na = N_ENVS_AT_SCALE
    env3 = wb.EnvBatch(na, floor_materials="Wood", stream=stream)
    rng = np.random.default_rng([SEED, rank, 7])
    acts3 = torch.from_numpy(rng.uniform(-1.0, 1.0, (8, na, 4)).astype(np.float32)).cuda()
    obs3 = torch.empty(na, 12, device="cuda")
    rew3 = torch.empty(na, device="cuda")
    done3 = torch.empty(na, dtype=torch.uint8, device="cuda")
    for w in range(PREROLL + 3):
        env3.step_dev(acts3[w % 8], obs3, rew3, done3)
    barrier()
    ka = max(3, min(K, 10))
    tot3 = 0.0
    for k in range(ka):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env3.step_dev(acts3[k % 8], obs3, rew3, done3)
        e1.record()
        e1.synchronize()
        tot3 += e0.elapsed_time(e1)
    barrier()
    t3 = torch.tensor([tot3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    ms3 = float(t3.item()) / ka
    hbm_peak = measured_peaks()[0]
    at_scale = {"value": world * na / (ms3 * 1e-3), "unit": "env-steps/s", "walkers_per_gpu": na, "ms_per_step": ms3, "steps": ka,
                "lanes_per_walker": env3.get_variant(),
                "roofline": {"bound": "hbm", "achieved": na * BYTES_PER_ENV_STEP / (ms3 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": na * BYTES_PER_ENV_STEP / (ms3 * 1e-3) / 1e9 / hbm_peak,
                             "substep_granular_frac": na * 50 * BYTES_PER_SUBSTEP / (ms3 * 1e-3) / 1e9 / hbm_peak,
                             "traffic": ncu_traffic("physics_lanes_kernel_262144")}}
    # ---- Iterations = 1 (SURVEY 8d asks for it): ONE substep per launch from the state of the same batch, so the 376-byte record really
    #      crosses HBM twice per substep -- the measured counterpart of the "substep-granular" accounting above
    hp1 = wb.default_hyperparams()
    hp1.iterations = 1
    env4 = wb.EnvBatch(na, floor_materials="Wood", hp=hp1, stream=stream)
    env4.set_state(*env3.get_state())
    del env3
    for w in range(3):
        env4.step_dev(acts3[w % 8], obs3, rew3, done3)
    barrier()
    tot4 = 0.0
    for k in range(ka):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env4.step_dev(acts3[k % 8], obs3, rew3, done3)
        e1.record()
        e1.synchronize()
        tot4 += e0.elapsed_time(e1)
    barrier()
    t4 = torch.tensor([tot4], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    ms4 = float(t4.item()) / ka
    at_scale["iterations_1"] = {
        "value": world * na / (ms4 * 1e-3), "unit": "substeps/s (Hyperparameters.Iterations = 1: one substep per env-step and per launch)",
        "ms_per_step": ms4, "steps": ka, "lanes_per_walker": env4.get_variant(),
        "roofline": {"bound": "hbm", "achieved": na * BYTES_PER_ENV_STEP / (ms4 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": na * BYTES_PER_ENV_STEP / (ms4 * 1e-3) / 1e9 / hbm_peak,
                     "note": "824 B per walker per launch (state in + out, actions, obs/reward/done)"}}
    del env4

    if rank == 0:
        hbm, tflops, which = measured_peaks()
        achieved = n * BYTES_PER_ENV_STEP / (kernel_ms_local * 1e-3) / 1e9
        achieved_sub = n * 50 * BYTES_PER_SUBSTEP / (kernel_ms_local * 1e-3) / 1e9
        line = {
            "metric": "walker env-steps/sec (SAT physics step)", "value": world * n / (dev_ms * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": world * n / (e2e_ms_step * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * (48 + 4 + 1), "ms_per_step": e2e_ms_step,
                    "path": "EnvBatch.step -> wb_env_step with pinned host buffers: the kernel reads the actions and writes "
                            "obs/reward/done through the buffers' device aliases (zero-copy over PCIe), one stream sync per step"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": ncu_traffic("physics_lanes_kernel_4096"),
                         "kernel": f"physics_lanes_kernel, {env.get_variant()} lanes per walker",
                         "note": f"824 B/env-step x {n} walkers per launch / CUDA-event launch time; peak = {which} copy bandwidth. "
                                 "The 50 substeps are fused on chip, so the kernel is issue/latency-bound, not HBM-bound",
                         "substep_granular": {"achieved": achieved_sub, "frac": achieved_sub / hbm,
                                              "note": "752 B/substep x 50 substeps, as if the state round-tripped HBM every substep"}},
            "clocks": clocks,
            "episodes_finished_in_timed_region": ndone,
        }
        line["at_scale"] = at_scale
        if world == 1:
            O = ge.load_oracle()
            line["cpu_baseline"], _ = cpu_baseline(O)
            if not args.no_ppo:
                line["secondary"] = bench_ppo(wb, torch, stream, max(5, K), max(3, W), flush)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-ppo", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
