"""GPU parity tests proper: the CUDA physics path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: contact pairs / axis indices / contact points / state all BIT-EXACT."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MATS = ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_state_equal(env, ref, what=""):
    f, iv = env.get_state()
    rf, riv = ref.get_state()
    bad = np.argwhere(bits(f) != bits(rf))
    assert bad.size == 0, f"{what}: {len(bad)} state words differ, first (env,field)={bad[0]} gpu={f[tuple(bad[0])]!r} ref={rf[tuple(bad[0])]!r}"
    assert np.array_equal(iv, riv), f"{what}: flags/steps differ"


def test_initial_state_matches_oracle(gpu, O):
    env = gpu.EnvBatch(5)
    ref = O.EnvBatch(5)
    assert_state_equal(env, ref, "constructor")
    assert np.array_equal(bits(env.get_obs()), bits(ref.get_obs()))


def test_first_step_trace_bit_exact(gpu, O):
    n = 16
    env = gpu.EnvBatch(n, floor_materials=[m for m in MATS] * 2)
    ref = O.EnvBatch(n, floor=MATS * 2)
    rng = np.random.default_rng(0)
    for t in range(4):
        a = rng.uniform(-1.5, 1.5, (n, 4)).astype(np.float32)
        env.take_actions(a)
        ref.take_actions(a)
        pt, jt = env.debug_contacts(gpu.DT_FRAME)
        rpt, rjt = ref.step_objects(O.DT_FRAME, 50, trace=True)
        assert np.array_equal(jt.view(np.uint8), rjt.view(np.uint8)), f"joint trace differs at step {t}"
        for name in pt.dtype.names:
            assert np.array_equal(pt[name].view(np.uint32), rpt[name].view(np.uint32)), f"pair trace field {name} differs at step {t}"
        assert_state_equal(env, ref, f"after StepObjects {t}")
        obs, rew, done = env.observe()
        robs, rrew, rdone = ref.observe()
        assert np.array_equal(bits(obs), bits(robs)) and np.array_equal(bits(rew), bits(rrew)) and np.array_equal(done, rdone)


@pytest.mark.parametrize("n,steps", [(64, 120), (1000, 40)])
def test_rollout_with_resets_bit_exact(gpu, O, n, steps):
    """Long enough that walkers fall, terminate and are auto-reset into the floor-first list order."""
    floors = [MATS[i % 8] for i in range(n)]
    env = gpu.EnvBatch(n, floor_materials=floors)
    ref = O.EnvBatch(n, floor=floors)
    rng = np.random.default_rng(n)
    ndone = 0
    for t in range(steps):
        a = rng.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
        obs, rew, done = env.step(a)
        robs, rrew, rdone = ref.step(a)
        assert np.array_equal(done, rdone), f"done differs at step {t}"
        assert np.array_equal(bits(obs), bits(robs)), f"obs differs at step {t}"
        assert np.array_equal(bits(rew), bits(rrew)), f"reward differs at step {t}"
        ndone += int(done.sum())
    assert_state_equal(env, ref, "end of rollout")
    if steps >= 100:
        assert ndone > 0, "rollout never exercised the reset path"


def test_set_state_roundtrip_and_identical_start(gpu, O):
    n = 24
    ref = O.EnvBatch(n, floor="Wood")
    rng = np.random.default_rng(3)
    for _ in range(30):
        ref.step(rng.uniform(-1, 1, (n, 4)).astype(np.float32))
    f, iv = ref.get_state()
    env = gpu.EnvBatch(n, floor_materials="Wood")
    env.set_state(f, iv)
    f2, iv2 = env.get_state()
    assert np.array_equal(bits(f), bits(f2)) and np.array_equal(iv, iv2)
    a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    obs, rew, done = env.step(a)
    robs, rrew, rdone = ref.step(a)
    assert np.array_equal(bits(obs), bits(robs)) and np.array_equal(bits(rew), bits(rrew)) and np.array_equal(done, rdone)
    assert_state_equal(env, ref, "one step from an injected state")


def test_contact_stress_all_materials(gpu, O):
    """cfg5-style: walkers dropped with spin on every floor material; contact pairs/axes bit-exact."""
    n = 64
    floors = [MATS[i % 8] for i in range(n)]
    ref = O.EnvBatch(n, floor=floors)
    env = gpu.EnvBatch(n, floor_materials=floors)
    f, iv = ref.get_state()
    rng = np.random.default_rng(11)
    f[:, 78:83] = rng.uniform(-5, 5, (n, 5)).astype(np.float32)  # initial angular velocities
    ref.set_state(f, iv)
    env.set_state(f, iv)
    total_contacts = 0
    for t in range(12):
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        env.take_actions(a)
        ref.take_actions(a)
        pt, jt = env.debug_contacts(gpu.DT_FRAME)
        rpt, rjt = ref.step_objects(O.DT_FRAME, 50, trace=True)
        for name in pt.dtype.names:
            assert np.array_equal(pt[name].view(np.uint32), rpt[name].view(np.uint32)), f"{name} differs at step {t}"
        total_contacts += int(rpt["ncontacts"].sum())
        env.observe()
        ref.observe()
    assert total_contacts > 1000
    assert_state_equal(env, ref, "stress")


def test_contact_stress_sweep_at_baseline_size(gpu, O):
    """BASELINE configs[4] at its full size: 16 384 walkers, 2 048 per floor material (Ice ... SuperRubber), every body started
    with a spin ~ U(-5, 5), random actions, 24 env-steps of 50 substeps (RigidBody.cs:66-96 under maximum contact load).
      * all 16 384 records (state, flags, step counters), observations, rewards and done flags bit-exact against the oracle
        after every env-step's observe, with the kernel wb_env_create picks for this size;
      * per-substep pair and joint traces (candidate, AABB, SAT result, axis index, normal, depth, contact points) bit-exact on a
        512-walker sample (the first 64 walkers of every material, same start state and actions as in the big batch), and the
        sample's final records equal the corresponding rows of the big batch."""
    import workloads
    n, per, steps = 16384, 2048, 24
    rng = np.random.default_rng(5)
    f0, iv0 = O.EnvBatch(n).get_state()
    floors = workloads.contact_stress_start(n, per, f0, rng)
    ref = O.EnvBatch(n, floor=floors)
    env = gpu.EnvBatch(n, floor_materials=floors)
    ref.set_state(f0, iv0)
    env.set_state(f0, iv0)
    sample = np.concatenate([np.arange(k * per, k * per + 64) for k in range(8)])
    sfloors = [floors[i] for i in sample]
    sref = O.EnvBatch(len(sample), floor=sfloors)
    senv = gpu.EnvBatch(len(sample), floor_materials=sfloors)
    sref.set_state(f0[sample].copy(), iv0[sample].copy())
    senv.set_state(f0[sample].copy(), iv0[sample].copy())
    contacts = sat = 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        obs, rew, done = env.step(a, auto_reset=False)
        robs, rrew, rdone = ref.step(a, auto_reset=False)
        assert np.array_equal(obs.view(np.uint32), robs.view(np.uint32)), f"observations differ at env-step {t}"
        assert np.array_equal(rew.view(np.uint32), rrew.view(np.uint32)) and np.array_equal(done, rdone), f"reward/done differ at env-step {t}"
        senv.take_actions(a[sample])
        sref.take_actions(a[sample])
        pt, jt = senv.debug_contacts(gpu.DT_FRAME)
        rpt, rjt = sref.step_objects(O.DT_FRAME, 50, trace=True)
        for name in pt.dtype.names:
            assert np.array_equal(pt[name].view(np.uint32), rpt[name].view(np.uint32)), f"pair trace field {name} differs at env-step {t}"
        for name in jt.dtype.names:
            assert np.array_equal(jt[name].view(np.uint32), rjt[name].view(np.uint32)), f"joint trace field {name} differs at env-step {t}"
        contacts += int(rpt["ncontacts"].sum())
        sat += int(rpt["sat"].sum())
        senv.observe()
        sref.observe()
    assert_state_equal(env, ref, "contact stress, 16384 walkers")
    assert_state_equal(senv, sref, "contact stress, traced sample")
    gf, giv = env.get_state()
    sf, siv = senv.get_state()
    # (flags only: Environment._steps is advanced by the fused step, not by the granular TakeActions / StepObjects / observe calls)
    assert np.array_equal(gf[sample].view(np.uint32), sf.view(np.uint32)) and np.array_equal(giv[sample][:, 0], siv[:, 0])
    # the sweep really is contact-heavy: ~55 SAT collisions and ~65 contact points per walker per env-step
    assert sat > 30 * len(sample) * steps and contacts > 30 * len(sample) * steps


def test_user_material_plugin(gpu, O):
    m = gpu.IMaterial(7.5, 0.55, 0.33).register()
    assert m.id >= 8
    env = gpu.EnvBatch(4, floor_materials=m)
    ref = O.EnvBatch(4, floor=(7.5, 0.55, 0.33))
    rng = np.random.default_rng(5)
    for _ in range(25):
        a = rng.uniform(-1, 1, (4, 4)).astype(np.float32)
        env.step(a)
        ref.step(a)
    assert_state_equal(env, ref, "custom material")


@pytest.mark.parametrize("variant", [0, 1, 4, 1001, 1002])
def test_ragged_batch_sizes(gpu, O, variant):
    """Batch sizes that leave partially filled warps / CTAs / SoA padding, for the auto-selected kernel and explicit variants."""
    for n in (1, 7, 9, 33, 300):
        env = gpu.EnvBatch(n)
        env.set_variant(variant)
        ref = O.EnvBatch(n)
        a = np.linspace(-1, 1, n * 4, dtype=np.float32).reshape(n, 4)
        for _ in range(3):
            env.step(a)
            ref.step(a)
        assert_state_equal(env, ref, f"n={n} variant={variant}")


def test_nonfinite_and_out_of_range_actions_are_clipped_like_matrix_clip(gpu, O):
    n = 8
    env = gpu.EnvBatch(n)
    ref = O.EnvBatch(n)
    a = np.array([[5, -5, 1, -1], [0.999, -0.999, 1e-30, -0.0]] * 4, np.float32)
    env.step(a)
    ref.step(a)
    assert_state_equal(env, ref, "clip")


def test_pinned_host_buffers_zero_copy_path_matches_oracle(gpu, O):
    """wb_env_step with PINNED host buffers: the kernel reads the actions and writes obs / reward / done straight through the
    device aliases of the caller's buffers (no staging copies).  Same bits as the oracle, and as the staged path."""
    import torch
    n = 300
    env = gpu.EnvBatch(n, floor_materials=[MATS[i % 8] for i in range(n)])
    staged = gpu.EnvBatch(n, floor_materials=[MATS[i % 8] for i in range(n)])
    ref = O.EnvBatch(n, floor=[MATS[i % 8] for i in range(n)])
    a_pin = torch.empty(n, 4).pin_memory()
    obs, rew, done = torch.empty(n, 12).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    out = (obs.numpy(), rew.numpy(), done.numpy())
    rng = np.random.default_rng(11)
    for t in range(30):
        a = rng.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
        a_pin.copy_(torch.from_numpy(a))
        env.step(a_pin, out=out)
        sobs, srew, sdone = staged.step(a)  # pageable numpy buffers: staged copies
        robs, rrew, rdone = ref.step(a)
        assert np.array_equal(bits(out[0]), bits(robs)) and np.array_equal(bits(out[1]), bits(rrew)) and np.array_equal(out[2], rdone), f"step {t}"
        assert np.array_equal(bits(sobs), bits(robs)) and np.array_equal(bits(srew), bits(rrew)) and np.array_equal(sdone, rdone)
    assert_state_equal(env, ref, "zero-copy path")


def test_host_pin_makes_numpy_buffers_zero_copy(gpu, O):
    """wb_host_pin / wb_host_unpin (what the C# shim does with its GCHandle-pinned arrays): registered numpy buffers take the
    zero-copy path of wb_env_step and give the oracle's bits -- page-aligned buffers, and the realistic case of four small
    arrays that SHARE one page (cudaHostRegister refuses overlapping pages: wb_host_pin registers what is still missing)."""
    for shared_page in (False, True):
        n = 3 if shared_page else 64
        env = gpu.EnvBatch(n)
        ref = O.EnvBatch(n)
        if shared_page:
            raw = np.zeros(8192, np.uint8)
            off = (-raw.ctypes.data) % 4096 + 64      # all four arrays inside one page, 16-byte aligned
            def carve(nbytes, dtype, shape):
                nonlocal off
                v = raw[off:off + nbytes].view(dtype).reshape(shape)
                off += (nbytes + 15) // 16 * 16
                return v
            a, obs = carve(n * 16, np.float32, (n, 4)), carve(n * 48, np.float32, (n, 12))
            rew, done = carve(n * 4, np.float32, (n,)), carve(n, np.uint8, (n,))
            keep = [raw]
        else:
            def aligned(shape, dtype):
                nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
                r = np.empty(nbytes + 8192, np.uint8)
                o = (-r.ctypes.data) % 4096
                return r[o:o + nbytes].view(dtype).reshape(shape), r
            (a, k0), (obs, k1), (rew, k2), (done, k3) = (aligned((n, 4), np.float32), aligned((n, 12), np.float32),
                                                          aligned((n,), np.float32), aligned((n,), np.uint8))
            keep = [k0, k1, k2, k3]
        bufs = [a, obs, rew, done]
        for b in bufs:
            gpu.pin_host(b)
        with pytest.raises(gpu.WalkerB200Error):
            gpu.pin_host(a)                           # the same buffer twice is an error, not a silent double registration
        try:
            rng = np.random.default_rng(3)
            for t in range(10):
                a[:] = rng.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
                env.step(a, out=(obs, rew, done))
                robs, rrew, rdone = ref.step(a.copy())
                assert np.array_equal(bits(obs), bits(robs)) and np.array_equal(bits(rew), bits(rrew)) and np.array_equal(done, rdone), f"step {t}"
        finally:
            for b in bufs:
                gpu.unpin_host(b)
        with pytest.raises(gpu.WalkerB200Error):
            gpu.unpin_host(a)
        assert_state_equal(env, ref, f"host_pin path (shared_page={shared_page})")
        del keep


def test_plain_c_program_drives_the_abi(gpu, tmp_path):
    """examples/c_abi_smoke.c (C99, no CUDA headers): wb_env_step with malloc'ed buffers and with wb_host_pin'ed buffers agree."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_abi_cpu import _build_c_example
    exe = _build_c_example(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "staged and zero-copy paths agree" in res.stdout


def test_committed_golden_rollout(gpu):
    """The CUDA path against the committed fixture (tests/golden/physics_rollout.npz): no oracle at run time."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics_rollout.npz"))
    env = gpu.EnvBatch(8, floor_materials=[str(x) for x in g["floors"]])
    for t in range(g["actions"].shape[0]):
        obs, rew, done = env.step(g["actions"][t])
        f, iv = env.get_state()
        assert np.array_equal(bits(f), bits(g["states"][t])), f"state differs from the golden fixture at step {t}"
        assert np.array_equal(iv, g["ints"][t]) and np.array_equal(bits(obs), bits(g["obs"][t]))
        assert np.array_equal(bits(rew), bits(g["reward"][t])) and np.array_equal(done, g["done"][t])


def test_full_size_properties_4096_walkers(gpu):
    """BASELINE configs[1] size: size-independent properties -- identical envs stay identical in lockstep, walkers never
    tunnel through the floor, every record stays finite, episodes terminate and reset."""
    n = 4096
    env = gpu.EnvBatch(n, floor_materials="Wood")
    rng = np.random.default_rng(0)
    base = rng.uniform(-1, 1, (60, 8, 4)).astype(np.float32)
    total_done = 0
    for t in range(60):
        a = np.tile(base[t], (n // 8, 1))  # env i and env i+8 receive identical actions
        obs, rew, done = env.step(a)
        total_done += int(done.sum())
        assert np.array_equal(obs[:8].view(np.uint32), obs[8:16].view(np.uint32))
        assert np.array_equal(obs.reshape(n // 8, 8, 12)[0].view(np.uint32), obs.reshape(n // 8, 8, 12)[-1].view(np.uint32))
    f, iv = env.get_state()
    assert np.isfinite(f).all() and f[:, 1:58:2].max() < 915
    assert total_done > 0 and (iv[:, 1] <= 60).all()


@pytest.mark.parametrize("variant,n,steps", [(0, 4096, 150), (1001, 4096, 150), (1, 1024, 400)])
def test_long_rollout_at_baseline_size_bit_exact(gpu, O, variant, n, steps):
    """BASELINE configs[1] size against the oracle itself (OpenMP over walkers): every observation, reward and done flag of a
    long random-action rollout with many episode ends, then the complete state.  0.6 M env-steps = 30 M substeps per case."""
    env = gpu.EnvBatch(n, floor_materials="Wood")
    env.set_variant(variant)
    ref = O.EnvBatch(n, floor="Wood")
    rng = np.random.default_rng(variant + n)
    ndone = 0
    for t in range(steps):
        a = rng.uniform(-1.1, 1.1, (n, 4)).astype(np.float32)
        obs, rew, done = env.step(a)
        robs, rrew, rdone = ref.step(a)
        assert np.array_equal(bits(obs), bits(robs)), f"obs differ at step {t}"
        assert np.array_equal(bits(rew), bits(rrew)) and np.array_equal(done, rdone), f"reward/done differ at step {t}"
        ndone += int(done.sum())
    assert ndone > n // 4  # plenty of resets (both list orders are exercised)
    assert_state_equal(env, ref, f"variant={variant}")


def test_kernels_agree_at_full_scale_through_many_resets(gpu):
    """BASELINE configs[3] size (65536 walkers, device-resident, 120 env-steps from the spawn state, so every walker goes
    through its first reset and both list orders occur): the compacting throughput kernel, the one-lane kernel and the 4-lane
    kernel must leave bit-identical state, observations, rewards and done flags -- three thread mappings of the same arithmetic."""
    import torch
    n, steps = 65536, 120
    rng = np.random.default_rng(99)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() for _ in range(8)]
    results = []
    for variant in (1001, 1, 4):
        env = gpu.EnvBatch(n, floor_materials=[MATS[i % 8] for i in range(n)])
        env.set_variant(variant)
        obs = torch.empty(n, 12, device="cuda")
        rew = torch.empty(n, device="cuda")
        done = torch.empty(n, dtype=torch.uint8, device="cuda")
        rew_sum = torch.zeros(n, device="cuda", dtype=torch.float64)
        ndone = torch.zeros((), device="cuda", dtype=torch.int64)
        env.set_stream(torch.cuda.current_stream().cuda_stream)
        for t in range(steps):
            env.step_dev(acts[t % 8], obs, rew, done)
            rew_sum += rew.double()
            ndone += done.sum()
        torch.cuda.synchronize()
        f, iv = env.get_state()
        results.append((f, iv, obs.cpu().numpy(), rew_sum.cpu().numpy(), int(ndone.item())))
        del env
    assert results[0][4] > n // 8            # plenty of episode ends
    for other in results[1:]:
        assert np.array_equal(bits(results[0][0]), bits(other[0])) and np.array_equal(results[0][1], other[1])
        assert np.array_equal(bits(results[0][2]), bits(other[2]))
        assert np.array_equal(results[0][3], other[3]) and results[0][4] == other[4]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_rotation_coefficients_match_libm_rounded_to_float(gpu, O, mode):
    """(float)cos((double)theta), (float)sin((double)theta): the polynomial fast path, the forced double-double path and
    the device sincos path against the host libm (what .NET's Math.Cos/Sin call) on 2M angles incl. per-substep-sized ones."""
    import ctypes as C
    from ppo_bipedalwalker_b200._lib import check, lib, ptr
    rng = np.random.default_rng(mode)
    small = (rng.uniform(-1, 1, 1_500_000) * 10.0 ** rng.uniform(-7, -1.6, 1_500_000)).astype(np.float32)
    edge = np.array([0.0, -0.0, 1e-38, -1e-38, 1e-45, 0.03124999, -0.03124999, 3.3333397e-4, 1e-20], np.float32)
    big = rng.uniform(-0.03125, 0.03125, 500_000).astype(np.float32) if mode == 1 else rng.uniform(-7, 7, 500_000).astype(np.float32)
    ang = np.concatenate([small, edge, big])
    if mode == 1:
        ang = ang[np.abs(ang) < 0.03125]
    c = np.empty_like(ang)
    s = np.empty_like(ang)
    check(lib().wb_debug_rotz(ang.size, ptr(ang), mode, ptr(c), ptr(s)))
    rc = np.cos(ang.astype(np.float64)).astype(np.float32)
    rs = np.sin(ang.astype(np.float64)).astype(np.float32)
    bad_c = int((c.view(np.uint32) != rc.view(np.uint32)).sum())
    bad_s = int(((s.view(np.uint32) != rs.view(np.uint32)) & ~((s == 0) & (rs == 0))).sum())
    if mode == 2:
        assert bad_c + bad_s <= 2   # two independent <=1-2 ulp libms: agreement after float rounding is statistical
    else:
        assert bad_c == 0 and bad_s == 0


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32, 104, 108, 116, 1001, 1002, 1003])
def test_kernel_variants_bit_exact(gpu, O, lanes):
    """Both thread mappings (two envs per warp / one env per warp) against the oracle, traces included."""
    n = 40
    floors = [MATS[i % 8] for i in range(n)]
    env = gpu.EnvBatch(n, floor_materials=floors)
    env.set_variant(lanes)
    ref = O.EnvBatch(n, floor=floors)
    rng = np.random.default_rng(lanes)
    for t in range(70):
        a = rng.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
        if t % 10 == 0:
            env.take_actions(a)
            ref.take_actions(a)
            pt, jt = env.debug_contacts(gpu.DT_FRAME)
            rpt, rjt = ref.step_objects(O.DT_FRAME, 50, trace=True)
            assert np.array_equal(jt.view(np.uint8), rjt.view(np.uint8))
            for name in pt.dtype.names:
                assert np.array_equal(pt[name].view(np.uint32), rpt[name].view(np.uint32)), f"{name} differs at step {t}"
            env.observe()
            ref.observe()
        else:
            obs, rew, done = env.step(a)
            robs, rrew, rdone = ref.step(a)
            assert np.array_equal(bits(obs), bits(robs)) and np.array_equal(bits(rew), bits(rrew)) and np.array_equal(done, rdone)
    assert_state_equal(env, ref, f"lanes={lanes}")


def test_branch_free_normalize_factor_is_exhaustively_exact(gpu):
    """rcp_sqrt_rn(s) (one MUFU.RSQ + 4 FMAs, what the thread-per-env kernel uses for Vector2.Normalize's 1f / sqrt(s)) equals
    __frcp_rn(__fsqrt_rn(s)) for EVERY one of the 2^32 float bit patterns."""
    import ctypes as C
    from ppo_bipedalwalker_b200._lib import check, lib
    bad = C.c_uint64(0)
    first = C.c_uint32(0)
    check(lib().wb_debug_rcp_sqrt_check(0, 1 << 32, C.byref(bad), C.byref(first)))
    assert bad.value == 0, f"{bad.value} mismatches, first at bits 0x{first.value:08x}"
