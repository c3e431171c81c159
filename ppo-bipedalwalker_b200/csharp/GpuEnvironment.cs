// GpuEnvironment.cs -- the reference's Environment surface (Environment.cs:64 Update, :126 StepObjects, :176 InitialState)
// forwarding to libwalker_b200 for N lockstep walkers. Source only (not compiled here); see INTEGRATION.md.
using System;
using System.Runtime.InteropServices;
using NEA.Native;
using Matrix = NEA.Walker.PPO.Matrix;

namespace NEA;

public class GpuEnvironment : IDisposable
{
    private readonly IntPtr _env;
    private readonly int _n;
    private readonly float[] _obs, _reward, _actions;
    private readonly byte[] _done;
    // the four I/O arrays stay pinned (GCHandle) and page-locked (wb_host_pin) for the life of the object: wb_env_step then
    // reads the actions and writes obs / reward / done through their device aliases (zero-copy path, include/walker_b200.h)
    private GCHandle _hObs, _hReward, _hActions, _hDone;

    public GpuEnvironment(int walkers = 1, byte floorMaterial = 6 /* Metal, Environment.cs:223 */)
    {
        _n = walkers;
        var hp = WbHyperparams.FromStatics();
        var floors = new byte[walkers];
        Array.Fill(floors, floorMaterial);
        Wb.Ok(Wb.wb_init(0), "wb_init");
        Wb.Ok(Wb.wb_env_create(walkers, floors, null, ref hp, out _env), "wb_env_create");
        _obs = new float[walkers * 12];
        _reward = new float[walkers];
        _done = new byte[walkers];
        _actions = new float[walkers * 4];
        _hObs = Pin(_obs, sizeof(float) * _obs.Length);
        _hReward = Pin(_reward, sizeof(float) * _reward.Length);
        _hActions = Pin(_actions, sizeof(float) * _actions.Length);
        _hDone = Pin(_done, _done.Length);
        InitialState();
    }

    private static GCHandle Pin(Array a, int bytes)
    {
        var h = GCHandle.Alloc(a, GCHandleType.Pinned);
        Wb.Ok(Wb.wb_host_pin(h.AddrOfPinnedObject(), (UIntPtr)bytes), "wb_host_pin");
        return h;
    }

    private static void Unpin(ref GCHandle h)
    {
        if (!h.IsAllocated) return;
        Wb.wb_host_unpin(h.AddrOfPinnedObject());
        h.Free();
    }

    // Environment.InitialState (Environment.cs:176-180)
    public void InitialState() => Wb.Ok(Wb.wb_env_get_obs(_env, _obs), "wb_env_get_obs");

    // Environment.StepObjects (Environment.cs:126-143)
    public void StepObjects(float deltaTime) => Wb.Ok(Wb.wb_env_step_objects(_env, deltaTime), "wb_env_step_objects");

    // Environment.Update (Environment.cs:64-92) for walker 0 with the action the caller sampled (PPOAgent.SampleActions).
    public Matrix Update(float deltaTime, Matrix action, out float reward, out bool terminal)
    {
        for (int k = 0; k < 4; k++) _actions[k] = action.GetValue(k, 0);   // the library applies Matrix.Clip(action, 1, -1)
        Wb.Ok(Wb.wb_env_step_pinned(_env, _hActions.AddrOfPinnedObject(), deltaTime, 1, _hObs.AddrOfPinnedObject(),
                                    _hReward.AddrOfPinnedObject(), _hDone.AddrOfPinnedObject()), "wb_env_step");
        reward = _reward[0];
        terminal = _done[0] != 0;
        var state = new float[12];
        Array.Copy(_obs, state, 12);
        return Matrix.FromValues(state);
    }

    public void Dispose()
    {
        Unpin(ref _hObs);
        Unpin(ref _hReward);
        Unpin(ref _hActions);
        Unpin(ref _hDone);
        Wb.wb_env_destroy(_env);
    }
}
