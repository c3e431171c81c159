// WalkerB200Native.cs -- P/Invoke surface of libwalker_b200.so (include/walker_b200.h).
// Source only: no .NET toolchain exists in the build image, so this file is not compiled or tested here.
// Drop it into the reference project (namespace NEA) next to Environment.cs; see INTEGRATION.md.
using System;
using System.Runtime.InteropServices;
using System.Text;

namespace NEA.Native;

[StructLayout(LayoutKind.Sequential)]
public struct WbHyperparams            // wb_hyperparams <- Walker/PPO/Hyperparameters.cs:83-121
{
    public int Iterations, MaxTimesteps, BatchSize, UseGae, NormalizeAdvantages;
    public float Alpha, Beta1, Beta2, AdamEpsilon, Gamma, Lambda, Epsilon, LogStd;

    public static WbHyperparams FromStatics() => new WbHyperparams
    {
        Iterations = NEA.Walker.PPO.Hyperparameters.Iterations,
        MaxTimesteps = NEA.Walker.PPO.Hyperparameters.MaxTimesteps,
        BatchSize = NEA.Walker.PPO.Hyperparameters.BatchSize,
        UseGae = NEA.Walker.PPO.Hyperparameters.UseGAE ? 1 : 0,
        NormalizeAdvantages = NEA.Walker.PPO.Hyperparameters.NormalizeAdvantages ? 1 : 0,
        Alpha = NEA.Walker.PPO.Hyperparameters.Alpha,
        Beta1 = NEA.Walker.PPO.Hyperparameters.Beta1,
        Beta2 = NEA.Walker.PPO.Hyperparameters.Beta2,
        AdamEpsilon = NEA.Walker.PPO.Hyperparameters.AdamEpsilon,
        Gamma = NEA.Walker.PPO.Hyperparameters.Gamma,
        Lambda = NEA.Walker.PPO.Hyperparameters.Lambda,
        Epsilon = NEA.Walker.PPO.Hyperparameters.Epsilon,
        LogStd = NEA.Walker.PPO.Hyperparameters.LogStandardDeviation,
    };
}

public static class Wb
{
    private const string Lib = "walker_b200";   // libwalker_b200.so on the probing path

    [DllImport(Lib)] public static extern int wb_init(int device);
    [DllImport(Lib)] public static extern int wb_last_error(StringBuilder buf, UIntPtr len);
    [DllImport(Lib)] public static extern int wb_material_register(float inverseMass, float restitution, float friction, out int id);

    // Environment / Walker / Joint / RigidBody (Environment.cs, Walker/Walker.cs, Bodies/, Objects/)
    [DllImport(Lib)] public static extern int wb_env_create(int nEnvs, byte[] floorMaterialIds, byte[] walkerMaterialIds, ref WbHyperparams hp, out IntPtr env);
    [DllImport(Lib)] public static extern int wb_env_destroy(IntPtr env);
    [DllImport(Lib)] public static extern int wb_env_reset(IntPtr env, byte[] mask, int firstEpisode);
    [DllImport(Lib)] public static extern int wb_env_set_state(IntPtr env, float[] stateF, int[] stateI);
    [DllImport(Lib)] public static extern int wb_env_get_state(IntPtr env, float[] stateF, int[] stateI);
    [DllImport(Lib)] public static extern int wb_env_take_actions(IntPtr env, float[] actions);                  // Walker.TakeActions(Matrix.Clip(a,1,-1))
    [DllImport(Lib)] public static extern int wb_env_step_objects(IntPtr env, float deltaTime);                  // Environment.StepObjects
    [DllImport(Lib)] public static extern int wb_env_observe(IntPtr env, float[] obs, float[] reward, byte[] done);
    [DllImport(Lib)] public static extern int wb_env_get_obs(IntPtr env, float[] obs);                            // Walker.GetState
    [DllImport(Lib)] public static extern int wb_env_step(IntPtr env, float[] actions, float deltaTime, int autoReset, float[] obs, float[] reward, byte[] done);
    // pinned-pointer overload + page-locking: with GCHandle-pinned arrays registered through wb_host_pin the step is zero-copy
    [DllImport(Lib, EntryPoint = "wb_env_step")] public static extern int wb_env_step_pinned(IntPtr env, IntPtr actions, float deltaTime, int autoReset, IntPtr obs, IntPtr reward, IntPtr done);
    [DllImport(Lib)] public static extern int wb_host_pin(IntPtr hostPtr, UIntPtr bytes);
    [DllImport(Lib)] public static extern int wb_host_unpin(IntPtr hostPtr);

    // PPOAgent / NeuralNetwork / Matrix (Walker/PPO/)
    [DllImport(Lib)] public static extern int wb_policy_create(int stateSize, int actionSize, int[] actorKinds, int[] actorSizes, int actorLayers,
                                                              int[] criticKinds, int[] criticSizes, int criticLayers, ref WbHyperparams hp, out IntPtr policy);
    [DllImport(Lib)] public static extern int wb_policy_destroy(IntPtr policy);
    [DllImport(Lib)] public static extern int wb_policy_set_weights(IntPtr policy, int which, float[] flat);
    [DllImport(Lib)] public static extern int wb_policy_get_weights(IntPtr policy, int which, float[] flat);
    [DllImport(Lib)] public static extern int wb_policy_forward(IntPtr policy, int n, float[] states, float[] mean, float[] value);
    [DllImport(Lib)] public static extern int wb_policy_sample(IntPtr policy, int n, float[] states, float[] uniforms, float[] actions, float[] logp, float[] mean);
    [DllImport(Lib)] public static extern int wb_ppo_grad(IntPtr policy, int n, float[] states, float[] actions, float[] oldLogp, float[] advantages, float[] returns,
                                                         float[] losses, out int skipped);
    [DllImport(Lib)] public static extern int wb_adam_step(IntPtr policy);
    [DllImport(Lib)] public static extern int wb_returns_advantages(IntPtr policy, int n, float[] rewards, float[] values, float[] returns, float[] advantages);

    // kernel choice and the device-resident lockstep loop (device pointers: IntPtr obtained from the CUDA allocator of the host process)
    [DllImport(Lib)] public static extern int wb_env_set_variant(IntPtr env, int lanesPerWalker);   // 0 = chosen from the batch size
    [DllImport(Lib)] public static extern int wb_env_get_variant(IntPtr env, out int lanesPerWalker);
    [DllImport(Lib)] public static extern int wb_policy_set_variant(IntPtr policy, int variant);     // 0 tcgen05, 1 fp32 CUDA cores
    [DllImport(Lib)] public static extern int wb_env_step_dev(IntPtr env, IntPtr actionsDev, float deltaTime, int autoReset, IntPtr obsDev, IntPtr rewardDev, IntPtr doneDev);
    [DllImport(Lib)] public static extern int wb_policy_act_dev(IntPtr policy, int n, IntPtr statesDev, ulong seed, ulong step, IntPtr actionsDev, IntPtr logpDev,
                                                               IntPtr meanDev, IntPtr valueDev);               // SampleActions + GetValueEstimate, N walkers
    [DllImport(Lib)] public static extern int wb_segment_returns_dev(IntPtr policy, int nEnvs, int horizon, IntPtr rewardsDev, IntPtr valuesDev, IntPtr donesDev,
                                                                    IntPtr returnsDev, IntPtr advantagesDev); // MonteCarloReturn / GAE per episode fragment
    [DllImport(Lib)] public static extern int wb_gather_minibatch_dev(IntPtr policy, int batch, IntPtr indexDev, IntPtr statesPool, IntPtr actionsPool, IntPtr logpPool,
                                                                     IntPtr advantagesPool, IntPtr returnsPool, IntPtr statesOut, IntPtr actionsOut, IntPtr logpOut,
                                                                     IntPtr advantagesOut, IntPtr returnsOut);   // CreateBatches
    [DllImport(Lib)] public static extern int wb_ppo_grad_dev(IntPtr policy, int n, IntPtr statesDev, IntPtr actionsDev, IntPtr oldLogpDev, IntPtr advantagesDev, IntPtr returnsDev);
    [DllImport(Lib)] public static extern int wb_policy_grad_buffer(IntPtr policy, out IntPtr devPtr, out int nFloats);   // for the NCCL all-reduce

    /// The reference never surfaces hot-path errors: it logs and continues (RigidBody.cs:91-94). Same here.
    public static bool Ok(int status, string what)
    {
        if (status == 0) return true;
        var sb = new StringBuilder(512);
        wb_last_error(sb, (UIntPtr)512);
        NEA.Rendering.ErrorLogger.LogError($"libwalker_b200: {what} failed with status {status}: {sb}");
        return false;
    }
}
