// GpuPPOAgent.cs -- the reference's PPOAgent surface (Walker/PPO/PPOAgent.cs:23 ctor, :147 Train(Trajectory, Renderer),
// :192 Save, :381 SampleActions) forwarding to libwalker_b200's wb_policy_* exports.  Source only (no .NET toolchain in the
// build image); see INTEGRATION.md.  Host-side pieces the reference keeps in C# stay in C# here: the network DSL
// (ParseLayers), Xavier initialisation (Matrix.FromXavier), the System.Random draws of the Box-Muller transform, the
// shuffling of CreateBatches and the .weights text format.  Forward passes, the clipped-surrogate gradient, the backward
// pass, Adam and the return / advantage recurrences run in the library.
using System;
using System.Collections.Generic;
using System.IO;
using System.Linq;
using System.Text.RegularExpressions;
using NEA.Native;
using NEA.Rendering;
using NEA.Walker.PPO.Network;

namespace NEA.Walker.PPO;

public sealed class GpuPPOAgent : IDisposable
{
    private readonly IntPtr _policy;
    private readonly int _stateSize, _actionSize;
    private readonly List<(int outSize, int inSize)> _actorDense = new(), _criticDense = new();
    private readonly Random _random = new Random();
    private const string WeightsLocation = "Data/Weights/";     // PPOAgent.cs:21

    // new PPOAgent(stateSize, actionSize) (PPOAgent.cs:23-37): parse both DSL strings (falling back to the defaults like
    // CreateNetworks, :41-93), create the device networks, Xavier-initialise, then Load("critic") / Load("actor").
    public GpuPPOAgent(int stateSize, int actionSize)
    {
        _stateSize = stateSize;
        _actionSize = actionSize;
        const string defaultCritic = "Input |64| (LeakyReLU) |1| Output", defaultActor = "Input |64| (LeakyReLU) |64| (LeakyReLU) |4| (TanH) Output";
        if (!TryParse(Hyperparameters.CriticNeuralNetwork, out var ck, out var cs) || cs.Last(s => s > 0) != 1)
        {
            ErrorLogger.LogError("Exception occurred while attempting to parse the critic neural network.");
            Hyperparameters.CriticNeuralNetwork = defaultCritic;
            TryParse(defaultCritic, out ck, out cs);
        }
        if (!TryParse(Hyperparameters.ActorNeuralNetwork, out var ak, out var asz) || asz.Last(s => s > 0) != actionSize)
        {
            ErrorLogger.LogError("Exception occurred while attempting to parse the actor neural network.");
            Hyperparameters.ActorNeuralNetwork = defaultActor;
            TryParse(defaultActor, out ak, out asz);
        }
        var hp = WbHyperparams.FromStatics();
        Wb.Ok(Wb.wb_init(0), "wb_init");
        Wb.Ok(Wb.wb_policy_create(stateSize, actionSize, ak, asz, ak.Length, ck, cs, ck.Length, ref hp, out _policy), "wb_policy_create");
        DenseShapes(stateSize, ak, asz, _actorDense);
        DenseShapes(stateSize, ck, cs, _criticDense);
        Wb.Ok(Wb.wb_policy_set_weights(_policy, 1, Xavier(_criticDense)), "wb_policy_set_weights(critic)");
        Wb.Ok(Wb.wb_policy_set_weights(_policy, 0, Xavier(_actorDense)), "wb_policy_set_weights(actor)");
        Load("critic");
        Load("actor");
    }

    internal IntPtr Handle => _policy;

    // PPOAgent.ParseLayers (PPOAgent.cs:96-143) -> kinds / sizes for wb_policy_create
    private static bool TryParse(string structure, out int[] kinds, out int[] sizes)
    {
        kinds = sizes = Array.Empty<int>();
        if (!Regex.IsMatch(structure, @"^Input( \|\d+\|| \((ReLU|TanH|LeakyReLU)\))+ Output$")) return false;
        string cleaned = Regex.Replace(structure, @"[\[|()]|( Output)|(Input )", "");
        var k = new List<int>();
        var s = new List<int>();
        foreach (var token in cleaned.Split(' '))
        {
            if (int.TryParse(token, out int size)) { k.Add(Wb.Dense); s.Add(size); }
            else { k.Add(token == "ReLU" ? Wb.ReLU : token == "LeakyReLU" ? Wb.LeakyReLU : Wb.TanH); s.Add(0); }
        }
        kinds = k.ToArray();
        sizes = s.ToArray();
        return k.Contains(Wb.Dense);
    }

    private static void DenseShapes(int input, int[] kinds, int[] sizes, List<(int, int)> shapes)
    {
        int width = input;
        for (int i = 0; i < kinds.Length; i++)
            if (kinds[i] == Wb.Dense) { shapes.Add((sizes[i], width)); width = sizes[i]; }
    }

    // DenseLayer ctor (DenseLayer.cs:22-33): weights Matrix.FromXavier(out, in) (Matrix.cs:59-80), zero biases;
    // flat layout per dense layer W[out][in] row-major then b[out] (DenseLayer.Save order)
    private static float[] Xavier(List<(int outSize, int inSize)> shapes)
    {
        var flat = new List<float>();
        foreach (var (o, i) in shapes)
        {
            flat.AddRange(Matrix.GetRepresentation(Matrix.FromXavier(o, i)));
            flat.AddRange(new float[o]);
        }
        return flat.ToArray();
    }

    // PPOAgent.SampleActions (PPOAgent.cs:381-398) for one state: the mean comes from the device actor, the two uniforms of
    // every Box-Muller draw from System.Random like Matrix.SampleNormal (Matrix.cs:541-555, NormalDistribution.cs:12-20)
    public Matrix SampleActions(Matrix state, out Matrix logProbabilities, out Matrix mean, out Matrix std)
    {
        var actions = new float[_actionSize];
        var logp = new float[_actionSize];
        var mu = new float[_actionSize];
        SampleActions(1, Matrix.GetRepresentation(state), actions, logp, mu);
        logProbabilities = Matrix.FromValues(logp);
        mean = Matrix.FromValues(mu);
        std = Matrix.FromValues(Enumerable.Repeat(MathF.Exp(Hyperparameters.LogStandardDeviation), _actionSize).ToArray());   // GetStandardDeviations, :367-378
        return Matrix.FromValues(actions);
    }

    // the same for n walkers at once: states [n][stateSize] -> actions / logp / mean [n][actionSize]
    public void SampleActions(int n, float[] states, float[] actions, float[] logp, float[] mean)
    {
        var uniforms = new float[n * _actionSize * 2];
        for (int i = 0; i < uniforms.Length; i++) uniforms[i] = (float)_random.NextDouble();
        Wb.Ok(Wb.wb_policy_sample(_policy, n, states, uniforms, actions, logp, mean), "wb_policy_sample");
    }

    // PPOAgent.GetValueEstimate (PPOAgent.cs:350-364) for n states
    public float[] ValueEstimates(int n, float[] states)
    {
        var values = new float[n];
        Wb.Ok(Wb.wb_policy_forward(_policy, n, states, null, values), "wb_policy_forward");
        return values;
    }

    // PPOAgent.Train(Trajectory, Renderer) (PPOAgent.cs:147-172)
    public void Train(Trajectory trajectory, Renderer renderer)
    {
        int T = trajectory.States.Count;
        if (T == 0) return;
        var hp = WbHyperparams.FromStatics();                      // the static hyper-parameters may have been edited in the menu
        Wb.Ok(Wb.wb_policy_set_hyperparams(_policy, ref hp), "wb_policy_set_hyperparams");
        float[] states = Flatten(trajectory.States, _stateSize), actions = Flatten(trajectory.Actions, _actionSize);
        float[] logp = Flatten(trajectory.LogProbabilities, _actionSize), rewards = trajectory.Rewards.ToArray();
        // CalculateValues (:175-189): value estimates, MC return / GAE, optional Normalize
        float[] values = ValueEstimates(T, states), returns = new float[T], advantages = new float[T];
        Wb.Ok(Wb.wb_returns_advantages(_policy, T, rewards, values, returns, advantages), "wb_returns_advantages");
        trajectory.Values.Clear(); trajectory.Values.AddRange(values);
        trajectory.Returns.Clear(); trajectory.Returns.AddRange(returns);
        trajectory.Advantages.Clear(); trajectory.Advantages.AddRange(advantages);
        renderer.AddTotalEpisodeReward(trajectory.Rewards.Sum());

        int B = Hyperparameters.BatchSize, batchCount = T / B;
        var losses = new float[2];
        float[] bs = new float[B * _stateSize], ba = new float[B * _actionSize], bl = new float[B * _actionSize], badv = new float[B], bret = new float[B];
        for (int epoch = 0; epoch < Hyperparameters.Epochs; epoch++)
        {
            // CreateBatches (:501-540): sampling without replacement, the remainder is dropped
            var pool = Enumerable.Range(0, T).ToList();
            for (int j = 0; j < batchCount; j++)
            {
                for (int b = 0; b < B; b++)
                {
                    int pick = _random.Next(0, pool.Count), idx = pool[pick];
                    pool.RemoveAt(pick);
                    Array.Copy(states, idx * _stateSize, bs, b * _stateSize, _stateSize);
                    Array.Copy(actions, idx * _actionSize, ba, b * _actionSize, _actionSize);
                    Array.Copy(logp, idx * _actionSize, bl, b * _actionSize, _actionSize);
                    badv[b] = advantages[idx];
                    bret[b] = returns[idx];
                }
                // Train(Batch) (:218-346): Zero, per-sample clipped-surrogate + value gradients, FeedBack, Optimise -- one call,
                // one kernel launch on the default networks (wb_ppo_grad + wb_adam_step are the two halves)
                Wb.Ok(Wb.wb_ppo_train(_policy, B, bs, ba, bl, badv, bret, losses, out _), "wb_ppo_train");
                renderer.UpdateConsole(epoch, j, batchCount, losses[0]);
            }
        }
        renderer.AddCriticLoss(losses[0]);
        renderer.AddActorLoss(losses[1]);
        if (Hyperparameters.SaveWeights) Save();
    }

    private static float[] Flatten(List<Matrix> rows, int width)
    {
        var flat = new float[rows.Count * width];
        for (int t = 0; t < rows.Count; t++)
            for (int k = 0; k < width; k++) flat[t * width + k] = rows[t].GetValue(k, 0);
        return flat;
    }

    // NeuralNetwork.Save (NeuralNetwork.cs:159-176) + PPOAgent.Save (PPOAgent.cs:192-213): structure line, then per dense layer
    // "W <weights> B <biases>" (DenseLayer.Save, DenseLayer.cs:73-79)
    public void Save()
    {
        Hyperparameters.CreateDirectories();
        File.WriteAllLines($"{Hyperparameters.FilePath}{WeightsLocation}{Hyperparameters.CriticWeightFileName}.weights", Lines(1, Hyperparameters.CriticNeuralNetwork, _criticDense));
        File.WriteAllLines($"{Hyperparameters.FilePath}{WeightsLocation}{Hyperparameters.ActorWeightFileName}.weights", Lines(0, Hyperparameters.ActorNeuralNetwork, _actorDense));
    }

    private string[] Lines(int which, string structure, List<(int outSize, int inSize)> shapes)
    {
        Wb.wb_policy_num_params(_policy, which, out int count);
        var flat = new float[count];
        Wb.Ok(Wb.wb_policy_get_weights(_policy, which, flat), "wb_policy_get_weights");
        var lines = new List<string> { structure };
        int p = 0;
        foreach (var (o, i) in shapes)
        {
            string w = string.Join(" ", flat.Skip(p).Take(o * i));
            p += o * i;
            string b = string.Join(" ", flat.Skip(p).Take(o));
            p += o;
            lines.Add("W " + w + " B " + b);
        }
        return lines.ToArray();
    }

    // NeuralNetwork.Load (NeuralNetwork.cs:94-115): structure line must match, ValidateWeights on every line, then the lines that
    // are present replace the leading dense layers
    public void Load(string type)
    {
        string[] contents = type == "critic" ? Hyperparameters.CriticWeights : Hyperparameters.ActorWeights;
        string network = type == "critic" ? Hyperparameters.CriticNeuralNetwork : Hyperparameters.ActorNeuralNetwork;
        var shapes = type == "critic" ? _criticDense : _actorDense;
        int which = type == "critic" ? 1 : 0;
        if (contents.Length < 2 || contents[0] != network) return;
        (bool ok, _) = NeuralNetwork.ValidateWeights(contents);
        if (!ok || contents.Length - 1 > shapes.Count) return;
        Wb.wb_policy_num_params(_policy, which, out int count);
        var flat = new float[count];
        Wb.Ok(Wb.wb_policy_get_weights(_policy, which, flat), "wb_policy_get_weights");
        int p = 0;
        for (int l = 0; l < contents.Length - 1; l++)
        {
            var (o, i) = shapes[l];
            string line = contents[l + 1];
            int wi = line.IndexOf("W", StringComparison.Ordinal) + 2, bi = line.IndexOf("B", StringComparison.Ordinal) + 2;
            float[] w = Array.ConvertAll(line.Substring(wi, bi - wi - 3).Split(), float.Parse);
            float[] b = Array.ConvertAll(line.Substring(bi).Split(), float.Parse);
            if (w.Length < o * i || b.Length < o) return;           // Matrix.Load would throw (Matrix.cs:118-127)
            Array.Copy(w, 0, flat, p, o * i);
            p += o * i;
            Array.Copy(b, 0, flat, p, o);
            p += o;
        }
        Wb.Ok(Wb.wb_policy_set_weights(_policy, which, flat), "wb_policy_set_weights");
    }

    public void Dispose() => Wb.wb_policy_destroy(_policy);
}
