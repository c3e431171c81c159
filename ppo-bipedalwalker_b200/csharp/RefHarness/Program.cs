// Program.cs -- RefHarness: replays a recorded rollout through the REFERENCE's physics classes and dumps the walker records.
//
// What runs is the reference's own code: RigidBody.Step / ResolveCollisions, SATCollision, ContactPoints, Impulses, Skeleton,
// Joint, Pole, Hull and the Materials (compiled from the reference checkout by RefHarness.csproj).  What is restated here,
// because Environment.cs and Walker/Walker.cs cannot be compiled without the MonoGame Game / Renderer / PPOAgent graph, is only
// the list handling around it, each line citing the reference lines it follows:
//   * Walker.CreateCreature: CreateBodies, CreateJoints, AddAssociatedBodies, AddAcceleration   (Walker.cs:40-46,155-209)
//   * Environment.CreateFloor (flat branch) with the floor material of the recording            (Environment.cs:211-226)
//   * Environment.Update up to the physics: Matrix.Clip, Walker.TakeActions, StepObjects         (Environment.cs:64-92,126-143)
//   * Environment.Reset / Walker.Reset: remove the walker's bodies, create a fresh creature      (Environment.cs:167-173, Walker.cs:212-236)
// Rewards, observations and the terminal test are NOT needed: the recording says at which steps a walker was reset.
//
// Input  (scripts/refharness_io.py export):  line 1 "n steps iterations"; line 2 the n floor material names; then steps x n lines
//         "a0 a1 a2 a3 done" with the four actions as 8-digit hex float bits.
// Output (compared by scripts/refharness_io.py compare): steps x n lines of 92 hex words in the wb record order --
//         vertices of LLL, LLU, Body, RLL, RLU (x, y), centroids, linear velocities, angular velocities, angles, joint torques --
//         taken after the step and after the reset that may follow it (the recording stores the post-reset record).
using System;
using System.Collections.Generic;
using System.Globalization;
using System.IO;
using Microsoft.Xna.Framework;
using NEA.Bodies;
using NEA.Materials;
using NEA.Objects;
using NEA.Objects.RigidBodies;

namespace RefHarness;

internal sealed class WalkerEnv
{
    private readonly List<RigidBody> _rigidBodies = new List<RigidBody>();
    private readonly List<Joint> _joints = new List<Joint>();
    private readonly IMaterial _walkerMaterial = new Carpet();                 // Walker.cs:30
    private Vector2 _position = new Vector2(125, 800);                        // Walker.cs:31
    private Hull _body;
    private Pole _llu, _lll, _rlu, _rll;

    public WalkerEnv(IMaterial floorMaterial)
    {
        CreateCreature();                                                     // Environment.cs:46-47
        Vector2[] floorPositions = { new(-50, 1050), new(-50, 900), new(1050, 900), new(1050, 1050) };   // Environment.cs:219-221
        _rigidBodies.Add(Hull.FromPositions(floorMaterial, floorPositions, isStatic: true, isFloor: true));
    }

    private void CreateCreature()                                             // Walker.cs:40-46
    {
        // CreateBodies, Walker.cs:155-178
        Skeleton bodySkeleton = new Skeleton();
        bodySkeleton.AddVectors(new Vector2[]
        {
            new(_position.X + 20, _position.Y + 20), new(_position.X, _position.Y + 20), new(_position.X - 20, _position.Y + 20),
            new(_position.X - 20, _position.Y - 20), new(_position.X + 20, _position.Y - 20)
        });
        _body = Hull.FromSkeleton(_walkerMaterial, bodySkeleton);
        _body.SetInverseInertia(0.0003f);
        _llu = Pole.FromSize(_walkerMaterial, _position + new Vector2(0, 30), 75);
        _lll = Pole.FromSize(_walkerMaterial, _position + new Vector2(0, 60f), 75);
        _rlu = Pole.FromSize(_walkerMaterial, _position + new Vector2(0, 30), 75);
        _rll = Pole.FromSize(_walkerMaterial, _position + new Vector2(0, 60f), 75);
        _rigidBodies.AddRange(new RigidBody[] { _lll, _llu, _body, _rll, _rlu });
        // CreateJoints, Walker.cs:181-189
        _joints.AddRange(new[] { new Joint(_body, _llu, 1, 4), new Joint(_body, _rlu, 1, 4), new Joint(_llu, _lll, 2, 3), new Joint(_rlu, _rll, 2, 3) });
        // AddAssociatedBodies, Walker.cs:204-209
        _llu.AddAssociatedBodies(new RigidBody[] { _rlu, _rll, _body });
        _lll.AddAssociatedBodies(new RigidBody[] { _rlu, _rll, _body });
        _rlu.AddAssociatedBodies(new RigidBody[] { _llu, _lll, _body });
        _rll.AddAssociatedBodies(new RigidBody[] { _llu, _lll, _body });
        _body.AddAssociatedBodies(new RigidBody[] { _llu, _rlu, _lll, _rll });
        // AddAcceleration(new Vector2(0, 980)), Walker.cs:45,192-199
        Vector2 g = new Vector2(0, 980);
        _llu.AddAcceleration(g);
        _lll.AddAcceleration(g);
        _rlu.AddAcceleration(g);
        _rll.AddAcceleration(g);
        _body.AddAcceleration(g);
    }

    public void Reset()                                                       // Environment.cs:167-173 -> Walker.cs:212-236
    {
        if (_rigidBodies.Count >= 6)
        {
            _rigidBodies.Remove(_body);
            _rigidBodies.Remove(_lll);
            _rigidBodies.Remove(_llu);
            _rigidBodies.Remove(_rll);
            _rigidBodies.Remove(_rlu);
        }
        _joints.Clear();
        _position = new Vector2(125, 800);
        CreateCreature();                                                     // appended AFTER the floor: the floor now comes first
    }

    private static float Clip(float v) => v >= 1f ? 1f : (v <= -1f ? -1f : v);  // Matrix.Clip(actions, 1, -1), Matrix.cs:377-405, Environment.cs:78

    public void Step(float[] actions, float deltaTime, int iterations)
    {
        for (int i = 0; i < _joints.Count; i++) _joints[i].SetTorque(Clip(actions[i]));   // Walker.TakeActions, Walker.cs:66-75
        deltaTime /= iterations;                                              // Environment.StepObjects, Environment.cs:126-143
        for (int i = 0; i < iterations; i++)
        {
            foreach (Joint joint in _joints) joint.Step();
            foreach (RigidBody body in _rigidBodies) ((IObject)body).Update(_rigidBodies, deltaTime);
        }
    }

    public void Dump(TextWriter w)
    {
        RigidBody[] order = { _lll, _llu, _body, _rll, _rlu };
        var words = new List<float>(92);
        foreach (RigidBody b in order)
            foreach (Vector2 v in b.GetVectors()) { words.Add(v.X); words.Add(v.Y); }
        foreach (RigidBody b in order) { Vector2 c = b.GetCentroid(); words.Add(c.X); words.Add(c.Y); }
        foreach (RigidBody b in order) { Vector2 v = b.GetLinearVelocity(); words.Add(v.X); words.Add(v.Y); }
        foreach (RigidBody b in order) words.Add(b.GetAngularVelocity());
        foreach (RigidBody b in order) words.Add(b.GetAngle());
        foreach (Joint j in _joints) words.Add(j.GetTorque());
        if (words.Count != 92) throw new Exception("record has " + words.Count + " words, expected 92");
        for (int i = 0; i < words.Count; i++)
        {
            if (i > 0) w.Write(' ');
            w.Write(BitConverter.SingleToInt32Bits(words[i]).ToString("x8"));
        }
        w.WriteLine();
    }
}

internal static class Program
{
    private static IMaterial MaterialByName(string name) => name switch
    {
        "Ice" => new Ice(), "Wood" => new Wood(), "Paper" => new Paper(), "Titanium" => new Titanium(), "Carpet" => new Carpet(),
        "Rubber" => new Rubber(), "Metal" => new Metal(), "SuperRubber" => new SuperRubber(),
        _ => throw new ArgumentException("unknown material " + name)
    };

    private static float FromHex(string s) => BitConverter.Int32BitsToSingle(int.Parse(s, NumberStyles.HexNumber, CultureInfo.InvariantCulture));

    private static int Main(string[] args)
    {
        if (args.Length != 2)
        {
            Console.Error.WriteLine("usage: RefHarness <golden_actions.txt> <refharness_states.txt>");
            return 2;
        }
        string[] lines = File.ReadAllLines(args[0]);
        string[] head = lines[0].Split(' ', StringSplitOptions.RemoveEmptyEntries);
        int n = int.Parse(head[0]), steps = int.Parse(head[1]), iterations = int.Parse(head[2]);
        string[] floors = lines[1].Split(' ', StringSplitOptions.RemoveEmptyEntries);
        var envs = new WalkerEnv[n];
        for (int e = 0; e < n; e++) envs[e] = new WalkerEnv(MaterialByName(floors[e]));
        // Game.TargetElapsedTime = TimeSpan.FromTicks(166667) with a fixed time step; Game1.cs:73 casts TotalSeconds to float
        float deltaTime = (float)TimeSpan.FromTicks(166667).TotalSeconds;
        using var w = new StreamWriter(args[1]);
        int line = 2;
        for (int t = 0; t < steps; t++)
        {
            for (int e = 0; e < n; e++, line++)
            {
                string[] f = lines[line].Split(' ', StringSplitOptions.RemoveEmptyEntries);
                float[] a = { FromHex(f[0]), FromHex(f[1]), FromHex(f[2]), FromHex(f[3]) };
                envs[e].Step(a, deltaTime, iterations);
                if (f[4] == "1") envs[e].Reset();          // the recording's record is the one AFTER Reset + InitialState (auto-reset)
                envs[e].Dump(w);
            }
        }
        Console.Error.WriteLine("RefHarness: " + steps + " env-steps x " + n + " walkers written; ErrorLogger calls: " + NEA.Rendering.ErrorLogger.Count);
        return 0;
    }
}
