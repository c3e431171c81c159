// Adapters.cs -- the reference's two plugin interfaces forwarded to libwalker_b200 (source only; see INTEGRATION.md):
//   IMaterial (Materials/IMaterial.cs:6-11)  -> wb_material_register: any user material becomes a material id
//   IObject   (Objects/IObject.cs:7-10)      -> wb_scene_create: any List<RigidBody> (Pole / Hull / Square / Triangle / Hexagon /
//                                               user shapes), its association lists and a List<Joint> become N lockstep copies of
//                                               the scene, stepped by wb_scene_step_objects (= Environment.StepObjects)
using System;
using System.Collections.Generic;
using System.Linq;
using System.Reflection;
using Microsoft.Xna.Framework;
using NEA.Bodies;
using NEA.Materials;
using NEA.Native;
using NEA.Objects.RigidBodies;

namespace NEA;

public static class MaterialAdapter
{
    private static readonly Dictionary<Type, int> Ids = new()
    {
        [typeof(Ice)] = 0, [typeof(Wood)] = 1, [typeof(Paper)] = 2, [typeof(Titanium)] = 3,
        [typeof(Carpet)] = 4, [typeof(Rubber)] = 5, [typeof(Metal)] = 6, [typeof(SuperRubber)] = 7,
    };

    // built-in materials map to their fixed ids; any other IMaterial implementation is registered once per type
    public static int IdOf(IMaterial material)
    {
        if (Ids.TryGetValue(material.GetType(), out int id)) return id;
        Wb.Ok(Wb.wb_material_register(material.InverseMass, material.Restitution, material.Friction, out id), "wb_material_register");
        Ids[material.GetType()] = id;
        return id;
    }

    // a RigidBody keeps its material's three numbers, not the IMaterial object (RigidBody ctor, RigidBody.cs:36-50); a static body
    // reports inverse mass 0 (the library applies that itself from is_static), so it is registered by restitution / friction alone
    private static readonly Dictionary<(float, float, float), int> ByValue = new();
    public static int IdOf(RigidBody body)
    {
        var key = (body.GetInverseMass(), body.GetRestitution(), body.GetFriction());
        if (ByValue.TryGetValue(key, out int id)) return id;
        Wb.Ok(Wb.wb_material_register(key.Item1, key.Item2, key.Item3, out id), "wb_material_register");
        ByValue[key] = id;
        return id;
    }
}

// N lockstep copies of an arbitrary scene: the list the reference keeps in Environment._rigidBodies, in list order
public sealed class GpuScene : IDisposable
{
    private readonly IntPtr _scene;
    private readonly int _copies, _floats;
    public GpuScene(List<RigidBody> rigidBodies, List<Joint> joints, int copies)
    {
        _copies = copies;
        var bodies = new WbBodyDesc[rigidBodies.Count];
        var vertices = new List<float>();
        for (int b = 0; b < rigidBodies.Count; b++)
        {
            RigidBody body = rigidBodies[b];
            List<Vector2> v = body.GetVectors();
            foreach (var p in v) { vertices.Add(p.X); vertices.Add(p.Y); }
            uint mask = 0;
            foreach (RigidBody other in Private<List<RigidBody>>(body, "_associatedBodies"))       // RigidBody.cs:143-152
                if (rigidBodies.IndexOf(other) is int j and >= 0) mask |= 1u << j;
            Vector2 accel = Private<Vector2>(body, "_acceleration");                                 // RigidBody.AddAcceleration, :156-159
            bodies[b] = new WbBodyDesc
            {
                NVertices = v.Count, IsStatic = body.IsStatic() ? 1 : 0, IsFloor = Private<bool>(body, "_isFloor") ? 1 : 0,
                Material = MaterialAdapter.IdOf(body), AssociatedMask = mask, AccelX = accel.X, AccelY = accel.Y,
                InverseInertia = body.GetInverseInertia(),                                           // incl. SetInverseInertia overrides (Walker.cs:168)
            };
        }
        var jd = joints.Select(j => new WbJointDesc
        {
            BodyA = rigidBodies.IndexOf(Private<RigidBody>(j, "_bodyA")), VertexA = Private<int>(j, "_indexA"),
            BodyB = rigidBodies.IndexOf(Private<RigidBody>(j, "_bodyB")), VertexB = Private<int>(j, "_indexB"),
        }).ToArray();
        Wb.Ok(Wb.wb_init(0), "wb_init");
        Wb.Ok(Wb.wb_scene_create(copies, bodies, bodies.Length, vertices.ToArray(), jd, jd.Length, NEA.Walker.PPO.Hyperparameters.Iterations, out _scene),
              "wb_scene_create");
        Wb.wb_scene_state_floats(_scene, out _floats);
    }

    // Joint.SetTorque for every joint of every copy (Joint.cs:56-61); torques [copies][joints]
    public void SetTorques(float[] torques) => Wb.Ok(Wb.wb_scene_set_torques(_scene, torques), "wb_scene_set_torques");

    // Environment.StepObjects (Environment.cs:126-143): every Joint.Step, then every IObject.Update in list order, Iterations times
    public void StepObjects(float deltaTime) => Wb.Ok(Wb.wb_scene_step_objects(_scene, deltaTime), "wb_scene_step_objects");

    // per-copy record (vertices, centroids, velocities, angular velocities, angles, torques) as SoA [floats][copies] + Collided bits
    public (float[] state, int[] collided) GetState()
    {
        var state = new float[_floats * _copies];
        var collided = new int[_copies];
        Wb.Ok(Wb.wb_scene_get_state(_scene, state, collided), "wb_scene_get_state");
        return (state, collided);
    }

    private static T Private<T>(object o, string field)
    {
        for (Type t = o.GetType(); t != null; t = t.BaseType)
        {
            FieldInfo f = t.GetField(field, BindingFlags.Instance | BindingFlags.NonPublic | BindingFlags.Public);
            if (f != null) return (T)f.GetValue(o);
        }
        throw new MissingFieldException(o.GetType().Name, field);
    }

    public void Dispose() => Wb.wb_scene_destroy(_scene);
}
