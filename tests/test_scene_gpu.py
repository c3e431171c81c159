"""GPU parity tests of the general-scene path (wb_scene_*, SURVEY 8f row 4: IObject plugins, other shapes, SmoothCorners,
arbitrary body lists) against the NumPy restatement of the reference engine (oracle/np_oracle.py), bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _np_oracle():
    import __graft_entry__ as ge
    ge.load_oracle()
    import np_oracle
    return np_oracle


MAT = {"Metal": (15, 0.3, 1.0), "Rubber": (11, 0.7, 0.5), "Wood": (20, 0.3, 0.01), "Ice": (11, 0.3, 0.0), "Carpet": (5, 0.3, 0.8)}


def build_pile(gpu, P, n_copies, spins):
    """A static floor hull, a static wall, and a pile of mixed shapes (one smoothed square), two of them jointed; gravity on."""
    V = P.Vec
    g = (0.0, 980.0)
    floor_pts = [(-50, 600), (-50, 480), (700, 480), (700, 600)]
    wall_pts = [(-50, 500), (-50, 100), (0, 100), (0, 500)]
    spec = [
        ("hull", "Metal", floor_pts, True, True), ("hull", "Metal", wall_pts, True, False),
        ("square", "Rubber", (120, 440), 60), ("square", "Wood", (150, 384), 50), ("triangle", "Ice", (260, 450), 70),
        ("hexagon", "Carpet", (330, 430), 80), ("pole", "Carpet", (420, 440), 75), ("hexagon", "Rubber", (140, 290), 60),
        ("smooth", "Wood", (250, 330), 64), ("triangle", "Metal", (40, 450), 60),
    ]
    objs, bodies = [], []
    for s in spec:
        kind, mat = s[0], s[1]
        if kind == "hull":
            objs.append(gpu.Hull.FromPositions(mat, s[2], isStatic=s[3], isFloor=s[4]))
            bodies.append(P.hull_from_positions(MAT[mat], s[2], is_static=s[3], is_floor=s[4], name=f"b{len(bodies)}"))
            continue
        c, size = s[2], s[3]
        if kind == "square":
            o, b = gpu.Square.FromSize(mat, c, size), P.square_from_size(MAT[mat], V(*c), size)
        elif kind == "triangle":
            o, b = gpu.Triangle.FromSize(mat, c, size), P.triangle_from_size(MAT[mat], V(*c), size)
        elif kind == "hexagon":
            o, b = gpu.Hexagon.FromSize(mat, c, size), P.hexagon_from_size(MAT[mat], V(*c), size)
        elif kind == "pole":
            o, b = gpu.Pole.FromSize(mat, c, size), P.pole_from_size(MAT[mat], V(*c), size, "pole")
        else:  # a square with smoothed corners: 8 vertices (Skeleton.SmoothCorners)
            o = gpu.Square.FromSize(mat, c, size).SmoothCorners(1)
            b = P.smooth_corners(P.square_from_size(MAT[mat], V(*c), size), 1)
        o.acceleration = g
        b.acceleration = V(*g)
        objs.append(o)
        bodies.append(b)
    # the jointed pair never collides with each other (association lists, RigidBody.cs:143-152); one custom inverse inertia
    objs[5].associated.append(objs[6]); objs[6].associated.append(objs[5])
    bodies[5].associated.append(bodies[6]); bodies[6].associated.append(bodies[5])
    objs[7].inverseInertia = 0.0004
    bodies[7].inverse_inertia = np.float32(0.0004)
    joints = [gpu.Joint(objs[5], objs[6], 5, 2)]
    pj = [P.Joint(bodies[5], bodies[6], 5, 2)]
    for i, w in enumerate(spins):
        bodies[2 + i].angular_velocity = np.float32(w)
    return objs, joints, P.Scene(bodies, pj, iterations=50)


def test_mixed_shape_pile_bit_exact(gpu):
    P = _np_oracle()
    rng = np.random.default_rng(3)
    n = 3
    spins = rng.uniform(-3, 3, 8).astype(np.float32)
    objs, joints, ref = build_pile(gpu, P, n, spins)
    scene = gpu.Scene(n, objs, joints)
    f, col = scene.get_state()
    rf, rcol = ref.flat_state()
    # identical start: shape factories, SmoothCorners and the cached centroids agree before any spin is injected
    rf0 = rf.copy()
    nb = len(objs)
    tv = sum(scene.vertex_counts)
    rf0[2 * tv + 4 * nb:2 * tv + 5 * nb] = 0
    assert np.array_equal(bits(f[0]), bits(rf0))
    f[:, 2 * tv + 4 * nb + 2:2 * tv + 4 * nb + 10] = spins  # angular velocities of the eight dynamic bodies
    scene.set_state(f, col)
    torques = np.zeros((n, 1), np.float32)
    for step in range(10):
        torques[:] = np.float32(0.3 * (step % 3 - 1))
        scene.SetTorques(torques)
        ref.set_torques([torques[0, 0]])
        scene.StepObjects(gpu.DT_FRAME)
        ref.step_objects(gpu.DT_FRAME)
        f, col = scene.get_state()
        rf, rcol = ref.flat_state()
        for c in range(n):
            assert np.array_equal(bits(f[c]), bits(rf)), f"state differs at step {step}, copy {c}"
            assert int(col[c]) == rcol, f"collided flags differ at step {step}"
    assert rcol != 0  # something touched the floor


def test_scene_rejects_bad_descriptions(gpu):
    tri = gpu.Triangle.FromSize("Wood", (0, 0), 10)
    two = gpu.IObject(np.zeros((2, 2), np.float32))
    with pytest.raises(gpu.WalkerB200Error):
        gpu.Scene(1, [two])
    with pytest.raises(gpu.WalkerB200Error):
        gpu.Scene(1, [tri] * 17)
    with pytest.raises(gpu.WalkerB200Error):
        gpu.Scene(1, [tri, gpu.Square.FromSize("Wood", (5, 5), 4)], [gpu.Joint(tri, tri, 0, 7)])


def _oracle_walker_scene(P, floor_bodies, floor_first=False):
    """The NumPy restatement's walker (Walker.CreateCreature) + the given static floor bodies as one P.Scene."""
    env = P.Environment()
    dyn = env.dyn()
    bodies = (floor_bodies + dyn) if floor_first else (dyn + floor_bodies)
    return P.Scene(bodies, env.joints, iterations=50)


def test_walker_scene_agrees_with_the_specialised_walker_kernel(gpu):
    """The reference's default scene (CreateCreature + CreateFloor) through the GENERAL engine (scene_step_kernel) against the
    specialised walker kernel (physics_lanes_kernel) on the same clipped actions: two independent CUDA implementations of
    Environment.StepObjects, bit for bit on every vertex, centroid, velocity, angular velocity, angle and joint torque."""
    n = 5
    objs, joints = gpu.CreateCreature()
    scene = gpu.Scene(n, objs + gpu.CreateFloor("Metal"), joints)      # constructor order: walker bodies, then the floor
    env = gpu.EnvBatch(n, floor_materials="Metal")
    rng = np.random.default_rng(17)
    for step in range(12):
        a = np.clip(rng.uniform(-1.3, 1.3, (n, 4)).astype(np.float32), -1, 1)
        env.take_actions(a)
        env.step_objects(gpu.DT_FRAME)
        scene.SetTorques(a)
        scene.StepObjects(gpu.DT_FRAME)
        f, _ = env.get_state()      # [n, 92]: 29 vertices, 5 centroids, 5 velocities, 5 omega, 5 angles, 4 torques
        g, col = scene.get_state()  # [n, 106]: 33 vertices, 6 centroids, 6 velocities, 6 omega, 6 angles, 4 torques
        assert g.shape[1] == 106
        pieces = [(f[:, 0:58], g[:, 0:58]), (f[:, 58:68], g[:, 66:76]), (f[:, 68:78], g[:, 78:88]), (f[:, 78:83], g[:, 90:95]),
                  (f[:, 83:88], g[:, 96:101]), (f[:, 88:92], g[:, 102:106])]
        for k, (x, y) in enumerate(pieces):
            assert np.array_equal(bits(x), bits(y)), f"piece {k} differs at env-step {step}"
    assert (col != 0).any()  # feet touched the floor


def test_walker_on_rough_floor_bit_exact(gpu):
    """Environment.CreateRoughFloor (Environment.cs:230-261; the random heights are injected) + the walker: 15 bodies, 4 joints,
    against the NumPy restatement, for both list orders (floor last as constructed, floor first as after Walker.Reset)."""
    P = _np_oracle()
    heights = [int(h) for h in np.random.default_rng(5).integers(0, 100, 11)]
    n = 2
    for floor_first in (False, True):
        objs, joints = gpu.CreateCreature()
        rough = gpu.CreateRoughFloor(heights)
        scene = gpu.Scene(n, (rough + objs) if floor_first else (objs + rough), joints)
        m = MAT["Metal"]
        initial_x, movement = -50, 120
        prev = (initial_x, 800 + heights[0])
        fb = []
        for i in range(10):
            x, y = initial_x + i * movement, 800 + heights[1 + i]
            fb.append(P.hull_from_positions(m, [(x, 1050), prev, (x, y), (x + movement, 1050)], is_static=True, is_floor=True, name=f"floor{i}"))
            prev = (x, y)
        ref = _oracle_walker_scene(P, fb, floor_first)
        f, col = scene.get_state()
        rf, rcol = ref.flat_state()
        assert np.array_equal(bits(f[0]), bits(rf)) and np.array_equal(bits(f[1]), bits(rf))
        rng = np.random.default_rng(23)
        for step in range(4):
            a = np.clip(rng.uniform(-1.2, 1.2, 4).astype(np.float32), -1, 1)
            scene.SetTorques(np.tile(a, (n, 1)))
            ref.set_torques(list(a))
            scene.StepObjects(gpu.DT_FRAME)
            ref.step_objects(gpu.DT_FRAME)
            f, col = scene.get_state()
            rf, rcol = ref.flat_state()
            for c in range(n):
                assert np.array_equal(bits(f[c]), bits(rf)), f"state differs at env-step {step} (floor_first={floor_first})"
                assert int(col[c]) == rcol
        assert rcol != 0
