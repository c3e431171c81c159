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
