"""Second, independent restatement of the reference physics step in NumPy float32 scalars.

TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle/walker_oracle.h).  Purpose: cross-check the C oracle
bit for bit -- two restatements written separately from the reference text must agree before either is
trusted.  Object-per-body, list-based like the reference (Skeleton / RigidBody / Joint / Walker /
Environment classes), pure-Python loops: use it for a handful of env-steps only.

Every method cites the reference file:line it follows.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
ZERO = f32(0.0)
FLT_MAX = np.finfo(np.float32).max
PI_F = f32(3.14159274)   # MathF.PI
TAU_F = f32(6.28318548)  # MathF.Tau

np.seterr(all="ignore")


class Vec:
    """Microsoft.Xna.Framework.Vector2 (MonoGame; formulas per SURVEY.md Appendix C)."""

    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x = f32(x)
        self.y = f32(y)

    def __add__(self, o):
        return Vec(self.x + o.x, self.y + o.y)

    def __sub__(self, o):
        return Vec(self.x - o.x, self.y - o.y)

    def __neg__(self):
        return Vec(-self.x, -self.y)

    def __mul__(self, s):
        s = f32(s)
        return Vec(self.x * s, self.y * s)

    def __truediv__(self, d):
        factor = f32(1.0) / f32(d)
        return Vec(self.x * factor, self.y * factor)

    def __eq__(self, o):
        return bool(self.x == o.x and self.y == o.y)

    def dot(self, o):
        return f32(f32(self.x * o.x) + f32(self.y * o.y))

    def length(self):
        return f32(np.sqrt(f32(f32(self.x * self.x) + f32(self.y * self.y))))

    def normalized(self):
        val = f32(1.0) / self.length()
        return Vec(self.x * val, self.y * val)

    def copy(self):
        return Vec(self.x, self.y)


def net_min(a, b):
    """.NET Math.Min(float, float): NaN-propagating IEEE minimum."""
    if a != b:
        if not np.isnan(a):
            return a if a < b else b
        return a
    return a if np.signbit(a) else b


def net_max(a, b):
    if a != b:
        if not np.isnan(a):
            return a if b < a else b
        return a
    return a if np.signbit(b) else b


class Skeleton:
    """Objects/RigidBodies/Skeleton.cs"""

    def __init__(self):
        self.vectors = []
        self.centroid = Vec(0, 0)
        self.corners = [Vec(0, 0), Vec(0, 0)]

    def add_vectors(self, vs):  # Skeleton.cs:56-61, FindCentroid :100-113
        self.vectors.extend(v.copy() for v in vs)
        s = Vec(0, 0)
        for v in self.vectors:
            s = s + v
        self.centroid = s / f32(len(self.vectors))
        self.update_box()

    def update_box(self):  # BoundingBox.FindSignificantCorners, Skeleton.cs:144-176
        maxx = maxy = -FLT_MAX
        minx = miny = FLT_MAX
        for p in self.vectors:
            if p.x > maxx:
                maxx = p.x
            if p.y > maxy:
                maxy = p.y
            if p.x < minx:
                minx = p.x
            if p.y < miny:
                miny = p.y
        self.corners = [Vec(minx, miny), Vec(maxx, maxy)]

    def move(self, d):  # Skeleton.cs:76-85
        for i in range(len(self.vectors)):
            self.vectors[i] = self.vectors[i] + d
        self.centroid = self.centroid + d
        self.update_box()

    def rotate(self, angle):  # Skeleton.cs:89-97 + Matrix.CreateRotationZ + Vector2.Transform
        m11 = f32(math.cos(float(angle)))
        m12 = f32(math.sin(float(angle)))
        m21 = -m12
        m22 = m11
        for i in range(len(self.vectors)):
            p = self.vectors[i] - self.centroid
            x = f32(f32(f32(p.x * m11) + f32(p.y * m21)) + ZERO)
            y = f32(f32(f32(p.x * m12) + f32(p.y * m22)) + ZERO)
            self.vectors[i] = Vec(x, y) + self.centroid
        self.update_box()

    @staticmethod
    def is_colliding(a, b):  # BoundingBox.IsColliding, Skeleton.cs:133-140
        ca, cb = a.corners, b.corners
        return bool(ca[0].x < cb[1].x and ca[1].x > cb[0].x and ca[0].y < cb[1].y and ca[1].y > cb[0].y)


def project_points(axis, vectors):  # SATCollision.cs:63-76
    mn, mx = FLT_MAX, -FLT_MAX
    for v in vectors:
        t = axis.dot(v)
        if t < mn:
            mn = t
        if t > mx:
            mx = t
    return mn, mx


def axis_checks(va, vb, state, base):  # SATCollision.cs:39-59
    n = len(va)
    for i in range(n):
        edge = va[(i + 1) % n] - va[i]
        axis = Vec(-edge.y, edge.x)
        if axis == Vec(0, 0):
            continue
        axis = axis.normalized()
        amin, amax = project_points(axis, va)
        bmin, bmax = project_points(axis, vb)
        temp = net_min(f32(bmax - amin), f32(amax - bmin))  # Projection.IsOverlapping :100-104
        if not (amin < bmax and bmin < amax):
            return False
        if temp >= state["depth"]:
            continue
        state["depth"] = temp
        state["normal"] = axis
        state["axis"] = base + i
    return True


def sat_is_colliding(va, vb, ca, cb):  # SATCollision.cs:15-35
    st = {"normal": Vec(0, 0), "depth": FLT_MAX, "axis": -1}
    result = axis_checks(va, vb, st, 0) and axis_checks(vb, va, st, len(va))
    direction = cb - ca
    if direction.dot(st["normal"]) > ZERO:
        st["normal"] = st["normal"] * f32(-1.0)
    return result, st["normal"], st["depth"], st["axis"]


def cp_mod(a, b):  # ContactPoints.cs:131-134
    a = f32(a)
    b = f32(b)
    return int(round(float(a) - float(b) * math.floor(float(f32(a / b)))))


def significant_face(vs, normal):  # ContactPoints.cs:79-113
    idx = -1
    sv = Vec(0, 0)
    md = FLT_MAX
    for i, v in enumerate(vs):
        pr = v.dot(normal)
        if not (pr < md):
            continue
        sv, idx, md = v, i, pr
    n = len(vs)
    after = (sv - vs[(idx + 1) % n]).normalized()
    before = (sv - vs[cp_mod(idx - 1, n)]).normalized()
    if normal.dot(before) >= normal.dot(after):
        return (sv, vs[cp_mod(idx - 1, n)], sv)
    return (vs[(idx + 1) % n], sv, sv)


def clip_vectors(a, b, normal, offset):  # ContactPoints.cs:56-76
    pts = []
    da = f32(a.dot(normal) - offset)
    db = f32(b.dot(normal) - offset)
    if da >= ZERO:
        pts.append(a)
    if db >= ZERO:
        pts.append(b)
    if f32(da * db) < ZERO:
        edge = b - a
        loc = f32(da / f32(da - db))
        edge = edge * loc
        edge = edge + a
        pts.append(edge)
    return pts


def list_remove(pts, value):  # List<T>.Remove: first equal element
    for i, p in enumerate(pts):
        if p == value:
            del pts[i]
            return


def get_contact_points(va, vb, normal):  # ContactPoints.cs:13-53
    ref = significant_face(va, normal)
    rf = ref[1] - ref[0]
    inc = significant_face(vb, -normal)
    iv = inc[1] - inc[0]
    if abs(rf.dot(normal)) > abs(iv.dot(normal)):
        ref, inc = inc, ref
        rf = ref[1] - ref[0]
    rf = rf.normalized()
    offset = rf.dot(ref[0])
    cp = clip_vectors(inc[0], inc[1], rf, offset)
    if len(cp) < 2:
        return []
    offset = rf.dot(ref[1])
    cp = clip_vectors(cp[0], cp[1], -rf, -offset)
    if len(cp) < 2:
        return []
    rn = Vec(rf.y, -rf.x)
    maximum = rn.dot(ref[2])
    if f32(rn.dot(cp[0]) - maximum) < ZERO:
        list_remove(cp, cp[0])
    if f32(rn.dot(cp[-1]) - maximum) < ZERO:
        list_remove(cp, cp[-1])
    return cp


class RigidBody:
    """Bodies/RigidBody.cs"""

    def __init__(self, material, skeleton, is_static=False, is_floor=False, name=""):
        inv_mass, restitution, friction = (f32(v) for v in material)
        self.skeleton = skeleton
        self.associated = []
        self.is_static = is_static
        self.is_floor = is_floor
        self.collided = False
        self.restitution = restitution
        self.friction = friction
        self.inverse_mass = ZERO if is_static else inv_mass
        self.inverse_inertia = ZERO if is_static else f32(f32(0.001) * inv_mass)
        self.acceleration = Vec(0, 0)
        self.linear_velocity = Vec(0, 0)
        self.angular_velocity = ZERO
        self.angle = ZERO
        self.name = name
        self.trace = None

    def step(self, objects, dt):  # RigidBody.cs:54-61
        self.linear_velocity = self.linear_velocity + self.acceleration * dt  # :116-120
        self.skeleton.move(self.linear_velocity * dt)
        if self.is_static:
            return
        self.angle = f32(self.angle + f32(self.angular_velocity * dt))  # :123-129
        if self.angle > PI_F:  # WrapAngle :132-140
            self.angle = f32(self.angle - TAU_F)
        elif self.angle < -PI_F:
            self.angle = f32(self.angle + TAU_F)
        self.skeleton.rotate(f32(self.angular_velocity * dt))
        self.resolve_collisions(objects)

    def resolve_collisions(self, objects):  # RigidBody.cs:66-96
        for body in objects:
            if body is self:
                continue
            if any(body is b for b in self.associated):
                continue
            rec = {"other": body.name, "aabb": 0, "sat": 0, "axis": -1, "n": (ZERO, ZERO), "depth": ZERO, "contacts": []}
            if self.trace is not None:
                self.trace.append(rec)
            if not Skeleton.is_colliding(self.skeleton, body.skeleton):
                continue
            rec["aabb"] = 1
            if body.is_floor:
                self.collided = True
            if self.is_floor:
                body.collided = True
            va, vb = self.skeleton.vectors, body.skeleton.vectors
            ok, normal, depth, axis = sat_is_colliding(va, vb, self.skeleton.centroid, body.skeleton.centroid)
            if ok:
                cps = get_contact_points(va, vb, normal)
                rec.update(sat=1, axis=axis, n=(normal.x, normal.y), depth=depth, contacts=[(p.x, p.y) for p in cps])
                move_objects(self, body, normal, depth)
                impulses_resolve_collisions(self, body, cps, normal)


def move_objects(a, b, normal, depth):  # RigidBody.cs:99-113
    if a.is_static:
        b.skeleton.move((-normal) * depth)
    elif b.is_static:
        a.skeleton.move(normal * depth)
    else:
        a.skeleton.move(normal * depth / f32(2))
        b.skeleton.move((-normal) * depth / f32(2))


def calculate_impulse(A, B, contact, force, normal):  # Impulses.cs:86-115
    ra = contact - A.skeleton.centroid
    perp_a = Vec(-ra.y, ra.x)
    ka = normal.dot(perp_a)
    rb = contact - B.skeleton.centroid
    perp_b = Vec(-rb.y, rb.x)
    kb = normal.dot(perp_b)
    va = A.linear_velocity + perp_a * A.angular_velocity
    vb = B.linear_velocity + perp_b * B.angular_velocity
    vel = vb - va
    vn = vel.dot(normal)
    impulse = f32(-f32(force) * vn)
    denom = f32(f32(f32(A.inverse_mass + B.inverse_mass) + f32(f32(ka * ka) * A.inverse_inertia)) +
                f32(f32(kb * kb) * B.inverse_inertia))
    impulse = f32(impulse / denom)
    return ra, rb, impulse


def apply_impulses(A, B, normal, impulse, ra, rb):  # Impulses.cs:57-82
    J = normal * impulse
    va = A.linear_velocity - J * A.inverse_mass
    vb = B.linear_velocity + J * B.inverse_mass
    A.linear_velocity = va
    B.linear_velocity = vb
    perp_a = Vec(-ra.y, ra.x)
    wa = f32(A.angular_velocity - f32(perp_a.dot(J) * A.inverse_inertia))
    perp_b = Vec(-rb.y, rb.x)
    wb = f32(B.angular_velocity + f32(perp_b.dot(J) * B.inverse_inertia))
    A.angular_velocity = wa
    B.angular_velocity = wb


def impulses_resolve_collisions(A, B, cps, normal):  # Impulses.cs:12-28
    if len(cps) == 0:
        return
    restitution = net_max(A.restitution, B.restitution)
    friction = net_min(A.friction, B.friction)
    contact = (cps[0] + cps[1]) / f32(2) if len(cps) == 2 else cps[0]
    ra, rb, j = calculate_impulse(A, B, contact, f32(f32(1) + restitution), normal)
    tangent = Vec(-normal.y, normal.x)
    raf, rbf, jf = calculate_impulse(A, B, contact, friction, tangent)
    apply_impulses(A, B, normal, j, ra, rb)
    apply_impulses(A, B, tangent, jf, raf, rbf)


class Joint:
    """Objects/RigidBodies/Joint.cs"""

    def __init__(self, a, b, ia, ib):
        self.a, self.b, self.ia, self.ib = a, b, ia, ib
        self.current_torque = ZERO

    def point_a(self):
        return self.a.skeleton.vectors[self.ia]

    def point_b(self):
        return self.b.skeleton.vectors[self.ib]

    def step(self):  # Joint.cs:31-41
        ab = self.point_b() - self.point_a()
        depth = ab.length()
        if depth < f32(0.1):
            return
        ab = ab.normalized()
        self.a.skeleton.move(ab * depth / f32(2))
        self.b.skeleton.move((-ab) * depth / f32(2))
        # Impulses.ResolveJoint(_bodyB, _bodyA, ...), Impulses.cs:31-40
        contact = (self.point_a() + self.point_b()) / f32(2)
        ra, rb, j = calculate_impulse(self.b, self.a, contact, f32(f32(1) + f32(1)), ab)
        apply_impulses(self.b, self.a, ab, j, ra, rb)

    def set_torque(self, amount):  # Joint.cs:56-61
        amount = f32(amount)
        change = f32(amount - self.current_torque)
        self.current_torque = amount
        self.b.angular_velocity = f32(self.b.angular_velocity + f32(change * f32(5)))


def pole_from_size(material, c, size, name):  # Pole.cs:18-34
    adj = f32(f32(0.1) * f32(size))
    h = f32(adj * f32(3.5))
    sk = Skeleton()
    sk.add_vectors([Vec(c.x + adj, c.y + h), Vec(c.x, c.y + h), Vec(c.x - adj, c.y + h), Vec(c.x - adj, c.y - h),
                    Vec(c.x, c.y - h), Vec(c.x + adj, c.y - h)])
    return RigidBody(material, sk, name=name)


class Environment:
    """Environment.cs + Walker/Walker.cs (policy removed: actions are injected)."""

    NAMES = ["LLL", "LLU", "Body", "RLL", "RLU"]

    def __init__(self, floor=(15, 0.3, 1.0), walker=(5, 0.3, 0.8), iterations=50, max_timesteps=1000):
        self.floor_mat, self.walker_mat = floor, walker
        self.iterations, self.max_timesteps = iterations, max_timesteps
        self.rigid_bodies = []
        self.joints = []
        self.position = Vec(125, 800)
        self.previous_position = self.position
        self.terminal = False
        self.create_creature()
        sk = Skeleton()  # Environment.CreateFloor, Environment.cs:211-226
        sk.add_vectors([Vec(-50, 1050), Vec(-50, 900), Vec(1050, 900), Vec(1050, 1050)])
        self.floor = RigidBody(floor, sk, is_static=True, is_floor=True, name="floor")
        self.rigid_bodies.append(self.floor)
        self.steps = 0
        self.walker_update()  # InitialState, Environment.cs:176-180

    def create_creature(self):  # Walker.cs:40-46,155-209
        p = self.position
        sk = Skeleton()
        sk.add_vectors([Vec(p.x + f32(20), p.y + f32(20)), Vec(p.x, p.y + f32(20)), Vec(p.x - f32(20), p.y + f32(20)),
                        Vec(p.x - f32(20), p.y - f32(20)), Vec(p.x + f32(20), p.y - f32(20))])
        m = self.walker_mat
        self.body = RigidBody(m, sk, name="Body")
        self.body.inverse_inertia = f32(0.0003)
        self.llu = pole_from_size(m, p + Vec(0, 30), 75, "LLU")
        self.lll = pole_from_size(m, p + Vec(0, 60), 75, "LLL")
        self.rlu = pole_from_size(m, p + Vec(0, 30), 75, "RLU")
        self.rll = pole_from_size(m, p + Vec(0, 60), 75, "RLL")
        self.rigid_bodies.extend([self.lll, self.llu, self.body, self.rll, self.rlu])
        self.joints = [Joint(self.body, self.llu, 1, 4), Joint(self.body, self.rlu, 1, 4), Joint(self.llu, self.lll, 2, 3),
                       Joint(self.rlu, self.rll, 2, 3)]
        self.llu.associated = [self.rlu, self.rll, self.body]
        self.lll.associated = [self.rlu, self.rll, self.body]
        self.rlu.associated = [self.llu, self.lll, self.body]
        self.rll.associated = [self.llu, self.lll, self.body]
        self.body.associated = [self.llu, self.rlu, self.lll, self.rll]
        for b in (self.llu, self.lll, self.rlu, self.rll, self.body):
            b.acceleration = b.acceleration + Vec(0, 980)

    def dyn(self):
        return [self.lll, self.llu, self.body, self.rll, self.rlu]

    def walker_update(self):  # Walker.cs:49-54
        self.previous_position = self.position
        self.position = self.body.skeleton.centroid
        if self.body.collided or self.llu.collided or self.rlu.collided:
            self.terminal = True

    def reset(self):  # Environment.cs:167-173, Walker.cs:212-236
        self.steps = 0
        for b in (self.body, self.lll, self.llu, self.rll, self.rlu):
            self.rigid_bodies.remove(b)
        self.terminal = False
        self.position = Vec(125, 800)
        self.previous_position = self.position
        self.create_creature()
        self.walker_update()

    def take_actions(self, actions):  # Environment.cs:78 (Matrix.Clip, Matrix.cs:377-405), Walker.cs:66-75
        for j, a in zip(self.joints, actions):
            a = f32(a)
            if a >= f32(1):
                a = f32(1)
            elif a <= f32(-1):
                a = f32(-1)
            j.set_torque(a)

    def step_objects(self, dt, trace=None):  # Environment.cs:126-143
        dt = f32(f32(dt) / f32(self.iterations))
        for it in range(self.iterations):
            for j in self.joints:
                j.step()
            for b in self.rigid_bodies:
                b.trace = None
                if trace is not None and not b.is_static:
                    b.trace = []
                b.step(self.rigid_bodies, dt)
                if b.trace is not None:
                    trace.append((it, b.name, b.trace))

    def get_state(self):  # Walker.cs:132-152
        j = self.joints
        return np.array([j[0].point_a().x / f32(900), j[0].point_a().y / f32(500), j[2].point_a().x / f32(900),
                         j[2].point_a().y / f32(500), j[3].point_a().x / f32(900), j[3].point_a().y / f32(500),
                         self.body.linear_velocity.x / f32(60), self.body.linear_velocity.y / f32(60), self.lll.angle,
                         self.llu.angle, self.rll.angle, self.rlu.angle], dtype=np.float32)

    def step(self, actions, dt, auto_reset=False):  # Environment.Update/Step, Environment.cs:64-122,148-154
        self.steps += 1
        self.take_actions(actions)
        self.step_objects(dt)
        self.walker_update()
        dx = f32(self.position.x - self.previous_position.x)
        h = f32(self.joints[0].point_a().y / f32(500))
        reward = ZERO
        reward = f32(reward + (dx if (dx > ZERO and h < f32(1.6)) else ZERO))
        reward = f32(reward - (f32(-0.1) if h > f32(1.65) else ZERO))
        terminal = False
        if self.terminal or self.steps > self.max_timesteps:
            if self.terminal:
                reward = f32(reward - f32(40))
            terminal = True
        if self.position.x > f32(900):
            reward = f32(reward + f32(80))
            terminal = True
        obs = self.get_state()
        if terminal and auto_reset:
            self.reset()
            obs = self.get_state()
        return obs, reward, terminal

    def flat_state(self):
        """Canonical 92-float + 2-int record (oracle/walker_oracle.h)."""
        f = []
        for b in self.dyn():
            for v in b.skeleton.vectors:
                f += [v.x, v.y]
        for b in self.dyn():
            f += [b.skeleton.centroid.x, b.skeleton.centroid.y]
        for b in self.dyn():
            f += [b.linear_velocity.x, b.linear_velocity.y]
        f += [b.angular_velocity for b in self.dyn()]
        f += [b.angle for b in self.dyn()]
        f += [j.current_torque for j in self.joints]
        flags = 0
        for i, b in enumerate(self.dyn()):
            if b.collided:
                flags |= 1 << i
        if self.terminal:
            flags |= 1 << 5
        if self.rigid_bodies[0] is self.floor:
            flags |= 1 << 6
        return np.array(f, np.float32), np.array([flags, self.steps], np.int32)


# ---------------------------------------------------------------------------------------------------------------
# General scenes: the other IObject shapes and a bare Environment.StepObjects over any body / joint lists
# (checker for the wb_scene_* path)

def _body_from(material, vectors, is_static=False, is_floor=False, name=""):
    sk = Skeleton()
    sk.add_vectors(vectors)
    return RigidBody(material, sk, is_static=is_static, is_floor=is_floor, name=name)


def square_from_size(material, c, size, is_static=False, name="square"):  # Square.cs:18-32
    adj = f32(f32(0.5) * f32(size))
    return _body_from(material, [Vec(c.x + adj, c.y + adj), Vec(c.x - adj, c.y + adj), Vec(c.x - adj, c.y - adj), Vec(c.x + adj, c.y - adj)],
                      is_static, name=name)


def triangle_from_size(material, c, size, is_static=False, name="triangle"):  # Triangle.cs:18-31
    adj = f32(f32(0.5) * f32(size))
    return _body_from(material, [Vec(c.x, c.y + adj), Vec(c.x - adj, c.y - adj), Vec(c.x + adj, c.y - adj)], is_static, name=name)


def hexagon_from_size(material, c, size, is_static=False, name="hexagon"):  # Hexagon.cs:18-34
    adj = f32(f32(0.5) * f32(size))
    half = f32(adj * f32(0.5))
    return _body_from(material, [Vec(c.x + half, c.y + adj), Vec(c.x - half, c.y + adj), Vec(c.x - adj, c.y), Vec(c.x - half, c.y - adj),
                                 Vec(c.x + half, c.y - adj), Vec(c.x + adj, c.y)], is_static, name=name)


def hull_from_positions(material, positions, is_static=False, is_floor=False, name="hull"):  # Hull.cs:18-29
    return _body_from(material, [Vec(x, y) for x, y in positions], is_static, is_floor, name)


def smooth_corners(body, count=1):  # Skeleton.SmoothCorners, Skeleton.cs:33-53 (centroid and box are NOT refreshed)
    vs = body.skeleton.vectors
    for _ in range(count):
        new = []
        n = len(vs)
        for j in range(n):
            face_ab = (vs[(j + 1) % n] - vs[j]) * f32(0.2)
            face_ac = (vs[cp_mod(j - 1, n)] - vs[j]) * f32(0.2)
            new.append(vs[j] + face_ac)
            new.append(vs[j] + face_ab)
        vs = new
    body.skeleton.vectors = vs
    return body


class Scene:
    """Environment.StepObjects (Environment.cs:126-143) over arbitrary lists of bodies and joints."""

    def __init__(self, bodies, joints=(), iterations=50):
        self.bodies = list(bodies)
        self.joints = list(joints)
        self.iterations = iterations

    def set_torques(self, torques):  # Joint.SetTorque for every joint
        for j, t in zip(self.joints, torques):
            j.set_torque(t)

    def step_objects(self, dt):
        dt = f32(f32(dt) / f32(self.iterations))
        for _ in range(self.iterations):
            for j in self.joints:
                j.step()
            for b in self.bodies:
                b.trace = None
                b.step(self.bodies, dt)

    def flat_state(self):
        """Record of the wb_scene_* path: vertices, centroids, velocities, omega, angle, joint torques; collided bit mask."""
        f = []
        for b in self.bodies:
            for v in b.skeleton.vectors:
                f += [v.x, v.y]
        for b in self.bodies:
            f += [b.skeleton.centroid.x, b.skeleton.centroid.y]
        for b in self.bodies:
            f += [b.linear_velocity.x, b.linear_velocity.y]
        f += [b.angular_velocity for b in self.bodies]
        f += [b.angle for b in self.bodies]
        f += [j.current_torque for j in self.joints]
        collided = 0
        for i, b in enumerate(self.bodies):
            if b.collided:
                collided |= 1 << i
        return np.array(f, np.float32), collided
