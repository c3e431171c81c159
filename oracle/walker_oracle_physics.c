/*
 * walker_oracle_physics.c -- CPU ORACLE (test infrastructure, NOT the product). PARITY UNPINNED, see walker_oracle.h.
 *
 * Strict-binary32 restatement of the reference's rigid-body step.  Every function cites the
 * reference file:line it follows (paths relative to the reference root).  All arithmetic is
 * float, never contracted to FMA (-ffp-contract=off); the only double excursions are the ones
 * the reference has: Math.Cos/Math.Sin inside Matrix.CreateRotationZ.
 */
#include "walker_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Vector2 ----
 * MonoGame Microsoft.Xna.Framework.Vector2 (not vendored in the reference). */
typedef struct {
  float x, y;
} v2;

static inline v2 V(float x, float y) {
  v2 r;
  r.x = x;
  r.y = y;
  return r;
}
static inline v2 v_add(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 v_sub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 v_neg(v2 a) { return V(-a.x, -a.y); }
static inline v2 v_mul(v2 a, float s) { return V(a.x * s, a.y * s); }
/* Vector2 / float and Vector2.Divide: reciprocal, then two multiplies */
static inline v2 v_div(v2 a, float d) {
  float factor = 1.0f / d;
  return V(a.x * factor, a.y * factor);
}
static inline float v_dot(v2 a, v2 b) { return (a.x * b.x) + (a.y * b.y); }
static inline float v_length(v2 a) { return sqrtf((a.x * a.x) + (a.y * a.y)); }
static inline v2 v_normalize(v2 a) {
  float val = 1.0f / sqrtf((a.x * a.x) + (a.y * a.y));
  return V(a.x * val, a.y * val);
}

/* .NET Math.Min/Math.Max(float,float): IEEE 754-2019 minimum/maximum, NaN-propagating */
static inline float net_min(float a, float b) {
  if (a != b) {
    if (!isnan(a)) return a < b ? a : b;
    return a;
  }
  return signbit(a) ? a : b;
}
static inline float net_max(float a, float b) {
  if (a != b) {
    if (!isnan(a)) return b < a ? a : b;
    return a;
  }
  return signbit(b) ? a : b;
}

/* Materials/{Ice..SuperRubber}.cs:7-9 */
static const wo_material k_materials[WO_NUM_MATERIALS] = {
    {11.0f, 0.3f, 0.0f},   /* Ice */
    {20.0f, 0.3f, 0.01f},  /* Wood */
    {1.0f, 0.3f, 0.1f},    /* Paper */
    {0.01f, 0.1f, 0.2f},   /* Titanium */
    {5.0f, 0.3f, 0.8f},    /* Carpet */
    {11.0f, 0.7f, 0.5f},   /* Rubber */
    {15.0f, 0.3f, 1.0f},   /* Metal */
    {11.0f, 1.0f, 1.0f},   /* SuperRubber */
};
const wo_material* wo_builtin_material(int id) {
  if (id < 0 || id >= WO_NUM_MATERIALS) return NULL;
  return &k_materials[id];
}

/* ------------------------------------------------------------------ Skeleton / RigidBody ---- */
#define MAXV 8
typedef struct {
  /* Skeleton (Objects/RigidBodies/Skeleton.cs:11-21) */
  v2 verts[MAXV];
  int n;
  v2 centroid;
  v2 bb_min, bb_max;
  /* RigidBody (Bodies/RigidBody.cs:16-33) */
  unsigned assoc; /* bit per body id */
  int is_static, is_floor, collided;
  float inv_mass, inv_inertia, restitution, friction;
  v2 accel, vel;
  float omega, angle;
} body_t;

struct wo_env {
  body_t body[6]; /* indexed by body id */
  int order[6];   /* Environment._rigidBodies list order */
  int nbodies;
  int joint_a[4], joint_b[4], joint_ia[4], joint_ib[4]; /* Walker.CreateJoints, Walker.cs:180-188 */
  float joint_torque[4];
  v2 pos, prev_pos; /* Walker._position/_previousPosition */
  int terminal;
  int steps;
  int floor_first;
  wo_material floor_mat, walker_mat;
};

int wo_env_sizeof(void) { return (int)sizeof(struct wo_env); }

/* BoundingBox.FindSignificantCorners, Skeleton.cs:144-176 */
static void bb_update(body_t* b) {
  float maxx = -FLT_MAX, maxy = -FLT_MAX, minx = FLT_MAX, miny = FLT_MAX;
  for (int i = 0; i < b->n; i++) {
    v2 p = b->verts[i];
    if (p.x > maxx) maxx = p.x;
    if (p.y > maxy) maxy = p.y;
    if (p.x < minx) minx = p.x;
    if (p.y < miny) miny = p.y;
  }
  b->bb_min = V(minx - 0.0f, miny - 0.0f);
  b->bb_max = V(maxx + 0.0f, maxy + 0.0f);
}

/* Skeleton.AddVectors + FindCentroid, Skeleton.cs:56-61,100-113 */
static void sk_add_vectors(body_t* b, const v2* verts, int n) {
  b->n = n;
  v2 sum = V(0.0f, 0.0f);
  for (int i = 0; i < n; i++) {
    b->verts[i] = verts[i];
    sum = v_add(sum, verts[i]);
  }
  b->centroid = v_div(sum, (float)n);
  bb_update(b);
}

/* Skeleton.Move, Skeleton.cs:76-85 */
static void sk_move(body_t* b, v2 d) {
  for (int i = 0; i < b->n; i++) b->verts[i] = v_add(b->verts[i], d);
  b->centroid = v_add(b->centroid, d);
  bb_update(b);
}

/* Matrix.CreateRotationZ: double trig on the widened float, rounded to float (Appendix C) */
void wo_rotz(float radians, float* c, float* s) {
  *c = (float)cos((double)radians);
  *s = (float)sin((double)radians);
}

/* Skeleton.Rotate, Skeleton.cs:89-97 (Vector2.Transform with M41 = M42 = 0) */
static void sk_rotate(body_t* b, float angle) {
  float m11, m12;
  wo_rotz(angle, &m11, &m12);
  float m21 = -m12, m22 = m11;
  for (int i = 0; i < b->n; i++) {
    v2 p = v_sub(b->verts[i], b->centroid);
    v2 t = V(((p.x * m11) + (p.y * m21)) + 0.0f, ((p.x * m12) + (p.y * m22)) + 0.0f);
    b->verts[i] = v_add(t, b->centroid);
  }
  bb_update(b);
}

/* BoundingBox.IsColliding, Skeleton.cs:133-140 */
static int bb_colliding(const body_t* a, const body_t* b) {
  return (a->bb_min.x < b->bb_max.x && a->bb_max.x > b->bb_min.x && a->bb_min.y < b->bb_max.y && a->bb_max.y > b->bb_min.y);
}

/* RigidBody ctor, RigidBody.cs:36-50 */
static void rb_init(body_t* b, wo_material m, int is_static, int is_floor) {
  b->assoc = 0;
  b->is_static = is_static;
  b->is_floor = is_floor;
  b->collided = 0;
  b->restitution = m.restitution;
  b->friction = m.friction;
  b->inv_mass = is_static ? 0.0f : m.inverse_mass;
  b->inv_inertia = is_static ? 0.0f : 0.001f * m.inverse_mass;
  b->accel = V(0.0f, 0.0f);
  b->vel = V(0.0f, 0.0f);
  b->omega = 0.0f;
  b->angle = 0.0f;
}

/* ------------------------------------------------------------------ SAT ----
 * SATCollision.ProjectPoints, SATCollision.cs:63-76 */
static void project(v2 axis, const v2* p, int n, float* omin, float* omax) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int i = 0; i < n; i++) {
    float t = v_dot(axis, p[i]);
    if (t < mn) mn = t;
    if (t > mx) mx = t;
  }
  *omin = mn;
  *omax = mx;
}

/* SATCollision.AxisChecks + Projection.IsOverlapping, SATCollision.cs:39-59,100-104 */
static int axis_checks(const v2* a, int na, const v2* b, int nb, v2* normal, float* depth, int* axis_idx, int idx_base) {
  for (int i = 0; i < na; i++) {
    v2 edge = v_sub(a[(i + 1) % na], a[i]);
    v2 axis = V(-edge.y, edge.x);
    if (axis.x == 0.0f && axis.y == 0.0f) continue;
    axis = v_normalize(axis);
    float amin, amax, bmin, bmax;
    project(axis, a, na, &amin, &amax);
    project(axis, b, nb, &bmin, &bmax);
    float temp = net_min(bmax - amin, amax - bmin);
    int overlapping = (amin < bmax) && (bmin < amax);
    if (!overlapping) return 0;
    if (temp >= *depth) continue;
    *depth = temp;
    *normal = axis;
    *axis_idx = idx_base + i;
  }
  return 1;
}

/* SATCollision.IsColliding, SATCollision.cs:15-35 */
static int sat_colliding(const v2* a, int na, const v2* b, int nb, v2 ca, v2 cb, v2* normal, float* depth, int* axis_idx) {
  *normal = V(0.0f, 0.0f);
  *depth = FLT_MAX;
  *axis_idx = -1;
  int result = axis_checks(a, na, b, nb, normal, depth, axis_idx, 0) && axis_checks(b, nb, a, na, normal, depth, axis_idx, na);
  v2 dir = v_sub(cb, ca);
  if (v_dot(dir, *normal) > 0.0f) *normal = v_mul(*normal, -1.0f);
  return result;
}

/* ------------------------------------------------------------------ contact points ----
 * ContactPoints.cs:116-128 */
typedef struct {
  v2 a, b, max;
} face_t;

/* ContactPoints.Mod, ContactPoints.cs:131-134 (exact for the small ints it is used with) */
static int cp_mod(int a, int b) { return (int)lrint((double)((float)a) - (double)((float)b) * floor((double)((float)a / (float)b))); }

/* ContactPoints.GetSignificantVertex, ContactPoints.cs:97-113 */
static v2 significant_vertex(const v2* p, int n, v2 normal, int* index) {
  v2 sv = V(0.0f, 0.0f);
  *index = -1;
  float min_dist = FLT_MAX;
  for (int i = 0; i < n; i++) {
    float proj = v_dot(p[i], normal);
    if (!(proj < min_dist)) continue;
    sv = p[i];
    *index = i;
    min_dist = proj;
  }
  return sv;
}

/* ContactPoints.GetSignificantFace, ContactPoints.cs:79-94 */
static face_t significant_face(const v2* p, int n, v2 normal) {
  int k;
  v2 sv = significant_vertex(p, n, normal, &k);
  v2 after = v_normalize(v_sub(sv, p[(k + 1) % n]));
  v2 before = v_normalize(v_sub(sv, p[cp_mod(k - 1, n)]));
  face_t f;
  if (v_dot(normal, before) >= v_dot(normal, after)) {
    f.a = sv;
    f.b = p[cp_mod(k - 1, n)];
    f.max = sv;
  } else {
    f.a = p[(k + 1) % n];
    f.b = sv;
    f.max = sv;
  }
  return f;
}

/* ContactPoints.ClipVectors, ContactPoints.cs:56-76 */
static int clip_vectors(v2 a, v2 b, v2 normal, float offset, v2* out) {
  int n = 0;
  float da = v_dot(a, normal) - offset;
  float db = v_dot(b, normal) - offset;
  if (da >= 0.0f) out[n++] = a;
  if (db >= 0.0f) out[n++] = b;
  if (da * db < 0.0f) {
    v2 edge = v_sub(b, a);
    float location = da / (da - db);
    edge = v_mul(edge, location);
    edge = v_add(edge, a);
    out[n++] = edge;
  }
  return n;
}

/* List<Vector2>.Remove(value): removes the first element equal to value */
static int list_remove(v2* pts, int n, v2 value) {
  for (int i = 0; i < n; i++) {
    if (pts[i].x == value.x && pts[i].y == value.y) {
      for (int j = i; j + 1 < n; j++) pts[j] = pts[j + 1];
      return n - 1;
    }
  }
  return n;
}

/* ContactPoints.GetContactPoints, ContactPoints.cs:13-53 */
static int contact_points(const v2* a, int na, const v2* b, int nb, v2 normal, v2* out) {
  face_t ref = significant_face(a, na, normal);
  v2 rf = v_sub(ref.b, ref.a);
  face_t inc = significant_face(b, nb, v_neg(normal));
  v2 ifv = v_sub(inc.b, inc.a);
  if (fabsf(v_dot(rf, normal)) > fabsf(v_dot(ifv, normal))) {
    face_t t = ref;
    ref = inc;
    inc = t;
    rf = v_sub(ref.b, ref.a);
  }
  rf = v_normalize(rf);
  float offset = v_dot(rf, ref.a);
  v2 pts[3];
  int n = clip_vectors(inc.a, inc.b, rf, offset, pts);
  if (n < 2) return 0;
  offset = v_dot(rf, ref.b);
  v2 pts2[3];
  n = clip_vectors(pts[0], pts[1], v_neg(rf), -offset, pts2);
  if (n < 2) return 0;
  v2 ref_normal = V(rf.y, -rf.x);
  float maximum = v_dot(ref_normal, ref.max);
  if (v_dot(ref_normal, pts2[0]) - maximum < 0.0f) n = list_remove(pts2, n, pts2[0]);
  /* clippedPoints.Last() throws on an empty list (n >= 2 at entry, so n >= 1 here) */
  if (v_dot(ref_normal, pts2[n - 1]) - maximum < 0.0f) n = list_remove(pts2, n, pts2[n - 1]);
  for (int i = 0; i < n; i++) out[i] = pts2[i];
  return n;
}

/* ------------------------------------------------------------------ impulses ----
 * Impulses.CalculateImpulse, Impulses.cs:86-115 */
static void calculate_impulse(const body_t* A, const body_t* B, v2 contact, float force, v2 normal, v2* rA, v2* rB, float* impulse) {
  *rA = v_sub(contact, A->centroid);
  v2 perpA = V(-rA->y, rA->x);
  float kA = v_dot(normal, perpA);
  *rB = v_sub(contact, B->centroid);
  v2 perpB = V(-rB->y, rB->x);
  float kB = v_dot(normal, perpB);
  v2 va = v_add(A->vel, v_mul(perpA, A->omega));
  v2 vb = v_add(B->vel, v_mul(perpB, B->omega));
  v2 vrel = v_sub(vb, va);
  float vn = v_dot(vrel, normal);
  float j = -force * vn;
  float denom = (A->inv_mass + B->inv_mass) + ((kA * kA) * A->inv_inertia) + ((kB * kB) * B->inv_inertia);
  j /= denom;
  *impulse = j;
}

/* Impulses.ApplyImpulses, Impulses.cs:57-82 */
static void apply_impulses(body_t* A, body_t* B, v2 normal, float impulse, v2 rA, v2 rB) {
  v2 J = v_mul(normal, impulse); /* impulse * normal: float * Vector2 */
  v2 velA = v_sub(A->vel, v_mul(J, A->inv_mass));
  v2 velB = v_add(B->vel, v_mul(J, B->inv_mass));
  A->vel = velA;
  B->vel = velB;
  v2 perpA = V(-rA.y, rA.x);
  float wA = A->omega - (v_dot(perpA, J) * A->inv_inertia);
  v2 perpB = V(-rB.y, rB.x);
  float wB = B->omega + (v_dot(perpB, J) * B->inv_inertia);
  A->omega = wA;
  B->omega = wB;
}

/* Impulses.ResolveCollisions, Impulses.cs:12-28 */
static void resolve_collisions(body_t* A, body_t* B, const v2* pts, int n, v2 normal) {
  if (n == 0) return;
  float restitution = net_max(A->restitution, B->restitution);
  float friction = net_min(A->friction, B->friction);
  v2 contact = (n == 2) ? v_div(v_add(pts[0], pts[1]), 2.0f) : pts[0];
  v2 rA, rB, rAf, rBf;
  float j, jf;
  calculate_impulse(A, B, contact, (1.0f + restitution), normal, &rA, &rB, &j);
  v2 tangent = V(-normal.y, normal.x);
  calculate_impulse(A, B, contact, friction, tangent, &rAf, &rBf, &jf);
  apply_impulses(A, B, normal, j, rA, rB);
  apply_impulses(A, B, tangent, jf, rAf, rBf);
}

/* Impulses.ResolveJoint, Impulses.cs:31-40 */
static void resolve_joint(body_t* A, body_t* B, v2 p0, v2 p1, v2 normal) {
  v2 contact = v_div(v_add(p0, p1), 2.0f);
  v2 rA, rB;
  float j;
  calculate_impulse(A, B, contact, (1.0f + 1.0f), normal, &rA, &rB, &j);
  apply_impulses(A, B, normal, j, rA, rB);
}

/* Joint.Step, Joint.cs:31-41 */
static void joint_step(struct wo_env* e, int k, wo_joint_trace* tr) {
  body_t* A = &e->body[e->joint_a[k]];
  body_t* B = &e->body[e->joint_b[k]];
  v2 ab = v_sub(B->verts[e->joint_ib[k]], A->verts[e->joint_ia[k]]);
  float depth = v_length(ab);
  if (tr) {
    tr->depth = depth;
    tr->active = !(depth < 0.1f);
  }
  if (depth < 0.1f) return;
  ab = v_normalize(ab);
  sk_move(A, v_div(v_mul(ab, depth), 2.0f));
  sk_move(B, v_div(v_mul(v_neg(ab), depth), 2.0f));
  resolve_joint(B, A, A->verts[e->joint_ia[k]], B->verts[e->joint_ib[k]], ab);
}

/* RigidBody.MoveObjects, RigidBody.cs:99-113 */
static void move_objects(body_t* A, body_t* B, v2 normal, float depth) {
  if (A->is_static) {
    sk_move(B, v_mul(v_neg(normal), depth));
  } else if (B->is_static) {
    sk_move(A, v_mul(normal, depth));
  } else {
    sk_move(A, v_div(v_mul(normal, depth), 2.0f));
    sk_move(B, v_div(v_mul(v_neg(normal), depth), 2.0f));
  }
}

static const int k_slot_base[5] = {0, 2, 4, 5, 7};

/* RigidBody.ResolveCollisions, RigidBody.cs:66-96 */
static void rb_resolve_collisions(struct wo_env* e, int id, wo_pair_trace* tr) {
  body_t* self = &e->body[id];
  int cand = 0;
  for (int oi = 0; oi < e->nbodies; oi++) {
    int oid = e->order[oi];
    if (oid == id) continue;
    if (self->assoc & (1u << oid)) continue;
    body_t* other = &e->body[oid];
    wo_pair_trace* rec = NULL;
    if (tr && id < 5) {
      rec = &tr[k_slot_base[id] + cand];
      memset(rec, 0, sizeof(*rec));
      rec->other = oid;
      rec->axis = -1;
    }
    cand++;
    if (!bb_colliding(self, other)) continue;
    if (rec) rec->aabb = 1;
    if (other->is_floor) self->collided = 1;
    if (self->is_floor) other->collided = 1;
    v2 normal;
    float depth;
    int axis;
    if (sat_colliding(self->verts, self->n, other->verts, other->n, self->centroid, other->centroid, &normal, &depth, &axis)) {
      v2 pts[3];
      int n = contact_points(self->verts, self->n, other->verts, other->n, normal, pts);
      if (rec) {
        rec->sat = 1;
        rec->axis = axis;
        rec->nx = normal.x;
        rec->ny = normal.y;
        rec->depth = depth;
        rec->ncontacts = n;
        if (n > 0) {
          rec->c0x = pts[0].x;
          rec->c0y = pts[0].y;
        }
        if (n > 1) {
          rec->c1x = pts[1].x;
          rec->c1y = pts[1].y;
        }
      }
      move_objects(self, other, normal, depth);
      resolve_collisions(self, other, pts, n, normal);
    }
  }
}

/* RigidBody.Step + StepLinearVelocity + StepAngularVelocity + WrapAngle, RigidBody.cs:54-61,116-140 */
static void rb_step(struct wo_env* e, int id, float dt, wo_pair_trace* tr) {
  body_t* b = &e->body[id];
  b->vel = v_add(b->vel, v_mul(b->accel, dt));
  sk_move(b, v_mul(b->vel, dt));
  if (b->is_static) return;
  const float PI_F = 3.14159274f, TAU_F = 6.28318548f; /* MathF.PI, MathF.Tau */
  float a = b->angle + b->omega * dt;
  if (a > PI_F)
    a = a - TAU_F;
  else if (a < -PI_F)
    a = a + TAU_F;
  b->angle = a;
  sk_rotate(b, b->omega * dt);
  rb_resolve_collisions(e, id, tr);
}

/* Environment.StepObjects, Environment.cs:126-143 */
void wo_env_step_objects(wo_env* e, float dt, int iterations, wo_pair_trace* pair_trace, wo_joint_trace* joint_trace) {
  dt /= (float)iterations;
  for (int it = 0; it < iterations; it++) {
    wo_pair_trace* ptr = pair_trace ? pair_trace + (size_t)it * WO_PAIR_SLOTS : NULL;
    if (ptr) {
      memset(ptr, 0, sizeof(wo_pair_trace) * WO_PAIR_SLOTS);
      for (int s = 0; s < WO_PAIR_SLOTS; s++) {
        ptr[s].other = -1;
        ptr[s].axis = -1;
      }
    }
    for (int k = 0; k < 4; k++) joint_step(e, k, joint_trace ? joint_trace + (size_t)it * 4 + k : NULL);
    for (int oi = 0; oi < e->nbodies; oi++) rb_step(e, e->order[oi], dt, ptr);
  }
}

/* ------------------------------------------------------------------ walker / environment ----
 * Pole.FromSize, Pole.cs:18-34 */
static void pole_from_size(body_t* b, wo_material m, v2 c, float size) {
  float adjustment = (float)0.1 * size;
  v2 v[6] = {
      V(c.x + adjustment, c.y + adjustment * 3.5f), V(c.x, c.y + adjustment * 3.5f),
      V(c.x - adjustment, c.y + adjustment * 3.5f), V(c.x - adjustment, c.y - adjustment * 3.5f),
      V(c.x, c.y - adjustment * 3.5f),              V(c.x + adjustment, c.y - adjustment * 3.5f),
  };
  rb_init(b, m, 0, 0);
  sk_add_vectors(b, v, 6);
}

void wo_pole_from_size(float cx, float cy, float size, float* verts12, float* centroid2) {
  body_t b;
  wo_material m = {1.0f, 0.0f, 0.0f};
  pole_from_size(&b, m, V(cx, cy), size);
  for (int i = 0; i < 6; i++) {
    verts12[2 * i] = b.verts[i].x;
    verts12[2 * i + 1] = b.verts[i].y;
  }
  centroid2[0] = b.centroid.x;
  centroid2[1] = b.centroid.y;
}

/* Walker.CreateCreature: CreateBodies/CreateJoints/AddAssociatedBodies/AddAcceleration, Walker.cs:40-46,155-209 */
static void create_creature(struct wo_env* e) {
  v2 pos = e->pos;
  wo_material m = e->walker_mat;
  body_t* body = &e->body[WO_BODY];
  v2 hull[5] = {V(pos.x + 20.0f, pos.y + 20.0f), V(pos.x, pos.y + 20.0f), V(pos.x - 20.0f, pos.y + 20.0f),
                V(pos.x - 20.0f, pos.y - 20.0f), V(pos.x + 20.0f, pos.y - 20.0f)};
  rb_init(body, m, 0, 0);
  sk_add_vectors(body, hull, 5);
  body->inv_inertia = 0.0003f;
  pole_from_size(&e->body[WO_LLU], m, v_add(pos, V(0.0f, 30.0f)), 75.0f);
  pole_from_size(&e->body[WO_LLL], m, v_add(pos, V(0.0f, 60.0f)), 75.0f);
  pole_from_size(&e->body[WO_RLU], m, v_add(pos, V(0.0f, 30.0f)), 75.0f);
  pole_from_size(&e->body[WO_RLL], m, v_add(pos, V(0.0f, 60.0f)), 75.0f);
  /* rigidBodies.AddRange({LLL, LLU, Body, RLL, RLU}), Walker.cs:176 */
  static const int add_order[5] = {WO_LLL, WO_LLU, WO_BODY, WO_RLL, WO_RLU};
  for (int i = 0; i < 5; i++) e->order[e->nbodies++] = add_order[i];
  /* joints, Walker.cs:182-187 */
  static const int ja[4] = {WO_BODY, WO_BODY, WO_LLU, WO_RLU};
  static const int jb[4] = {WO_LLU, WO_RLU, WO_LLL, WO_RLL};
  static const int jia[4] = {1, 1, 2, 2};
  static const int jib[4] = {4, 4, 3, 3};
  for (int k = 0; k < 4; k++) {
    e->joint_a[k] = ja[k];
    e->joint_b[k] = jb[k];
    e->joint_ia[k] = jia[k];
    e->joint_ib[k] = jib[k];
    e->joint_torque[k] = 0.0f;
  }
  /* association ("no collide") lists, Walker.cs:204-208 */
  e->body[WO_LLU].assoc = (1u << WO_RLU) | (1u << WO_RLL) | (1u << WO_BODY);
  e->body[WO_LLL].assoc = (1u << WO_RLU) | (1u << WO_RLL) | (1u << WO_BODY);
  e->body[WO_RLU].assoc = (1u << WO_LLU) | (1u << WO_LLL) | (1u << WO_BODY);
  e->body[WO_RLL].assoc = (1u << WO_LLU) | (1u << WO_LLL) | (1u << WO_BODY);
  e->body[WO_BODY].assoc = (1u << WO_LLU) | (1u << WO_RLU) | (1u << WO_LLL) | (1u << WO_RLL);
  /* gravity, Walker.cs:45,191-198 */
  for (int i = 0; i < 5; i++) e->body[i].accel = v_add(e->body[i].accel, V(0.0f, 980.0f));
}

/* Environment.CreateFloor, Environment.cs:211-226 */
static void create_floor(struct wo_env* e) {
  v2 fp[4] = {V(-50.0f, 1050.0f), V(-50.0f, 900.0f), V(1050.0f, 900.0f), V(1050.0f, 1050.0f)};
  body_t* f = &e->body[WO_FLOOR];
  rb_init(f, e->floor_mat, 1, 1);
  sk_add_vectors(f, fp, 4);
  e->order[e->nbodies++] = WO_FLOOR;
}

/* Walker.Update, Walker.cs:49-54 */
static void walker_update(struct wo_env* e) {
  e->prev_pos = e->pos;
  e->pos = e->body[WO_BODY].centroid;
  if (e->body[WO_BODY].collided || e->body[WO_LLU].collided || e->body[WO_RLU].collided) e->terminal = 1;
}

/* Environment ctor + InitialState, Environment.cs:39-51,176-180 */
void wo_env_init(wo_env* e, wo_material floor, wo_material walker) {
  memset(e, 0, sizeof(*e));
  e->floor_mat = floor;
  e->walker_mat = walker;
  e->pos = V(125.0f, 800.0f); /* Walker ctor, Walker.cs:31-32 */
  e->prev_pos = e->pos;
  e->terminal = 0;
  e->nbodies = 0;
  create_creature(e);
  create_floor(e);
  e->floor_first = 0;
  e->steps = 0;
  walker_update(e);
}

/* Environment.Reset + Walker.Reset + InitialState, Environment.cs:167-180, Walker.cs:212-236 */
void wo_env_reset(wo_env* e) {
  e->steps = 0;
  /* RemoveRigidObjects: drop the five walker bodies, floor stays at the head */
  e->order[0] = WO_FLOOR;
  e->nbodies = 1;
  e->terminal = 0;
  e->pos = V(125.0f, 800.0f);
  e->prev_pos = e->pos;
  create_creature(e);
  e->floor_first = 1;
  walker_update(e);
}

/* Matrix.Clip (Matrix.cs:377-405) + Walker.TakeActions (Walker.cs:66-75) + Joint.SetTorque (Joint.cs:56-61) */
void wo_env_take_actions(wo_env* e, const float* actions4) {
  for (int k = 0; k < 4; k++) {
    float a = actions4[k];
    if (a >= 1.0f)
      a = 1.0f;
    else if (a <= -1.0f)
      a = -1.0f;
    float change = a - e->joint_torque[k];
    e->joint_torque[k] = a;
    body_t* B = &e->body[e->joint_b[k]];
    B->omega += change * 5.0f;
  }
}

/* Walker.GetState, Walker.cs:132-152 (Game1.FrameRate = 60, Game1.cs:15) */
void wo_env_get_obs(const wo_env* e, float* o) {
  v2 j0 = e->body[e->joint_a[0]].verts[e->joint_ia[0]];
  v2 j2 = e->body[e->joint_a[2]].verts[e->joint_ia[2]];
  v2 j3 = e->body[e->joint_a[3]].verts[e->joint_ia[3]];
  o[0] = j0.x / 900.0f;
  o[1] = j0.y / 500.0f;
  o[2] = j2.x / 900.0f;
  o[3] = j2.y / 500.0f;
  o[4] = j3.x / 900.0f;
  o[5] = j3.y / 500.0f;
  o[6] = e->body[WO_BODY].vel.x / 60.0f;
  o[7] = e->body[WO_BODY].vel.y / 60.0f;
  o[8] = e->body[WO_LLL].angle;
  o[9] = e->body[WO_LLU].angle;
  o[10] = e->body[WO_RLL].angle;
  o[11] = e->body[WO_RLU].angle;
}

/* tail of Environment.Step, Environment.cs:101-121, and CalculateReward, :148-154 */
void wo_env_observe(wo_env* e, int max_timesteps, float* obs12, float* reward, uint8_t* done) {
  walker_update(e);
  float dx = e->pos.x - e->prev_pos.x;
  float h = e->body[e->joint_a[0]].verts[e->joint_ia[0]].y / 500.0f;
  float r = 0.0f;
  r += (dx > 0.0f && h < 1.6f) ? dx : 0.0f;
  r -= (h > 1.65f) ? -0.1f : 0.0f;
  int terminal = 0;
  if (e->terminal || e->steps > max_timesteps) {
    if (e->terminal) r -= 40.0f;
    terminal = 1;
  }
  if (e->pos.x > 900.0f) {
    r += 80.0f;
    terminal = 1;
  }
  *reward = r;
  *done = (uint8_t)terminal;
  wo_env_get_obs(e, obs12);
}

/* Environment.Update minus the policy, Environment.cs:64-92 */
void wo_env_step(wo_env* e, const float* actions4, float dt, int iterations, int max_timesteps, int auto_reset, float* obs12,
                 float* reward, uint8_t* done) {
  e->steps++;
  wo_env_take_actions(e, actions4);
  wo_env_step_objects(e, dt, iterations, NULL, NULL);
  wo_env_observe(e, max_timesteps, obs12, reward, done);
  if (*done && auto_reset) {
    wo_env_reset(e);
    wo_env_get_obs(e, obs12);
  }
}

void wo_env_get_state(const wo_env* e, float* f, int32_t* iv) {
  static const int ids[5] = {WO_LLL, WO_LLU, WO_BODY, WO_RLL, WO_RLU};
  int p = 0;
  for (int b = 0; b < 5; b++)
    for (int i = 0; i < e->body[ids[b]].n; i++) {
      f[p++] = e->body[ids[b]].verts[i].x;
      f[p++] = e->body[ids[b]].verts[i].y;
    }
  for (int b = 0; b < 5; b++) {
    f[p++] = e->body[b].centroid.x;
    f[p++] = e->body[b].centroid.y;
  }
  for (int b = 0; b < 5; b++) {
    f[p++] = e->body[b].vel.x;
    f[p++] = e->body[b].vel.y;
  }
  for (int b = 0; b < 5; b++) f[p++] = e->body[b].omega;
  for (int b = 0; b < 5; b++) f[p++] = e->body[b].angle;
  for (int k = 0; k < 4; k++) f[p++] = e->joint_torque[k];
  int flags = 0;
  for (int b = 0; b < 5; b++)
    if (e->body[b].collided) flags |= (1 << b);
  if (e->terminal) flags |= WO_FLAG_TERMINAL;
  if (e->floor_first) flags |= WO_FLAG_FLOOR_FIRST;
  iv[0] = flags;
  iv[1] = e->steps;
}

void wo_env_set_state(wo_env* e, const float* f, const int32_t* iv) {
  int p = 0;
  for (int b = 0; b < 5; b++) {
    for (int i = 0; i < e->body[b].n; i++) {
      e->body[b].verts[i].x = f[p++];
      e->body[b].verts[i].y = f[p++];
    }
    bb_update(&e->body[b]);
  }
  for (int b = 0; b < 5; b++) {
    e->body[b].centroid.x = f[p++];
    e->body[b].centroid.y = f[p++];
  }
  for (int b = 0; b < 5; b++) {
    e->body[b].vel.x = f[p++];
    e->body[b].vel.y = f[p++];
  }
  for (int b = 0; b < 5; b++) e->body[b].omega = f[p++];
  for (int b = 0; b < 5; b++) e->body[b].angle = f[p++];
  for (int k = 0; k < 4; k++) e->joint_torque[k] = f[p++];
  int flags = iv[0];
  for (int b = 0; b < 5; b++) e->body[b].collided = (flags >> b) & 1;
  e->terminal = (flags & WO_FLAG_TERMINAL) ? 1 : 0;
  e->floor_first = (flags & WO_FLAG_FLOOR_FIRST) ? 1 : 0;
  e->steps = iv[1];
  if (e->floor_first) {
    static const int o[6] = {WO_FLOOR, WO_LLL, WO_LLU, WO_BODY, WO_RLL, WO_RLU};
    memcpy(e->order, o, sizeof(o));
  } else {
    static const int o[6] = {WO_LLL, WO_LLU, WO_BODY, WO_RLL, WO_RLU, WO_FLOOR};
    memcpy(e->order, o, sizeof(o));
  }
  e->nbodies = 6;
  /* Walker._position always equals the Body centroid at a step boundary (Walker.cs:52) */
  e->pos = e->body[WO_BODY].centroid;
  e->prev_pos = e->pos;
}

void wo_batch_step(void* envs, int n, const float* actions, float dt, int iterations, int max_timesteps, int auto_reset,
                   float* obs, float* reward, uint8_t* done, int nthreads) {
  struct wo_env* E = (struct wo_env*)envs;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int i = 0; i < n; i++)
    wo_env_step(&E[i], actions + (size_t)i * 4, dt, iterations, max_timesteps, auto_reset, obs + (size_t)i * 12, reward + i,
                done + i);
  (void)nthreads;
}

/* ---- stand-alone wrappers for unit tests ---- */
int wo_sat(const float* a, int na, const float* b, int nb, const float* ca, const float* cb, float* normal2, float* depth,
           int* axis) {
  v2 A[MAXV], B[MAXV];
  for (int i = 0; i < na; i++) A[i] = V(a[2 * i], a[2 * i + 1]);
  for (int i = 0; i < nb; i++) B[i] = V(b[2 * i], b[2 * i + 1]);
  v2 n;
  int r = sat_colliding(A, na, B, nb, V(ca[0], ca[1]), V(cb[0], cb[1]), &n, depth, axis);
  normal2[0] = n.x;
  normal2[1] = n.y;
  return r;
}

int wo_contacts(const float* a, int na, const float* b, int nb, const float* normal2, float* pts4) {
  v2 A[MAXV], B[MAXV], out[3];
  for (int i = 0; i < na; i++) A[i] = V(a[2 * i], a[2 * i + 1]);
  for (int i = 0; i < nb; i++) B[i] = V(b[2 * i], b[2 * i + 1]);
  int n = contact_points(A, na, B, nb, V(normal2[0], normal2[1]), out);
  for (int i = 0; i < n; i++) {
    pts4[2 * i] = out[i].x;
    pts4[2 * i + 1] = out[i].y;
  }
  return n;
}
