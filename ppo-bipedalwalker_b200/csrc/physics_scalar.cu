// physics_scalar.cu -- lockstep rigid-body step, ONE THREAD PER ENVIRONMENT (sm_100a).
//
// Same path and the same arithmetic contract as physics.cu (Environment.StepObjects and everything under it:
// Environment.cs:126-143; Joint.cs:31-61; RigidBody.cs:54-140; Skeleton.cs:76-176; SATCollision.cs:15-104;
// ContactPoints.cs:13-134; Impulses.cs:12-115; walker glue Walker.cs:49-75,132-152, Environment.cs:96-122,148-180),
// but mapped for THROUGHPUT: the reference's step is a strictly sequential Gauss-Seidel sweep per environment, so the
// only unlimited parallelism is across environments.  Here every lane of a warp advances a different walker:
//   * no lane does redundant scalar work, no shuffles, no intra-env synchronisation at all;
//   * per-env state lives in shared memory as columns, slot-major / env-minor ([slot][128 envs]): every access, including
//     the data-dependent ones (support vertex k, its neighbours, runtime body ids), is bank-conflict free, and one copy of
//     each helper serves all bodies (runtime slot arithmetic), which keeps the SASS small;
//   * the polygons of the pair under test are pulled into registers once and feed AABB, all SAT axes and the move.
// The cooperative 16-lanes-per-env kernel in physics.cu needs ~2950 warp-instructions per env-substep; this one ~330.
//
// Bit-exactness: IEEE binary32, each multiply/add individually rounded (never FMA), correctly rounded 1/x, sqrt, division,
// (float)cos/sin((double)theta) for Skeleton.Rotate, reference operation order.  No CPU path.
#include "physics.cuh"
#include "physics_math.cuh"

namespace wb {
namespace t1 {

constexpr int kT = kScalarEnvsPerCta;  // environments (= threads) per CTA

__constant__ Material c_materials[WB_MAX_MATERIALS];
__constant__ float c_init_state[kStateFloats];  // state record of a freshly created walker
__constant__ FloorConst c_floor;

enum { LLL = 0, LLU = 1, BODY = 2, RLL = 3, RLU = 4, FLOOR = 5 };

// ---------------------------------------------------------------- shared-memory columns of one env
// float2 slots: vertices body*6 + i (Body's 6th slot mirrors its vertex 0, see below), centroids, velocities
constexpr int kV2Cen = 30;
constexpr int kV2Vel = 35;
constexpr int kV2Count = 40;
// float slots
constexpr int kFOmega = 0;
constexpr int kFAngle = 5;
constexpr int kFCount = 10;

// The Body hull has 5 vertices; its 6th slot holds a COPY of vertex 0 that goes through the same Move/Rotate operations
// (so it stays bit-identical to vertex 0).  Every polygon is then a 6-gon for the kernels: the duplicate never changes a
// min/max projection, never wins the strict "first smallest" support-vertex scan, edge 4 (P5 - P4) equals the hull's closing
// edge P0 - P4, and edge 5 (P0 - P5) is the zero vector, which SATCollision.AxisChecks itself skips (SATCollision.cs:45).
__device__ __forceinline__ int nverts(int b) { return b == BODY ? 5 : 6; }

struct Env {
  float2* v2;  // this env's column of float2 slots: slot s at v2[s * kT]
  float* f;    // this env's column of float slots
  int flags;   // Collided bits, Terminal, floor-first
  // material-derived constants (RigidBody ctor, RigidBody.cs:36-50; Impulses.cs:16-17)
  float im_w;     // walker inverse mass
  float ii_pole;  // 0.001f * inverse mass
  float e_ww, mu_ww, e_wf, mu_wf;
};

__device__ __forceinline__ float2& V2(const Env& e, int slot) { return e.v2[slot * kT]; }
__device__ __forceinline__ float& F1(const Env& e, int slot) { return e.f[slot * kT]; }
__device__ __forceinline__ float inv_inertia(const Env& e, int b) { return b == BODY ? 0.0003f : e.ii_pole; }  // Walker.cs:168

struct BodyDyn {
  float2 c, v;
  float w, im, ii;
};

__device__ __forceinline__ BodyDyn load_dyn(const Env& e, int b) {
  BodyDyn d;
  d.c = V2(e, kV2Cen + b);
  d.v = V2(e, kV2Vel + b);
  d.w = F1(e, kFOmega + b);
  d.im = e.im_w;
  d.ii = inv_inertia(e, b);
  return d;
}
__device__ __forceinline__ BodyDyn floor_dyn() {
  BodyDyn d;
  d.c = c_floor.cen;
  d.v = mk2(0.0f, 0.0f);
  d.w = 0.0f;
  d.im = 0.0f;
  d.ii = 0.0f;
  return d;
}
__device__ __forceinline__ void store_dyn(const Env& e, int b, const BodyDyn& d) {
  V2(e, kV2Vel + b) = d.v;
  F1(e, kFOmega + b) = d.w;
}

// Impulses.CalculateImpulse, Impulses.cs:86-115
__device__ __forceinline__ void calculate_impulse(const BodyDyn& A, const BodyDyn& B, float2 contact, float force, float2 n,
                                                  float2& rA, float2& rB, float& impulse) {
  rA = vsub(contact, A.c);
  const float2 perpA = mk2(-rA.y, rA.x);
  const float kA = vdot(n, perpA);
  rB = vsub(contact, B.c);
  const float2 perpB = mk2(-rB.y, rB.x);
  const float kB = vdot(n, perpB);
  const float2 va = vadd(A.v, vmul(perpA, A.w));
  const float2 vb = vadd(B.v, vmul(perpB, B.w));
  const float2 vrel = vsub(vb, va);
  const float vn = vdot(vrel, n);
  const float j = fmul(-force, vn);
  const float denom = fadd(fadd(fadd(A.im, B.im), fmul(fmul(kA, kA), A.ii)), fmul(fmul(kB, kB), B.ii));
  impulse = fdiv(j, denom);
}

// Impulses.ApplyImpulses, Impulses.cs:57-82
__device__ __forceinline__ void apply_impulses(BodyDyn& A, BodyDyn& B, float2 n, float impulse, float2 rA, float2 rB) {
  const float2 J = vmul(n, impulse);
  const float2 velA = vsub(A.v, vmul(J, A.im));
  const float2 velB = vadd(B.v, vmul(J, B.im));
  const float2 perpA = mk2(-rA.y, rA.x);
  const float wA = fsub(A.w, fmul(vdot(perpA, J), A.ii));
  const float2 perpB = mk2(-rB.y, rB.x);
  const float wB = fadd(B.w, fmul(vdot(perpB, J), B.ii));
  A.v = velA;
  B.v = velB;
  A.w = wA;
  B.w = wB;
}

// Skeleton.Move, Skeleton.cs:76-85 (all six slots: the Body's mirror slot moves with its vertex 0)
__device__ __forceinline__ void move_body(const Env& e, int b, float2 d) {
#pragma unroll
  for (int i = 0; i < 6; i++) V2(e, b * 6 + i) = vadd(V2(e, b * 6 + i), d);
  V2(e, kV2Cen + b) = vadd(V2(e, kV2Cen + b), d);
}
__device__ __forceinline__ void move_body_regs(const Env& e, int b, const float2 (&P)[6], float2 d) {
#pragma unroll
  for (int i = 0; i < 6; i++) V2(e, b * 6 + i) = vadd(P[i], d);
  V2(e, kV2Cen + b) = vadd(V2(e, kV2Cen + b), d);
}

// ---------------------------------------------------------------- Joint.Step, Joint.cs:31-41
template <bool TRACE>
__device__ __forceinline__ void joint_step(const Env& e, int A, int ia, int B, int ib, wb_joint_trace* tr) {
  const float2 pA = V2(e, A * 6 + ia);
  const float2 pB = V2(e, B * 6 + ib);
  float2 ab = vsub(pB, pA);
  const float depth = fsqrt(fadd(fmul(ab.x, ab.x), fmul(ab.y, ab.y)));  // Vector2.Length
  if (TRACE) {
    if (tr) {
      tr->active = !(depth < 0.1f);
      tr->depth = depth;
    }
  }
  if (depth < 0.1f) return;
  ab = vnormalize_fast(ab);
  const float2 dA = vhalf(vmul(ab, depth));
  const float2 dB = vhalf(vmul(vneg(ab), depth));
  move_body(e, A, dA);
  move_body(e, B, dB);
  BodyDyn X = load_dyn(e, B);  // Manifold(bodyA := joint._bodyB, bodyB := joint._bodyA), Joint.cs:40
  BodyDyn Y = load_dyn(e, A);
  const float2 contact = vhalf(vadd(vadd(pA, dA), vadd(pB, dB)));  // Vector2.Divide(p0 + p1, 2), Impulses.cs:35
  float2 rX, rY;
  float j;
  calculate_impulse(X, Y, contact, fadd(1.0f, 1.0f), ab, rX, rY, j);
  apply_impulses(X, Y, ab, j, rX, rY);
  store_dyn(e, B, X);
  store_dyn(e, A, Y);
}

// ---------------------------------------------------------------- SAT, SATCollision.cs:15-104
// float.MaxValue / float.MinValue seeds take part in the min/max exactly like the reference's "if (t < min) min = t"
__device__ __forceinline__ void project6(const float2 (&P)[6], float2 ax, float& mn, float& mx) {
  const float t0 = vdot(ax, P[0]), t1 = vdot(ax, P[1]), t2 = vdot(ax, P[2]);
  const float t3 = vdot(ax, P[3]), t4 = vdot(ax, P[4]), t5 = vdot(ax, P[5]);
  mn = fminf(fminf(fminf(FLT_MAX, t0), fminf(t1, t2)), fminf(fminf(t3, t4), t5));
  mx = fmaxf(fmaxf(fmaxf(-FLT_MAX, t0), fmaxf(t1, t2)), fmaxf(fmaxf(t3, t4), t5));
}
__device__ __forceinline__ void project_floor(float2 ax, float& mn, float& mx) {
  const float t0 = vdot(ax, c_floor.v[0]), t1 = vdot(ax, c_floor.v[1]);
  const float t2 = vdot(ax, c_floor.v[2]), t3 = vdot(ax, c_floor.v[3]);
  mn = fminf(fminf(fminf(FLT_MAX, t0), t1), fminf(t2, t3));
  mx = fmaxf(fmaxf(fmaxf(-FLT_MAX, t0), t1), fmaxf(t2, t3));
}

struct Sat {
  float depth;
  float2 normal;
  int idx;
  bool sep;  // an axis separated the polygons: AxisChecks returned false (every later axis is ignored)
};

// one iteration of the AxisChecks loop body after the projections (SATCollision.cs:47-56); symmetric in the two projections.
// Once an axis separates, the reference returns false and nothing computed afterwards is used, so only `sep` needs guarding;
// a skipped (zero) axis has NaN projections and must not touch anything.  "tempDepth >= depth -> continue" keeps the FIRST
// minimal axis: update only on a strictly smaller depth (temp is never NaN on a live, overlapping axis).
__device__ __forceinline__ void sat_accumulate(Sat& s, bool skip, float2 axis, int idx, float omin, float omax, float tmin, float tmax) {
  const float temp = fminf(fsub(tmax, omin), fsub(omax, tmin));
  const bool overlapping = (omin < tmax) && (tmin < omax);
  s.sep = s.sep || (!skip && !overlapping);
  if (!skip && temp < s.depth) {
    s.depth = temp;
    s.normal = axis;
    s.idx = idx;
  }
}

// left normal of an edge, zero test, Vector2.Normalize (SATCollision.cs:43-46)
__device__ __forceinline__ float2 edge_axis(float2 p0, float2 p1, bool& skip) {
  const float2 edge = vsub(p1, p0);
  const float2 axis = mk2(-edge.y, edge.x);
  skip = (axis.x == 0.0f) && (axis.y == 0.0f);
  return vnormalize_fast(axis);
}

// ---------------------------------------------------------------- contact points, ContactPoints.cs:13-134
struct Face {
  float2 a, b, max;
};

// GetSignificantFace given the support vertex sv = P[k] and its neighbours (ContactPoints.cs:79-94)
__device__ __forceinline__ Face face_from(float2 sv, float2 next, float2 prev, float2 nrm) {
  const float2 after = vnormalize_fast(vsub(sv, next));
  const float2 before = vnormalize_fast(vsub(sv, prev));
  const bool use_before = vdot(nrm, before) >= vdot(nrm, after);
  Face f;
  f.a = use_before ? sv : next;
  f.b = use_before ? prev : sv;
  f.max = sv;
  return f;
}

// GetSignificantVertex (ContactPoints.cs:97-113): first index with the strictly smallest projection
__device__ __forceinline__ int support_index6(const float2 (&P)[6], float2 nrm) {
  float best = FLT_MAX;
  int k = 0;  // (k stays 0 only if no projection is below float.MaxValue: non-finite state)
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float pr = vdot(P[i], nrm);
    const bool lt = pr < best;
    k = lt ? i : k;
    best = lt ? pr : best;
  }
  return k;
}

__device__ __forceinline__ Face significant_face_body(const Env& e, int b, const float2 (&P)[6], float2 nrm) {
  const int n = nverts(b);
  const int k = support_index6(P, nrm);
  const int kn = (k + 1 == n) ? 0 : k + 1;
  const int kp = (k == 0) ? n - 1 : k - 1;
  return face_from(V2(e, b * 6 + k), V2(e, b * 6 + kn), V2(e, b * 6 + kp), nrm);
}

__device__ __forceinline__ Face significant_face_floor(float2 nrm) {
  float best = FLT_MAX;
  int k = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float pr = vdot(c_floor.v[i], nrm);
    const bool lt = pr < best;
    k = lt ? i : k;
    best = lt ? pr : best;
  }
  return face_from(c_floor.v[k], c_floor.v[(k + 1) & 3], c_floor.v[(k + 3) & 3], nrm);
}

// ClipVectors, ContactPoints.cs:56-76: appends up to 3 points; only the first two are ever used
__device__ __forceinline__ int clip_vectors(float2 a, float2 b, float2 nrm, float offset, float2& o0, float2& o1) {
  int cnt = 0;
  const float da = fsub(vdot(a, nrm), offset);
  const float db = fsub(vdot(b, nrm), offset);
  if (da >= 0.0f) {
    o0 = a;
    cnt = 1;
  }
  if (db >= 0.0f) {
    if (cnt == 0) o0 = b; else o1 = b;
    cnt++;
  }
  if (fmul(da, db) < 0.0f) {
    float2 ed = vsub(b, a);
    const float location = fdiv(da, fsub(da, db));
    ed = vmul(ed, location);
    ed = vadd(ed, a);
    if (cnt == 0) o0 = ed; else if (cnt == 1) o1 = ed;
    cnt++;
  }
  return cnt;
}

__device__ __forceinline__ bool veq(float2 a, float2 b) { return a.x == b.x && a.y == b.y; }

// GetContactPoints after the two significant faces are known, ContactPoints.cs:13-53
__device__ __forceinline__ int contact_points(Face ref, Face inc, float2 normal, float2& c0, float2& c1) {
  float2 rf = vsub(ref.b, ref.a);
  const float2 ifv = vsub(inc.b, inc.a);
  if (fabsf(vdot(rf, normal)) > fabsf(vdot(ifv, normal))) {
    const Face t = ref;
    ref = inc;
    inc = t;
    rf = vsub(ref.b, ref.a);
  }
  rf = vnormalize_fast(rf);
  float offset = vdot(rf, ref.a);
  float2 p0 = mk2(0.f, 0.f), p1 = mk2(0.f, 0.f);
  int cnt = clip_vectors(inc.a, inc.b, rf, offset, p0, p1);
  if (cnt < 2) return 0;
  offset = vdot(rf, ref.b);
  float2 q0 = mk2(0.f, 0.f), q1 = mk2(0.f, 0.f);
  cnt = clip_vectors(p0, p1, vneg(rf), -offset, q0, q1);
  if (cnt < 2) return 0;
  cnt = 2;  // ClipVectors returns 3 points only if da >= 0, db >= 0 and da*db < 0 at once, which is impossible
  const float2 rn = mk2(rf.y, -rf.x);
  const float maximum = vdot(rn, ref.max);
  // List.Remove(First()) then List.Remove(Last()): Remove deletes the first element EQUAL to the value
  if (fsub(vdot(rn, q0), maximum) < 0.0f) {
    q0 = q1;
    cnt = 1;
  }
  const float2 last = (cnt == 2) ? q1 : q0;
  if (fsub(vdot(rn, last), maximum) < 0.0f) {
    if (cnt == 2) {
      if (veq(q0, q1)) q0 = q1;  // removes index 0 when both points are equal (same value survives)
      cnt = 1;
    } else {
      cnt = 0;
    }
  }
  c0 = q0;
  c1 = q1;
  return cnt;
}

__device__ __forceinline__ void trace_init(wb_pair_trace& rec, int other) {
  rec.other = other;
  rec.aabb = 0;
  rec.sat = 0;
  rec.axis = -1;
  rec.nx = rec.ny = rec.depth = 0.0f;
  rec.ncontacts = 0;
  rec.c0x = rec.c0y = rec.c1x = rec.c1y = 0.0f;
}
__device__ __forceinline__ void trace_hit(wb_pair_trace& rec, const Sat& s, float2 normal, int ncp, float2 c0, float2 c1) {
  rec.sat = 1;
  rec.axis = s.idx;
  rec.nx = normal.x;
  rec.ny = normal.y;
  rec.depth = s.depth;
  rec.ncontacts = ncp;
  if (ncp > 0) {
    rec.c0x = c0.x;
    rec.c0y = c0.y;
  }
  if (ncp > 1) {
    rec.c1x = c1.x;
    rec.c1y = c1.y;
  }
}

// Impulses.ResolveCollisions, Impulses.cs:12-28: both impulses come from the PRE-impulse velocities (:23-24), then both are applied (:26-27)
__device__ __forceinline__ void resolve_impulses(BodyDyn& X, BodyDyn& Y, int ncp, float2 c0, float2 c1, float2 normal, float e, float mu) {
  const float2 contact = (ncp == 2) ? vhalf(vadd(c0, c1)) : c0;
  const float2 tangent = mk2(-normal.y, normal.x);
  float2 rA, rB, rAf, rBf;
  float j, jf;
  calculate_impulse(X, Y, contact, fadd(1.0f, e), normal, rA, rB, j);
  calculate_impulse(X, Y, contact, mu, tangent, rAf, rBf, jf);
  apply_impulses(X, Y, normal, j, rA, rB);
  apply_impulses(X, Y, tangent, jf, rAf, rBf);
}

__device__ __forceinline__ void load_poly(const Env& e, int b, float2 (&P)[6]) {
#pragma unroll
  for (int i = 0; i < 6; i++) P[i] = V2(e, b * 6 + i);
}
__device__ __forceinline__ void aabb6(const float2 (&P)[6], float2& mn, float2& mx) {
  mn.x = fminf(fminf(fminf(P[0].x, P[1].x), fminf(P[2].x, P[3].x)), fminf(P[4].x, P[5].x));
  mn.y = fminf(fminf(fminf(P[0].y, P[1].y), fminf(P[2].y, P[3].y)), fminf(P[4].y, P[5].y));
  mx.x = fmaxf(fmaxf(fmaxf(P[0].x, P[1].x), fmaxf(P[2].x, P[3].x)), fmaxf(P[4].x, P[5].x));
  mx.y = fmaxf(fmaxf(fmaxf(P[0].y, P[1].y), fmaxf(P[2].y, P[3].y)), fmaxf(P[4].y, P[5].y));
}
// BoundingBox.IsColliding, Skeleton.cs:133-140
__device__ __forceinline__ bool aabb_hit(float2 amin, float2 amax, float2 bmin, float2 bmax) {
  return amin.x < bmax.x && amax.x > bmin.x && amin.y < bmax.y && amax.y > bmin.y;
}

// ---------------------------------------------------------------- one candidate of RigidBody.ResolveCollisions (RigidBody.cs:66-96):
// a leg segment against the other segment of its own leg (both dynamic poles)
template <bool TRACE>
__device__ __forceinline__ void pole_pair(Env& e, int A, int B, wb_pair_trace* tr) {
  wb_pair_trace rec;
  if (TRACE) trace_init(rec, B);
  float2 PA[6], PB[6];
  load_poly(e, A, PA);
  load_poly(e, B, PB);
  float2 amin, amax, bmin, bmax;
  aabb6(PA, amin, amax);
  aabb6(PB, bmin, bmax);
  if (aabb_hit(amin, amax, bmin, bmax)) {
    if (TRACE) rec.aabb = 1;
    Sat s;
    s.depth = FLT_MAX;
    s.normal = mk2(0.0f, 0.0f);
    s.idx = -1;
    s.sep = false;
    // AxisChecks(A, B) then, only while nothing separated, AxisChecks(B, A): one loop over the 12 edges.  The polygons stay in
    // registers for the projections; only the two edge endpoints are fetched by (runtime) index.  A lane leaves the loop at its
    // first separating axis, exactly like the reference's early "return false".
    const int nA = nverts(A);
#pragma unroll 1
    for (int i = 0; i < 12 && !s.sep; i++) {
      const int own = i < 6 ? A : B;
      const int k = i < 6 ? i : i - 6;
      bool skip;
      const float2 axis = edge_axis(V2(e, own * 6 + k), V2(e, own * 6 + (k == 5 ? 0 : k + 1)), skip);
      float amn, amx, bmn, bmx;
      project6(PA, axis, amn, amx);
      project6(PB, axis, bmn, bmx);
      sat_accumulate(s, skip, axis, i < 6 ? i : nA + k, amn, amx, bmn, bmx);
    }
    if (!s.sep) {
      // orient: normal points from B towards A (SATCollision.cs:31-32, cached centroids)
      BodyDyn X = load_dyn(e, A);
      BodyDyn Y = load_dyn(e, B);
      float2 normal = s.normal;
      if (vdot(vsub(Y.c, X.c), normal) > 0.0f) normal = vmul(normal, -1.0f);
      const Face ref = significant_face_body(e, A, PA, normal);
      const Face inc = significant_face_body(e, B, PB, vneg(normal));
      float2 c0 = mk2(0.f, 0.f), c1 = mk2(0.f, 0.f);
      const int ncp = contact_points(ref, inc, normal, c0, c1);
      if (TRACE) trace_hit(rec, s, normal, ncp, c0, c1);
      // RigidBody.MoveObjects, RigidBody.cs:99-113 (both dynamic), applied even with 0 contact points
      const float2 dA = vhalf(vmul(normal, s.depth));
      const float2 dB = vhalf(vmul(vneg(normal), s.depth));
      move_body_regs(e, A, PA, dA);
      move_body_regs(e, B, PB, dB);
      if (ncp > 0) {  // impulses read the PRE-move velocities but the POST-move centroids
        X.c = vadd(X.c, dA);
        Y.c = vadd(Y.c, dB);
        resolve_impulses(X, Y, ncp, c0, c1, normal, e.e_ww, e.mu_ww);
        store_dyn(e, A, X);
        store_dyn(e, B, Y);
      }
    }
  }
  if (TRACE) {
    if (tr) *tr = rec;
  }
}

// a walker body against the static floor (scene constants in c_floor)
template <bool TRACE>
__device__ __forceinline__ void floor_pair(Env& e, int A, wb_pair_trace* tr) {
  wb_pair_trace rec;
  if (TRACE) trace_init(rec, FLOOR);
  float2 PA[6];
  load_poly(e, A, PA);
  float2 amin, amax;
  aabb6(PA, amin, amax);
  if (aabb_hit(amin, amax, c_floor.bb_min, c_floor.bb_max)) {
    e.flags |= (1 << A);  // if (body._isFloor) Collided = true  (before SAT: RigidBody.cs:75)
    if (TRACE) rec.aabb = 1;
    Sat s;
    s.depth = FLT_MAX;
    s.normal = mk2(0.0f, 0.0f);
    s.idx = -1;
    s.sep = false;
#pragma unroll 1
    for (int i = 0; i < 6 && !s.sep; i++) {  // AxisChecks(A, floor)
      bool skip;
      const float2 axis = edge_axis(V2(e, A * 6 + i), V2(e, A * 6 + (i == 5 ? 0 : i + 1)), skip);
      float omin, omax, tmin, tmax;
      project6(PA, axis, omin, omax);
      project_floor(axis, tmin, tmax);
      sat_accumulate(s, skip, axis, i, omin, omax, tmin, tmax);
    }
    const int nA = nverts(A);
#pragma unroll 1
    for (int i = 0; i < 4 && !s.sep; i++) {  // AxisChecks(floor, A): constant axes and constant own projection
      float tmin, tmax;
      project6(PA, c_floor.axis[i], tmin, tmax);
      sat_accumulate(s, c_floor.skip[i] != 0, c_floor.axis[i], nA + i, c_floor.pmin[i], c_floor.pmax[i], tmin, tmax);
    }
    if (!s.sep) {
      BodyDyn X = load_dyn(e, A);
      BodyDyn Y = floor_dyn();
      float2 normal = s.normal;
      if (vdot(vsub(Y.c, X.c), normal) > 0.0f) normal = vmul(normal, -1.0f);
      const Face ref = significant_face_body(e, A, PA, normal);
      const Face inc = significant_face_floor(vneg(normal));
      float2 c0 = mk2(0.f, 0.f), c1 = mk2(0.f, 0.f);
      const int ncp = contact_points(ref, inc, normal, c0, c1);
      if (TRACE) trace_hit(rec, s, normal, ncp, c0, c1);
      // MoveObjects with a static B: A.Move(normal * depth)
      const float2 dA = vmul(normal, s.depth);
      move_body_regs(e, A, PA, dA);
      if (ncp > 0) {
        X.c = vadd(X.c, dA);
        resolve_impulses(X, Y, ncp, c0, c1, normal, e.e_wf, e.mu_wf);
        store_dyn(e, A, X);  // the floor is never written (inverse mass/inertia 0)
      }
    }
  }
  if (TRACE) {
    if (tr) *tr = rec;
  }
}

// ---------------------------------------------------------------- RigidBody.Step, RigidBody.cs:54-61,116-140
template <bool TRACE>
__device__ __forceinline__ void body_step(Env& e, int b, float dt, wb_pair_trace* tr_base) {
  // StepLinearVelocity: v += a * dt (gravity (0, 980), Walker.cs:45); Skeleton.Move(v * dt)
  float2 v = V2(e, kV2Vel + b);
  v = vadd(v, vmul(mk2(0.0f, 980.0f), dt));
  const float2 d = vmul(v, dt);
  // StepAngularVelocity: angle = WrapAngle(angle + w * dt); Skeleton.Rotate(w * dt)
  const float w = F1(e, kFOmega + b);
  const float theta = fmul(w, dt);
  float ang = fadd(F1(e, kFAngle + b), theta);
  const float PI_F = 3.14159274f, TAU_F = 6.28318548f;
  if (ang > PI_F) ang = fsub(ang, TAU_F);
  else if (ang < -PI_F) ang = fadd(ang, TAU_F);
  float m11, m12;
  rotz(theta, m11, m12);
  const float m21 = -m12, m22 = m11;
  const float2 cen = vadd(V2(e, kV2Cen + b), d);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    // Move then Rotate (Vector2.Transform(p - centroid, R) + centroid, Skeleton.cs:93)
    float2 p = vadd(V2(e, b * 6 + i), d);
    p = vsub(p, cen);
    float2 t;
    t.x = fadd(fadd(fmul(p.x, m11), fmul(p.y, m21)), 0.0f);
    t.y = fadd(fadd(fmul(p.x, m12), fmul(p.y, m22)), 0.0f);
    V2(e, b * 6 + i) = vadd(t, cen);
  }
  V2(e, kV2Cen + b) = cen;
  V2(e, kV2Vel + b) = v;
  F1(e, kFAngle + b) = ang;
  // ResolveCollisions: candidates in Environment._rigidBodies order, skipping self and associated bodies (Walker.cs:204-208):
  // a leg segment meets the other segment of its own leg and the floor; the Body only the floor.  The floor comes first in the
  // list after the first Reset (Walker.cs:212-223).  Three uniform phases keep a warp whose environments disagree on the list
  // order from running the pole-pole pair twice: [floor if floor-first] [leg partner] [floor if floor-last].
  const int slot = (0x75420 >> (4 * b)) & 0xF;     // trace slot base {0,2,4,5,7}
  const int partner = (0x34F01 >> (4 * b)) & 0xF;  // {LLU, LLL, -, RLU, RLL}
  const bool floor_first = (e.flags & WB_FLAG_FLOOR_FIRST) != 0;
#pragma unroll 1
  for (int phase = 0; phase < 3; phase++) {
    if (phase == 1) {
      if (b != BODY) pole_pair<TRACE>(e, b, partner, (TRACE && tr_base) ? tr_base + slot + (floor_first ? 1 : 0) : nullptr);
    } else {
      const bool run = (b == BODY) ? (phase == 0) : ((phase == 0) == floor_first);
      if (run) floor_pair<TRACE>(e, b, (TRACE && tr_base) ? tr_base + slot + ((b == BODY || floor_first) ? 0 : 1) : nullptr);
    }
  }
}

// Walker.GetState, Walker.cs:132-152
__device__ __forceinline__ void observation(const Env& e, float* o) {
  const float2 j0 = V2(e, BODY * 6 + 1), j2 = V2(e, LLU * 6 + 2), j3 = V2(e, RLU * 6 + 2), bv = V2(e, kV2Vel + BODY);
  o[0] = fdiv(j0.x, 900.0f);
  o[1] = fdiv(j0.y, 500.0f);
  o[2] = fdiv(j2.x, 900.0f);
  o[3] = fdiv(j2.y, 500.0f);
  o[4] = fdiv(j3.x, 900.0f);
  o[5] = fdiv(j3.y, 500.0f);
  o[6] = fdiv(bv.x, 60.0f);
  o[7] = fdiv(bv.y, 60.0f);
  o[8] = F1(e, kFAngle + LLL);
  o[9] = F1(e, kFAngle + LLU);
  o[10] = F1(e, kFAngle + RLL);
  o[11] = F1(e, kFAngle + RLU);
}
__device__ __forceinline__ void store_observation(const Env& e, float* dst) {  // dst is 16-byte aligned (48 B per env)
  float o[12];
  observation(e, o);
  float4* d4 = reinterpret_cast<float4*>(dst);
  d4[0] = make_float4(o[0], o[1], o[2], o[3]);
  d4[1] = make_float4(o[4], o[5], o[6], o[7]);
  d4[2] = make_float4(o[8], o[9], o[10], o[11]);
}

// canonical record index (walker_b200.h) -> float index inside this env's columns; rows 88..91 (joint torques) stay in HBM
__device__ __forceinline__ float& record_ref(const Env& e, int f) {
  float* v2f = reinterpret_cast<float*>(e.v2);
  if (f < 58) {
    const int vtx = f >> 1, slot = vtx + (vtx >= 17 ? 1 : 0);
    return v2f[slot * kT * 2 + (f & 1)];
  }
  if (f < 78) {
    const int slot = kV2Cen + ((f - 58) >> 1);
    return v2f[slot * kT * 2 + (f & 1)];
  }
  return e.f[(f - 78) * kT];
}

// Walker.Reset + CreateCreature: fresh walker record (constants computed on the host with the reference's formulas)
__device__ __noinline__ void write_initial_record(const Env& e, float* torque_rows, int n_pad) {
#pragma unroll 1
  for (int f = 0; f < 88; f++) record_ref(e, f) = c_init_state[f];
  V2(e, BODY * 6 + 5) = V2(e, BODY * 6);
  for (int k = 0; k < 4; k++) torque_rows[(size_t)k * n_pad] = c_init_state[88 + k];
}

template <bool TRACE>
__global__ void __launch_bounds__(kT, 4) physics_scalar_kernel(const PhysicsParams p) {
  __shared__ float2 s_v2[kV2Count * kT];
  __shared__ float s_f[kFCount * kT];
  const int tid = threadIdx.x;
  const int env = blockIdx.x * kT + tid;
  if (env >= p.n) return;  // no block-level synchronisation anywhere: each thread owns its columns
  Env e;
  e.v2 = s_v2 + tid;
  e.f = s_f + tid;

  // ---- stage the record: SoA rows are contiguous over envs, so every row is one coalesced 128-byte line per warp
#pragma unroll
  for (int f = 0; f < 88; f++) record_ref(e, f) = p.state[(size_t)f * p.n_pad + env];
  V2(e, BODY * 6 + 5) = V2(e, BODY * 6);
  float* torque_rows = p.state + (size_t)88 * p.n_pad + env;

  e.flags = p.flags[env];
  int steps = p.steps[env];
  {
    const Material mw = c_materials[p.walker_mat[env]];
    const Material mf = c_materials[p.floor_mat[env]];
    e.im_w = mw.inverse_mass;
    e.ii_pole = fmul(0.001f, mw.inverse_mass);
    e.e_ww = net_max(mw.restitution, mw.restitution);
    e.mu_ww = net_min(mw.friction, mw.friction);
    e.e_wf = net_max(mw.restitution, mf.restitution);
    e.mu_wf = net_min(mw.friction, mf.friction);
  }
  float2 pos = mk2(p.pos[env], p.pos[p.n_pad + env]);

  if (p.phases & kPhaseResetMasked) {
    if (p.reset_mask == nullptr || p.reset_mask[env]) {
      write_initial_record(e, torque_rows, p.n_pad);
      e.flags = (p.phases & kPhaseFirstEpisode) ? 0 : WB_FLAG_FLOOR_FIRST;
      steps = 0;
      pos = V2(e, kV2Cen + BODY);  // InitialState -> Walker.Update
    }
  }
  if (p.phases & kPhaseIncSteps) steps++;

  if (p.phases & kPhaseTakeActions) {
    // Matrix.Clip (Matrix.cs:377-405), Walker.TakeActions (Walker.cs:66-75), Joint.SetTorque (Joint.cs:56-61)
    const float4 a4 = *reinterpret_cast<const float4*>(p.actions + (size_t)env * 4);
    const float act[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float a = act[k];
      if (a >= 1.0f) a = 1.0f;
      else if (a <= -1.0f) a = -1.0f;
      const float change = fsub(a, torque_rows[(size_t)k * p.n_pad]);
      torque_rows[(size_t)k * p.n_pad] = a;
      const int bodyB = (k == 0) ? LLU : (k == 1) ? RLU : (k == 2) ? LLL : RLL;  // Walker.cs:182-185
      F1(e, kFOmega + bodyB) = fadd(F1(e, kFOmega + bodyB), fmul(change, 5.0f));
    }
  }

  if (p.phases & kPhaseStepObjects) {
    const float dt = fdiv(p.dt, (float)p.iterations);  // deltaTime /= Hyperparameters.Iterations
#pragma unroll 1
    for (int it = 0; it < p.iterations; it++) {
      wb_joint_trace* jt = nullptr;
      wb_pair_trace* pt = nullptr;
      if (TRACE) {
        if (p.joint_trace) jt = p.joint_trace + ((size_t)env * p.iterations + it) * 4;
        if (p.pair_trace) pt = p.pair_trace + ((size_t)env * p.iterations + it) * WB_PAIR_SLOTS;  // all 9 slots are written every substep
      }
      // joints in creation order (Walker.cs:182-187): (Body v1, LLU v4) (Body v1, RLU v4) (LLU v2, LLL v3) (RLU v2, RLL v3)
#pragma unroll 1
      for (int k = 0; k < 4; k++) {
        const int A = (0x4122 >> (4 * k)) & 0xF, B = (0x3041 >> (4 * k)) & 0xF;
        joint_step<TRACE>(e, A, k < 2 ? 1 : 2, B, k < 2 ? 4 : 3, jt ? jt + k : nullptr);
      }
      // bodies in list order; the static floor's Update is a no-op (a = v = 0, returns before rotation/collisions)
#pragma unroll 1
      for (int b = 0; b < 5; b++) body_step<TRACE>(e, b, dt, pt);
    }
  }

  if (p.phases & kPhaseObserve) {
    // Walker.Update, Walker.cs:49-54
    const float2 prev = pos;
    pos = V2(e, kV2Cen + BODY);
    if (e.flags & ((1 << BODY) | (1 << LLU) | (1 << RLU))) e.flags |= WB_FLAG_TERMINAL;
    // CalculateReward, Environment.cs:148-154 (incl. the "-= -0.1f" sign quirk)
    const float dx = fsub(pos.x, prev.x);
    const float h = fdiv(V2(e, BODY * 6 + 1).y, 500.0f);
    float r = 0.0f;
    r = fadd(r, (dx > 0.0f && h < 1.6f) ? dx : 0.0f);
    r = fsub(r, (h > 1.65f) ? -0.1f : 0.0f);
    bool terminal = false;
    if ((e.flags & WB_FLAG_TERMINAL) || steps > p.max_timesteps) {  // Environment.cs:106-110
      if (e.flags & WB_FLAG_TERMINAL) r = fsub(r, 40.0f);
      terminal = true;
    }
    if (pos.x > 900.0f) {  // :113-117
      r = fadd(r, 80.0f);
      terminal = true;
    }
    if (terminal && (p.phases & kPhaseAutoReset)) {
      write_initial_record(e, torque_rows, p.n_pad);
      e.flags = WB_FLAG_FLOOR_FIRST;
      steps = 0;
      pos = V2(e, kV2Cen + BODY);
    }
    store_observation(e, p.obs + (size_t)env * WB_OBS);
    p.reward[env] = r;
    p.done[env] = terminal ? 1 : 0;
  } else if (p.phases & kPhaseObsOnly) {
    store_observation(e, p.obs + (size_t)env * WB_OBS);
  }

  p.flags[env] = e.flags;
  p.steps[env] = steps;
  p.pos[env] = pos.x;
  p.pos[p.n_pad + env] = pos.y;
  // ---- write the record back (same coalesced pattern)
#pragma unroll
  for (int f = 0; f < 88; f++) p.state[(size_t)f * p.n_pad + env] = record_ref(e, f);
}

// test hook: compares rcp_sqrt_rn(s) with __frcp_rn(__fsqrt_rn(s)) for every bit pattern in [first, first + count)
__global__ void rcp_sqrt_check_kernel(uint32_t first, uint64_t count, unsigned long long* mismatches, uint32_t* first_bad) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const uint32_t bits = first + (uint32_t)i;
    const float s = __uint_as_float(bits);
    const uint32_t a = __float_as_uint(rcp_sqrt_rn(s));
    const uint32_t b = __float_as_uint(__frcp_rn(__fsqrt_rn(s)));
    const bool both_nan = ((a & 0x7FFFFFFFu) > 0x7F800000u) && ((b & 0x7FFFFFFFu) > 0x7F800000u);
    if (a != b && !both_nan) {
      bad++;
      atomicMin(first_bad, bits);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace t1

// ---------------------------------------------------------------- host side
cudaError_t launch_rcp_sqrt_check(uint32_t first, uint64_t count, unsigned long long* mismatches_dev, uint32_t* first_bad_dev,
                                  cudaStream_t stream) {
  t1::rcp_sqrt_check_kernel<<<148 * 8, 256, 0, stream>>>(first, count, mismatches_dev, first_bad_dev);
  return cudaGetLastError();
}

cudaError_t upload_materials_scalar(const Material* table, int count) {
  return cudaMemcpyToSymbol(t1::c_materials, table, sizeof(Material) * count);
}

cudaError_t upload_scene_constants_scalar(const float* init_state92, const FloorConst* floor) {
  cudaError_t e = cudaMemcpyToSymbol(t1::c_init_state, init_state92, sizeof(float) * kStateFloats);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(t1::c_floor, floor, sizeof(FloorConst));
}

cudaError_t launch_physics_scalar(const PhysicsParams& p, bool trace, cudaStream_t stream) {
  const int grid = (p.n + t1::kT - 1) / t1::kT;
  if (trace)
    t1::physics_scalar_kernel<true><<<grid, t1::kT, 0, stream>>>(p);
  else
    t1::physics_scalar_kernel<false><<<grid, t1::kT, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace wb
