// physics.cu -- lockstep rigid-body step for N independent BipedalWalker environments (sm_100a).
//
// Replaces, for a whole batch per launch, the reference's Environment.StepObjects and everything under it
// (Environment.cs:126-143; Joint.cs:31-61; RigidBody.cs:54-140; Skeleton.cs:76-176; SATCollision.cs:15-104;
// ContactPoints.cs:13-134; Impulses.cs:12-115) plus the walker glue around it (Walker.cs:49-75,132-152,
// Environment.cs:96-122,148-154,167-180).
//
// Mapping: one GROUP of 16 lanes owns one environment; 8 environments per CTA.  The per-env state record
// (92 floats) is read once from the SoA arrays in HBM with coalesced float4 loads (one 32-byte sector per
// SoA row per CTA), transposed into shared memory, advanced by all `iterations` substeps on chip and
// written back once.  Inside an environment the joint/body order is the reference's sequential
// Gauss-Seidel order; lanes parallelise over vertices (Move/Rotate), over SAT axes (one axis per lane,
// REDUX min over the group picks the first minimal-depth axis) and over the two bodies of a pair.
//
// Arithmetic contract: IEEE binary32, every multiply/add individually rounded (intrinsics below never
// contract to FMA; the TU is additionally compiled with -fmad=false), correctly rounded 1/x, sqrt and
// division, and double-precision sin/cos rounded to float for Skeleton.Rotate -- the same operation order
// as the reference's C# (SURVEY.md Appendix A/C).  There is no CPU path.
#include "physics.cuh"
#include "physics_math.cuh"

#include <cfloat>

namespace wb {

__constant__ Material c_materials[WB_MAX_MATERIALS];
__constant__ float c_init_state[kStateFloats];  // state record of a freshly created walker
__constant__ float c_floor[10];                 // 4 floor vertices (x,y) + cached centroid

// ---------------------------------------------------------------- shared-memory record of one env
// vertices of 6 bodies (floor = body 5) padded to 6 each, then centroids, velocities, omega, angle, torque
constexpr int kSVert = 0;     // + body*12 + i*2
constexpr int kSCen = 72;     // + body*2  (6 bodies)
constexpr int kSVel = 84;     // + body*2  (5 bodies)
constexpr int kSOmega = 94;   // + body
constexpr int kSAngle = 99;   // + body
constexpr int kSTorque = 104; // + joint
constexpr int kSStride = 108; // floats per env; 108 % 32 = 12 keeps the two envs of a warp on disjoint banks

enum { LLL = 0, LLU = 1, BODY = 2, RLL = 3, RLU = 4, FLOOR = 5 };

__device__ __forceinline__ int svert(int b, int i) { return kSVert + b * 12 + i * 2; }
__device__ __forceinline__ int nverts(int b) { return b == BODY ? 5 : (b == FLOOR ? 4 : 6); }

// canonical record index (walker_b200.h) -> shared-memory slot
__device__ __forceinline__ int record_to_smem(int f) {
  return f + (f >= 34 ? 2 : 0) + (f >= 58 ? 12 : 0) + (f >= 68 ? 2 : 0);
}


// ---------------------------------------------------------------- per-group context
struct Ctx {
  float* s;       // this env's shared-memory record
  unsigned mask;  // the 16 lanes of the group
  int gshift;     // 0 or 16: first lane of the group inside the warp
  int gl;         // lane inside the group
  int flags;      // Collided bits, Terminal, floor-first (group-uniform)
  // material-derived constants (RigidBody ctor, RigidBody.cs:36-50; Impulses.cs:16-17)
  float im_w;     // walker inverse mass
  float ii_pole;  // 0.001f * inverse mass
  float ii_body;  // Walker.cs:168
  float e_ww, mu_ww, e_wf, mu_wf;
};

__device__ __forceinline__ float inv_mass(const Ctx& c, int b) { return b == FLOOR ? 0.0f : c.im_w; }
__device__ __forceinline__ float inv_inertia(const Ctx& c, int b) {
  return b == FLOOR ? 0.0f : (b == BODY ? c.ii_body : c.ii_pole);
}

struct BodyDyn {
  float2 c, v;
  float w, im, ii;
};

__device__ __forceinline__ BodyDyn load_dyn(const Ctx& c, int b) {
  BodyDyn d;
  d.c = lds2(c.s, kSCen + b * 2);
  if (b == FLOOR) {
    d.v = mk2(0.0f, 0.0f);
    d.w = 0.0f;
  } else {
    d.v = lds2(c.s, kSVel + b * 2);
    d.w = c.s[kSOmega + b];
  }
  d.im = inv_mass(c, b);
  d.ii = inv_inertia(c, b);
  return d;
}

// Impulses.CalculateImpulse, Impulses.cs:86-115
__device__ __forceinline__ void calculate_impulse(const BodyDyn& A, const BodyDyn& B, float2 contact, float force, float2 n,
                                                  float2& rA, float2& rB, float& impulse) {
  rA = vsub(contact, A.c);
  float2 perpA = mk2(-rA.y, rA.x);
  float kA = vdot(n, perpA);
  rB = vsub(contact, B.c);
  float2 perpB = mk2(-rB.y, rB.x);
  float kB = vdot(n, perpB);
  float2 va = vadd(A.v, vmul(perpA, A.w));
  float2 vb = vadd(B.v, vmul(perpB, B.w));
  float2 vrel = vsub(vb, va);
  float vn = vdot(vrel, n);
  float j = fmul(-force, vn);
  float denom = fadd(fadd(fadd(A.im, B.im), fmul(fmul(kA, kA), A.ii)), fmul(fmul(kB, kB), B.ii));
  impulse = fdiv(j, denom);
}

// Impulses.ApplyImpulses, Impulses.cs:57-82
__device__ __forceinline__ void apply_impulses(BodyDyn& A, BodyDyn& B, float2 n, float impulse, float2 rA, float2 rB) {
  float2 J = vmul(n, impulse);
  float2 velA = vsub(A.v, vmul(J, A.im));
  float2 velB = vadd(B.v, vmul(J, B.im));
  float2 perpA = mk2(-rA.y, rA.x);
  float wA = fsub(A.w, fmul(vdot(perpA, J), A.ii));
  float2 perpB = mk2(-rB.y, rB.x);
  float wB = fadd(B.w, fmul(vdot(perpB, J), B.ii));
  A.v = velA;
  B.v = velB;
  A.w = wA;
  B.w = wB;
}

// min over the 16 lanes of the group / over the 8 lanes of my half (butterflies stay inside the group)
__device__ __forceinline__ unsigned group_umin16(const Ctx& c, unsigned v) {
  v = min(v, __shfl_xor_sync(c.mask, v, 1));
  v = min(v, __shfl_xor_sync(c.mask, v, 2));
  v = min(v, __shfl_xor_sync(c.mask, v, 4));
  v = min(v, __shfl_xor_sync(c.mask, v, 8));
  return v;
}
__device__ __forceinline__ unsigned half_umin8(const Ctx& c, unsigned v) {
  v = min(v, __shfl_xor_sync(c.mask, v, 1));
  v = min(v, __shfl_xor_sync(c.mask, v, 2));
  v = min(v, __shfl_xor_sync(c.mask, v, 4));
  return v;
}

// lane 0 stores body X's (v, w), lane 8 body Y's; the floor is never written (inverse mass/inertia 0)
__device__ __forceinline__ void store_dyn_pair(const Ctx& c, int X, const BodyDyn& dx, int Y, const BodyDyn& dy) {
  __syncwarp(c.mask);  // every lane has read the pre-impulse values
  if (c.gl == 0 && X != FLOOR) {
    sts2(c.s, kSVel + X * 2, dx.v);
    c.s[kSOmega + X] = dx.w;
  }
  if (c.gl == 8 && Y != FLOOR) {
    sts2(c.s, kSVel + Y * 2, dy.v);
    c.s[kSOmega + Y] = dy.w;
  }
  __syncwarp(c.mask);
}

// Skeleton.Move on up to two bodies at once: lanes 0..5 body A's vertices, lane 6 its centroid; lanes 8..13 / 14 body B
__device__ __forceinline__ void move_pair(const Ctx& c, bool moveA, int A, float2 dA, bool moveB, int B, float2 dB) {
  const bool hi = c.gl >= 8;
  const int i = c.gl & 7;
  const int b = hi ? B : A;
  const float2 d = hi ? dB : dA;
  const bool doit = hi ? moveB : moveA;
  __syncwarp(c.mask);  // every lane has read the pre-move vertices/centroids it needs
  if (doit) {
    if (i < nverts(b)) {
      sts2(c.s, svert(b, i), vadd(lds2(c.s, svert(b, i)), d));
    } else if (i == 6) {
      sts2(c.s, kSCen + b * 2, vadd(lds2(c.s, kSCen + b * 2), d));
    }
  }
  __syncwarp(c.mask);
}

// ---------------------------------------------------------------- Joint.Step, Joint.cs:31-41
template <bool TRACE>
__device__ __forceinline__ void joint_step(const Ctx& c, int A, int ia, int B, int ib, wb_joint_trace* tr) {
  const float2 pA = lds2(c.s, svert(A, ia));
  const float2 pB = lds2(c.s, svert(B, ib));
  float2 ab = vsub(pB, pA);
  const float depth = fsqrt(fadd(fmul(ab.x, ab.x), fmul(ab.y, ab.y)));  // Vector2.Length
  if (TRACE) {
    if (tr && c.gl == 0) {
      tr->active = !(depth < 0.1f);
      tr->depth = depth;
    }
  }
  if (depth < 0.1f) return;
  ab = vnormalize(ab);
  const float2 dA = vhalf(vmul(ab, depth));
  const float2 dB = vhalf(vmul(vneg(ab), depth));
  BodyDyn X = load_dyn(c, B);  // Manifold(bodyA := joint._bodyB, bodyB := joint._bodyA), Joint.cs:40
  BodyDyn Y = load_dyn(c, A);
  move_pair(c, true, A, dA, true, B, dB);
  // every lane redoes the two moved joint points and centroids in registers (same ops as the stores above)
  const float2 pA2 = vadd(pA, dA);
  const float2 pB2 = vadd(pB, dB);
  Y.c = vadd(Y.c, dA);
  X.c = vadd(X.c, dB);
  const float2 contact = vhalf(vadd(pA2, pB2));  // Vector2.Divide(p0 + p1, 2), Impulses.cs:35
  float2 rX, rY;
  float j;
  calculate_impulse(X, Y, contact, fadd(1.0f, 1.0f), ab, rX, rY, j);
  apply_impulses(X, Y, ab, j, rX, rY);
  store_dyn_pair(c, B, X, A, Y);
}

// ---------------------------------------------------------------- broadphase, Skeleton.cs:133-176
// lanes 0..3 hold A's {minX, maxX, minY, maxY}, lanes 4..7 B's; each lane folds one coordinate over the vertices
__device__ __forceinline__ bool aabb_overlap(const Ctx& c, int A, int B) {
  const int q = c.gl & 3;
  const int b = (c.gl & 4) ? B : A;
  const int comp = q >> 1;      // 0: x, 1: y
  const bool want_max = q & 1;
  const int n = nverts(b);
  float acc = want_max ? -FLT_MAX : FLT_MAX;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < n) {
      float t = c.s[svert(b, j) + comp];
      acc = want_max ? fmaxf(acc, t) : fminf(acc, t);
    }
  }
  // A.min < B.max (lanes 0,2 vs 5,7)  and  A.max > B.min (lanes 1,3 vs 4,6)
  const float other = __shfl_sync(c.mask, acc, c.gshift + ((c.gl & 3) ^ 1) + 4);
  bool ok = true;
  if (c.gl < 4) ok = want_max ? (acc > other) : (acc < other);
  return __ballot_sync(c.mask, !ok) == 0;  // bits outside the group are not in mask
}

// ---------------------------------------------------------------- SAT, SATCollision.cs:15-104
// lanes 0..5: axes from A's edges, lanes 8..13: axes from B's edges (the reference's evaluation order).
__device__ __forceinline__ bool sat(const Ctx& c, int A, int B, float2& normal, float& depth, int& axis_idx) {
  const int nA = nverts(A), nB = nverts(B);
  const bool hi = c.gl >= 8;
  const int i = c.gl & 7;
  const int own = hi ? B : A;
  const int n = hi ? nB : nA;
  const bool act = i < n;
  const int i0 = act ? i : 0;
  const int i1 = (i0 + 1 == n) ? 0 : i0 + 1;
  const float2 p0 = lds2(c.s, svert(own, i0));
  const float2 p1 = lds2(c.s, svert(own, i1));
  const float2 edge = vsub(p1, p0);
  float2 axis = mk2(-edge.y, edge.x);
  const bool skip = (axis.x == 0.0f) && (axis.y == 0.0f);
  axis = vnormalize(axis);
  float minA = FLT_MAX, maxA = -FLT_MAX, minB = FLT_MAX, maxB = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < nA) {
      float t = vdot(axis, lds2(c.s, svert(A, j)));
      minA = fminf(minA, t);
      maxA = fmaxf(maxA, t);
    }
  }
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < nB) {
      float t = vdot(axis, lds2(c.s, svert(B, j)));
      minB = fminf(minB, t);
      maxB = fmaxf(maxB, t);
    }
  }
  // Projection.IsOverlapping: both the depth and the test are symmetric in the two projections
  const float temp = fminf(fsub(maxB, minA), fsub(maxA, minB));
  const bool overlapping = (minA < maxB) && (minB < maxA);
  const bool elig = act && !skip;
  const unsigned separating = __ballot_sync(c.mask, elig && !overlapping);
  normal = mk2(0.0f, 0.0f);
  depth = FLT_MAX;
  axis_idx = -1;
  if (separating) return false;
  // every eligible depth is > 0 here, so the uint order of the bit patterns is the float order;
  // "tempDepth >= depth -> continue" keeps the FIRST minimal axis: lowest lane among the minima.
  const unsigned key = elig ? __float_as_uint(temp) : 0xFFFFFFFFu;
  const unsigned best = group_umin16(c, key);
  const unsigned winners = (__ballot_sync(c.mask, elig && key == best) >> c.gshift) & 0xFFFFu;
  if (winners) {
    const int w = __ffs(winners) - 1;
    normal.x = __shfl_sync(c.mask, axis.x, c.gshift + w);
    normal.y = __shfl_sync(c.mask, axis.y, c.gshift + w);
    depth = __uint_as_float(best);
    axis_idx = (w < 8) ? w : nA + (w - 8);
  }
  // orient: normal points from B towards A (SATCollision.cs:31-32, cached centroids)
  const float2 dir = vsub(lds2(c.s, kSCen + B * 2), lds2(c.s, kSCen + A * 2));
  if (vdot(dir, normal) > 0.0f) normal = vmul(normal, -1.0f);
  return true;
}

// ---------------------------------------------------------------- contact points, ContactPoints.cs:13-134
struct Face {
  float2 a, b, max;
};

// GetSignificantVertex + GetSignificantFace, ContactPoints.cs:79-113, for BOTH polygons at once:
// lanes 0..7 work on polygon A with `normal` (reference face), lanes 8..15 on polygon B with -normal (incident face).
// Lane i projects vertex i; the first strictly-smallest projection wins (min over the half on an order-preserving key,
// lowest lane among equals), exactly as the sequential "if (!(projection < minimumDistance)) continue" scan.
__device__ __forceinline__ void significant_faces(const Ctx& c, int A, int B, float2 normal, Face& fa, Face& fb) {
  const bool hi = c.gl >= 8;
  const int i = c.gl & 7;
  const int P = hi ? B : A;
  const int n = nverts(P);
  const float2 nrm = hi ? vneg(normal) : normal;
  unsigned key = 0xFFFFFFFFu;
  if (i < n) {
    const float pr = vdot(lds2(c.s, svert(P, i)), nrm);
    if (pr < FLT_MAX) {  // NaN and values >= float.MaxValue never replace the initial minimum
      const unsigned b = __float_as_uint(fadd(pr, 0.0f));  // -0 -> +0: equal values get equal keys
      key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
  }
  const unsigned best = half_umin8(c, key);
  const unsigned winners = (__ballot_sync(c.mask, key == best) >> (c.gshift + (hi ? 8 : 0))) & 0xFFu;
  const int k = (best == 0xFFFFFFFFu) ? -1 : (__ffs(winners) - 1);
  const float2 sv = (k >= 0) ? lds2(c.s, svert(P, k)) : mk2(0.0f, 0.0f);
  int ka = k + 1;
  if (ka >= n) ka -= n;
  int kb = k - 1;
  if (kb < 0) kb += n;
  const float2 pa = lds2(c.s, svert(P, ka));
  const float2 pb = lds2(c.s, svert(P, kb));
  const float2 after = vnormalize(vsub(sv, pa));
  const float2 before = vnormalize(vsub(sv, pb));
  const bool use_before = vdot(nrm, before) >= vdot(nrm, after);
  const float2 a = use_before ? sv : pa;
  const float2 b = use_before ? pb : sv;
  const int la = c.gshift, lb = c.gshift + 8;
  fa.a = mk2(__shfl_sync(c.mask, a.x, la), __shfl_sync(c.mask, a.y, la));
  fa.b = mk2(__shfl_sync(c.mask, b.x, la), __shfl_sync(c.mask, b.y, la));
  fa.max = mk2(__shfl_sync(c.mask, sv.x, la), __shfl_sync(c.mask, sv.y, la));
  fb.a = mk2(__shfl_sync(c.mask, a.x, lb), __shfl_sync(c.mask, a.y, lb));
  fb.b = mk2(__shfl_sync(c.mask, b.x, lb), __shfl_sync(c.mask, b.y, lb));
  fb.max = mk2(__shfl_sync(c.mask, sv.x, lb), __shfl_sync(c.mask, sv.y, lb));
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
    float2 e = vsub(b, a);
    const float location = fdiv(da, fsub(da, db));
    e = vmul(e, location);
    e = vadd(e, a);
    if (cnt == 0) o0 = e; else if (cnt == 1) o1 = e;
    cnt++;
  }
  return cnt;
}

__device__ __forceinline__ bool veq(float2 a, float2 b) { return a.x == b.x && a.y == b.y; }

// GetContactPoints, ContactPoints.cs:13-53
__device__ __forceinline__ int contact_points(const Ctx& c, int A, int B, float2 normal, float2& c0, float2& c1) {
  Face ref, inc;
  significant_faces(c, A, B, normal, ref, inc);
  float2 rf = vsub(ref.b, ref.a);
  const float2 ifv = vsub(inc.b, inc.a);
  if (fabsf(vdot(rf, normal)) > fabsf(vdot(ifv, normal))) {
    Face t = ref;
    ref = inc;
    inc = t;
    rf = vsub(ref.b, ref.a);
  }
  rf = vnormalize(rf);
  float offset = vdot(rf, ref.a);
  float2 p0 = mk2(0.f, 0.f), p1 = mk2(0.f, 0.f);
  int cnt = clip_vectors(inc.a, inc.b, rf, offset, p0, p1);
  if (cnt < 2) return 0;
  offset = vdot(rf, ref.b);
  float2 q0 = mk2(0.f, 0.f), q1 = mk2(0.f, 0.f);
  cnt = clip_vectors(p0, p1, vneg(rf), -offset, q0, q1);
  if (cnt < 2) return 0;
  cnt = 2;  // a third point (never read by the reference's First()/Last() logic below unless it is Last())
  // NOTE: ClipVectors can return 3 points only when da >= 0, db >= 0 and da*db < 0 at once, which is impossible.
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

// ---------------------------------------------------------------- one candidate pair of RigidBody.ResolveCollisions, RigidBody.cs:66-96
template <bool TRACE>
__device__ __forceinline__ void resolve_pair(Ctx& c, int A, int B, wb_pair_trace* tr) {
  wb_pair_trace rec;
  if (TRACE) {
    rec.other = B;
    rec.aabb = 0;
    rec.sat = 0;
    rec.axis = -1;
    rec.nx = rec.ny = rec.depth = 0.0f;
    rec.ncontacts = 0;
    rec.c0x = rec.c0y = rec.c1x = rec.c1y = 0.0f;
  }
  const bool hit = aabb_overlap(c, A, B);
  if (hit) {
    if (B == FLOOR) c.flags |= (1 << A);  // if (body._isFloor) Collided = true  (before SAT: RigidBody.cs:75)
    float2 normal;
    float depth;
    int axis;
    const bool colliding = sat(c, A, B, normal, depth, axis);
    if (TRACE) rec.aabb = 1;
    if (colliding) {
      float2 c0 = mk2(0.f, 0.f), c1 = mk2(0.f, 0.f);
      const int ncp = contact_points(c, A, B, normal, c0, c1);
      if (TRACE) {
        rec.sat = 1;
        rec.axis = axis;
        rec.nx = normal.x;
        rec.ny = normal.y;
        rec.depth = depth;
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
      // impulses read the PRE-move velocities but the POST-move centroids (MoveObjects precedes ResolveCollisions)
      BodyDyn X = load_dyn(c, A);
      BodyDyn Y = load_dyn(c, B);
      // RigidBody.MoveObjects, RigidBody.cs:99-113 (A is never static: static bodies return before ResolveCollisions)
      float2 dA, dB;
      bool moveB;
      if (B == FLOOR) {
        dA = vmul(normal, depth);
        dB = mk2(0.f, 0.f);
        moveB = false;
      } else {
        dA = vhalf(vmul(normal, depth));
        dB = vhalf(vmul(vneg(normal), depth));
        moveB = true;
      }
      move_pair(c, true, A, dA, moveB, B, dB);
      if (ncp > 0) {  // Impulses.ResolveCollisions, Impulses.cs:12-28
        X.c = vadd(X.c, dA);
        if (moveB) Y.c = vadd(Y.c, dB);
        const float e = (B == FLOOR) ? c.e_wf : c.e_ww;
        const float mu = (B == FLOOR) ? c.mu_wf : c.mu_ww;
        const float2 contact = (ncp == 2) ? vhalf(vadd(c0, c1)) : c0;
        // both impulses come from the PRE-impulse velocities (Impulses.cs:23-24): lanes 0..7 evaluate the normal one,
        // lanes 8..15 the tangential one, then both are applied in the reference's order (:26-27)
        const bool hi = c.gl >= 8;
        const float2 tangent = mk2(-normal.y, normal.x);
        float2 rA, rB;
        float jmine;
        calculate_impulse(X, Y, contact, hi ? mu : fadd(1.0f, e), hi ? tangent : normal, rA, rB, jmine);
        const float j = __shfl_sync(c.mask, jmine, c.gshift);
        const float jf = __shfl_sync(c.mask, jmine, c.gshift + 8);
        apply_impulses(X, Y, normal, j, rA, rB);
        apply_impulses(X, Y, tangent, jf, rA, rB);
        store_dyn_pair(c, A, X, B, Y);
      }
    }
  }
  if (TRACE) {
    if (tr && c.gl == 0) *tr = rec;
  }
}

// ---------------------------------------------------------------- RigidBody.Step, RigidBody.cs:54-61,116-140
template <bool TRACE>
__device__ __forceinline__ void body_step(Ctx& c, int b, float dt, wb_pair_trace* tr_base) {
  const int n = nverts(b);
  // StepLinearVelocity: v += a * dt (gravity (0, 980), Walker.cs:45); Skeleton.Move(v * dt)
  float2 v = lds2(c.s, kSVel + b * 2);
  v = vadd(v, vmul(mk2(0.0f, 980.0f), dt));
  const float2 d = vmul(v, dt);
  // StepAngularVelocity: angle = WrapAngle(angle + w * dt); Skeleton.Rotate(w * dt)
  const float w = c.s[kSOmega + b];
  const float theta = fmul(w, dt);
  float ang = fadd(c.s[kSAngle + b], theta);
  const float PI_F = 3.14159274f, TAU_F = 6.28318548f;
  if (ang > PI_F) ang = fsub(ang, TAU_F);
  else if (ang < -PI_F) ang = fadd(ang, TAU_F);
  float m11, m12;
  rotz(theta, m11, m12);
  const float m21 = -m12, m22 = m11;
  const float2 cen = vadd(lds2(c.s, kSCen + b * 2), d);
  __syncwarp(c.mask);  // all reads of the old centroid/velocity/angle are done
  if (c.gl < n) {
    // Move then Rotate of vertex gl (Vector2.Transform(p - centroid, R) + centroid, Skeleton.cs:93)
    float2 p = vadd(lds2(c.s, svert(b, c.gl)), d);
    p = vsub(p, cen);
    float2 t;
    t.x = fadd(fadd(fmul(p.x, m11), fmul(p.y, m21)), 0.0f);
    t.y = fadd(fadd(fmul(p.x, m12), fmul(p.y, m22)), 0.0f);
    sts2(c.s, svert(b, c.gl), vadd(t, cen));
  } else if (c.gl == 8) {
    sts2(c.s, kSCen + b * 2, cen);
    sts2(c.s, kSVel + b * 2, v);
    c.s[kSAngle + b] = ang;
  }
  __syncwarp(c.mask);
  // ResolveCollisions: candidates in Environment._rigidBodies order, skipping self and associated bodies (Walker.cs:204-208):
  // a leg segment meets the other segment of its own leg and the floor; the Body only the floor.
  const int slot = (0x75420 >> (4 * b)) & 0xF;             // trace slot base {0,2,4,5,7}
  const int partner = (0x34F01 >> (4 * b)) & 0xF;          // {LLU, LLL, -, RLU, RLL}
  const int ncand = (b == BODY) ? 1 : 2;
  const bool floor_first = (c.flags & WB_FLAG_FLOOR_FIRST) != 0;
#pragma unroll 1
  for (int k = 0; k < ncand; k++) {
    const int other = (b == BODY || (k == 0) == floor_first) ? FLOOR : partner;
    wb_pair_trace* tr = (TRACE && tr_base) ? tr_base + slot + k : nullptr;
    resolve_pair<TRACE>(c, b, other, tr);
  }
}

// Walker.GetState, Walker.cs:132-152: lane k produces observation k
__device__ __forceinline__ float observation(const Ctx& c, int k) {
  switch (k) {
    case 0: return fdiv(c.s[svert(BODY, 1)], 900.0f);
    case 1: return fdiv(c.s[svert(BODY, 1) + 1], 500.0f);
    case 2: return fdiv(c.s[svert(LLU, 2)], 900.0f);
    case 3: return fdiv(c.s[svert(LLU, 2) + 1], 500.0f);
    case 4: return fdiv(c.s[svert(RLU, 2)], 900.0f);
    case 5: return fdiv(c.s[svert(RLU, 2) + 1], 500.0f);
    case 6: return fdiv(c.s[kSVel + BODY * 2], 60.0f);
    case 7: return fdiv(c.s[kSVel + BODY * 2 + 1], 60.0f);
    case 8: return c.s[kSAngle + LLL];
    case 9: return c.s[kSAngle + LLU];
    case 10: return c.s[kSAngle + RLL];
    default: return c.s[kSAngle + RLU];
  }
}

// Walker.Reset + CreateCreature: fresh walker record (constants computed on the host with the reference's formulas)
__device__ __forceinline__ void write_initial_record(const Ctx& c) {
  for (int f = c.gl; f < kStateFloats; f += 16) c.s[record_to_smem(f)] = c_init_state[f];
  __syncwarp(c.mask);
}

// LANES = 16: two environments share a warp (fewest warp-instructions per env; best when the GPU is full).
// LANES = 32: one environment per warp, lanes 0..15 work: the group mask is the compile-time constant 0xFFFF, so every
//             warp-level primitive is a plain WARPSYNC/SHFL/VOTE (a per-group mask costs MATCH+REDUX+VOTE+branch each),
//             there is no cross-environment divergence, and twice as many warps hide latency when N is small.
template <int LANES, bool TRACE>
__global__ void __launch_bounds__(kEnvsPerCta * LANES, LANES == 32 ? 4 : 1) physics_step_kernel(const PhysicsParams p) {
  __shared__ __align__(16) float smem[kEnvsPerCta * kSStride];
  const int tid = threadIdx.x;
  const int env0 = blockIdx.x * kEnvsPerCta;

  // ---- stage the 8 records: SoA rows are contiguous over envs, one float4 = 4 envs of one field
  for (int idx = tid; idx < kStateFloats * 2; idx += blockDim.x) {
    const int f = idx >> 1, h = idx & 1;
    const float4 v = *reinterpret_cast<const float4*>(p.state + (size_t)f * p.n_pad + env0 + h * 4);
    const int so = record_to_smem(f);
    smem[(h * 4 + 0) * kSStride + so] = v.x;
    smem[(h * 4 + 1) * kSStride + so] = v.y;
    smem[(h * 4 + 2) * kSStride + so] = v.z;
    smem[(h * 4 + 3) * kSStride + so] = v.w;
  }
  const int g = tid / LANES;  // group = env inside the CTA
  const int env = env0 + g;
  Ctx c;
  c.s = smem + g * kSStride;
  c.gl = tid % LANES;
  if (LANES == 16) {
    c.gshift = (tid & 16);
    c.mask = 0xFFFFu << c.gshift;
  } else {
    c.gshift = 0;
    c.mask = 0x0000FFFFu;
  }
  if (c.gl < 10) c.s[(c.gl < 8 ? svert(FLOOR, 0) : kSCen + FLOOR * 2 - 8) + c.gl] = c_floor[c.gl];
  __syncthreads();

  const bool live = env < p.n && c.gl < 16;
  if (live) {
    c.flags = p.flags[env];
    int steps = p.steps[env];
    {
      const Material mw = c_materials[p.walker_mat[env]];
      const Material mf = c_materials[p.floor_mat[env]];
      c.im_w = mw.inverse_mass;
      c.ii_pole = fmul(0.001f, mw.inverse_mass);
      c.ii_body = 0.0003f;
      c.e_ww = net_max(mw.restitution, mw.restitution);
      c.mu_ww = net_min(mw.friction, mw.friction);
      c.e_wf = net_max(mw.restitution, mf.restitution);
      c.mu_wf = net_min(mw.friction, mf.friction);
    }
    float2 pos = mk2(p.pos[env], p.pos[p.n_pad + env]);

    if (p.phases & kPhaseResetMasked) {
      if (p.reset_mask == nullptr || p.reset_mask[env]) {
        write_initial_record(c);
        c.flags = (p.phases & kPhaseFirstEpisode) ? 0 : WB_FLAG_FLOOR_FIRST;
        steps = 0;
        pos = lds2(c.s, kSCen + BODY * 2);  // InitialState -> Walker.Update
      }
    }
    if (p.phases & kPhaseIncSteps) steps++;

    if (p.phases & kPhaseTakeActions) {
      // Matrix.Clip (Matrix.cs:377-405), Walker.TakeActions (Walker.cs:66-75), Joint.SetTorque (Joint.cs:56-61)
      if (c.gl < 4) {
        float a = p.actions[(size_t)env * 4 + c.gl];
        if (a >= 1.0f) a = 1.0f;
        else if (a <= -1.0f) a = -1.0f;
        const float change = fsub(a, c.s[kSTorque + c.gl]);
        c.s[kSTorque + c.gl] = a;
        const int bodyB = (c.gl == 0) ? LLU : (c.gl == 1) ? RLU : (c.gl == 2) ? LLL : RLL;  // Walker.cs:182-185
        c.s[kSOmega + bodyB] = fadd(c.s[kSOmega + bodyB], fmul(change, 5.0f));
      }
      __syncwarp(c.mask);
    }

    if (p.phases & kPhaseStepObjects) {
      const float dt = fdiv(p.dt, (float)p.iterations);  // deltaTime /= Hyperparameters.Iterations
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
          joint_step<TRACE>(c, A, k < 2 ? 1 : 2, B, k < 2 ? 4 : 3, jt ? jt + k : nullptr);
        }
        // bodies in list order; the static floor's Update is a no-op (a = v = 0, returns before rotation/collisions)
#pragma unroll 1
        for (int b = 0; b < 5; b++) body_step<TRACE>(c, b, dt, pt);
      }
    }

    if (p.phases & kPhaseObserve) {
      // Walker.Update, Walker.cs:49-54
      const float2 prev = pos;
      pos = lds2(c.s, kSCen + BODY * 2);
      if (c.flags & ((1 << BODY) | (1 << LLU) | (1 << RLU))) c.flags |= WB_FLAG_TERMINAL;
      // CalculateReward, Environment.cs:148-154 (incl. the "-= -0.1f" sign quirk)
      const float dx = fsub(pos.x, prev.x);
      const float h = fdiv(c.s[svert(BODY, 1) + 1], 500.0f);
      float r = 0.0f;
      r = fadd(r, (dx > 0.0f && h < 1.6f) ? dx : 0.0f);
      r = fsub(r, (h > 1.65f) ? -0.1f : 0.0f);
      bool terminal = false;
      if ((c.flags & WB_FLAG_TERMINAL) || steps > p.max_timesteps) {  // Environment.cs:106-110
        if (c.flags & WB_FLAG_TERMINAL) r = fsub(r, 40.0f);
        terminal = true;
      }
      if (pos.x > 900.0f) {  // :113-117
        r = fadd(r, 80.0f);
        terminal = true;
      }
      if (terminal && (p.phases & kPhaseAutoReset)) {
        __syncwarp(c.mask);
        write_initial_record(c);
        c.flags = WB_FLAG_FLOOR_FIRST;
        steps = 0;
        pos = lds2(c.s, kSCen + BODY * 2);
      }
      if (c.gl < WB_OBS) p.obs[(size_t)env * WB_OBS + c.gl] = observation(c, c.gl);
      if (c.gl == 12) p.reward[env] = r;
      if (c.gl == 13) p.done[env] = terminal ? 1 : 0;
    } else if (p.phases & kPhaseObsOnly) {
      if (c.gl < WB_OBS) p.obs[(size_t)env * WB_OBS + c.gl] = observation(c, c.gl);
    }

    if (c.gl == 0) {
      p.flags[env] = c.flags;
      p.steps[env] = steps;
      p.pos[env] = pos.x;
      p.pos[p.n_pad + env] = pos.y;
    }
  }
  __syncthreads();
  // ---- write the records back (same coalesced pattern)
  for (int idx = tid; idx < kStateFloats * 2; idx += blockDim.x) {
    const int f = idx >> 1, h = idx & 1;
    const int so = record_to_smem(f);
    float4 v;
    v.x = smem[(h * 4 + 0) * kSStride + so];
    v.y = smem[(h * 4 + 1) * kSStride + so];
    v.z = smem[(h * 4 + 2) * kSStride + so];
    v.w = smem[(h * 4 + 3) * kSStride + so];
    *reinterpret_cast<float4*>(p.state + (size_t)f * p.n_pad + env0 + h * 4) = v;
  }
}

// test hook: evaluates the rotation coefficients for an array of angles (mode 0: production path, 1: force the
// double-double slow path, 2: force the libm-style sincos path)
__global__ void rotz_debug_kernel(const float* radians, int n, int mode, float* c_out, float* s_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c, s;
  if (mode == 1) {
    sincos_dd_small((double)radians[i], c, s);
  } else if (mode == 2) {
    double sd, cd;
    sincos((double)radians[i], &sd, &cd);
    c = (float)cd;
    s = (float)sd;
  } else {
    rotz(radians[i], c, s);
  }
  c_out[i] = c;
  s_out[i] = s;
}

// ---------------------------------------------------------------- host side
cudaError_t launch_rotz_debug(const float* radians, int n, int mode, float* c_out, float* s_out, cudaStream_t stream) {
  rotz_debug_kernel<<<(n + 255) / 256, 256, 0, stream>>>(radians, n, mode, c_out, s_out);
  return cudaGetLastError();
}

cudaError_t upload_materials(const Material* table, int count) {
  return cudaMemcpyToSymbol(c_materials, table, sizeof(Material) * count);
}

cudaError_t upload_scene_constants(const float* init_state92, const float* floor10) {
  cudaError_t e = cudaMemcpyToSymbol(c_init_state, init_state92, sizeof(float) * kStateFloats);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_floor, floor10, sizeof(float) * 10);
}

cudaError_t launch_physics(const PhysicsParams& p, int lanes_per_env, bool trace, cudaStream_t stream) {
  const int grid = p.n_pad / kEnvsPerCta;
  if (lanes_per_env == 32) {
    if (trace)
      physics_step_kernel<32, true><<<grid, kEnvsPerCta * 32, 0, stream>>>(p);
    else
      physics_step_kernel<32, false><<<grid, kEnvsPerCta * 32, 0, stream>>>(p);
  } else {
    if (trace)
      physics_step_kernel<16, true><<<grid, kEnvsPerCta * 16, 0, stream>>>(p);
    else
      physics_step_kernel<16, false><<<grid, kEnvsPerCta * 16, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace wb
