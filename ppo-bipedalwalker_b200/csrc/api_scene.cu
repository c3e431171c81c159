// api_scene.cu -- C ABI for general scenes (include/walker_b200.h, wb_scene_*): the reference's IObject plugin surface
// (Objects/IObject.cs:7-10; Objects/RigidBodies/{Square,Triangle,Hexagon,Pole,Hull,Joint}.cs) for N lockstep copies of an
// arbitrary list of convex polygons and joints.  All compute is in physics_scene.cu; there is no CPU path.
#include <new>
#include <vector>

#include "common.h"
#include "physics.cuh"

using namespace wb;

struct wb_scene {
  int32_t n = 0, n_pad = 0;
  int32_t iterations = 50;
  int32_t rows = 0;  // state floats per copy
  SceneConst sc{};
  cudaStream_t stream = nullptr;
  float* d_state = nullptr;     // [rows][n_pad]
  int32_t* d_collided = nullptr;
  float* d_torques = nullptr;   // [n][J]
  int64_t launches = 0;
};

// strict-fp32 host arithmetic (volatile stores keep every intermediate in binary32)
static inline float f_add(float a, float b) { volatile float r = a + b; return r; }
static inline float f_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float f_div(float a, float b) { volatile float r = a / b; return r; }

extern "C" {

int32_t wb_scene_create(int32_t n_envs, const wb_body_desc* bodies, int32_t n_bodies, const float* vertices_xy, const wb_joint_desc* joints,
                        int32_t n_joints, int32_t iterations, wb_scene** out) {
  WB_REQUIRE(out, "out is null");
  *out = nullptr;
  WB_REQUIRE(n_envs > 0 && bodies && vertices_xy, "bad argument");
  WB_REQUIRE(n_bodies > 0 && n_bodies <= kSceneMaxBodies, "a scene holds 1..16 bodies");
  WB_REQUIRE(n_joints >= 0 && n_joints <= kSceneMaxJoints && (n_joints == 0 || joints), "a scene holds 0..16 joints");
  WB_REQUIRE(iterations > 0 && iterations < 200, "iterations must be in (0, 200)");  // Hyperparameters.cs:189-217
  if (int32_t rc = require_device()) return rc;
  wb_scene* s = new (std::nothrow) wb_scene();
  WB_REQUIRE(s, "out of host memory");
  s->n = n_envs;
  s->n_pad = (n_envs + kEnvPad - 1) / kEnvPad * kEnvPad;
  s->iterations = iterations;
  SceneConst& c = s->sc;
  c.n_bodies = n_bodies;
  c.n_joints = n_joints;
  int off = 0;
  for (int b = 0; b < n_bodies; b++) {
    const wb_body_desc& d = bodies[b];
    if (d.n_vertices < 3 || d.n_vertices > kSceneMaxVerts) {
      delete s;
      return fail(WB_ERR_INVALID, "body %d: a polygon has 3..16 vertices", b);
    }
    float im, e, mu;
    if (wb_material_get(d.material, &im, &e, &mu) != WB_OK) {
      delete s;
      return fail(WB_ERR_INVALID, "body %d uses an unregistered material id %d", b, d.material);
    }
    c.n_verts[b] = d.n_vertices;
    c.vert_offset[b] = off;
    off += d.n_vertices;
    c.is_static[b] = d.is_static != 0;
    c.is_floor[b] = d.is_floor != 0;
    c.assoc[b] = d.associated_mask;
    // RigidBody ctor, RigidBody.cs:36-50: static => inverse mass = inverse inertia = 0; else 0.001 * inverse mass (or the override)
    c.inv_mass[b] = d.is_static ? 0.0f : im;
    c.inv_inertia[b] = d.is_static ? 0.0f : (d.inverse_inertia >= 0.0f ? d.inverse_inertia : f_mul(0.001f, im));
    c.restitution[b] = e;
    c.friction[b] = mu;
    c.accel_x[b] = d.accel_x;
    c.accel_y[b] = d.accel_y;
  }
  c.total_verts = off;
  for (int k = 0; k < n_joints; k++) {
    const wb_joint_desc& j = joints[k];
    const bool ok = j.body_a >= 0 && j.body_a < n_bodies && j.body_b >= 0 && j.body_b < n_bodies && j.vertex_a >= 0 &&
                    j.vertex_a < c.n_verts[j.body_a] && j.vertex_b >= 0 && j.vertex_b < c.n_verts[j.body_b];
    if (!ok) {
      delete s;
      return fail(WB_ERR_INVALID, "joint %d references a body / vertex that does not exist", k);
    }
    c.joint_a[k] = j.body_a;
    c.joint_ia[k] = j.vertex_a;
    c.joint_b[k] = j.body_b;
    c.joint_ib[k] = j.vertex_b;
  }
  if (scene_smem_bytes(c) > 200 * 1024) {
    delete s;
    return fail(WB_ERR_UNSUPPORTED, "scene too large for one SM's shared memory (%zu bytes per 32 copies)", scene_smem_bytes(c));
  }
  s->rows = 2 * c.total_verts + 4 * n_bodies + 2 * n_bodies + n_joints;
  // one copy of the initial state: vertices, Skeleton.FindCentroid (sum, then multiply by 1f / count, Skeleton.cs:100-113), zeros
  std::vector<float> one((size_t)s->rows, 0.0f);
  for (int i = 0; i < 2 * c.total_verts; i++) one[i] = vertices_xy[i];
  for (int b = 0; b < n_bodies; b++) {
    float sx = 0.0f, sy = 0.0f;
    for (int i = 0; i < c.n_verts[b]; i++) {
      sx = f_add(sx, vertices_xy[2 * (c.vert_offset[b] + i)]);
      sy = f_add(sy, vertices_xy[2 * (c.vert_offset[b] + i) + 1]);
    }
    const float factor = f_div(1.0f, (float)c.n_verts[b]);
    one[2 * c.total_verts + 2 * b] = f_mul(sx, factor);
    one[2 * c.total_verts + 2 * b + 1] = f_mul(sy, factor);
  }
  std::vector<float> all((size_t)s->rows * s->n_pad);
  for (int r = 0; r < s->rows; r++)
    for (int i = 0; i < s->n_pad; i++) all[(size_t)r * s->n_pad + i] = one[r];
  cudaError_t e = cudaMalloc(&s->d_state, sizeof(float) * all.size());
  if (e == cudaSuccess) e = cudaMalloc(&s->d_collided, sizeof(int32_t) * s->n_pad);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_torques, sizeof(float) * (size_t)s->n * (n_joints > 0 ? n_joints : 1));
  if (e == cudaSuccess) e = cudaMemcpy(s->d_state, all.data(), sizeof(float) * all.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(s->d_collided, 0, sizeof(int32_t) * s->n_pad);
  if (e != cudaSuccess) {  // nothing leaks behind a failed create
    wb_scene_destroy(s);
    return fail(WB_ERR_CUDA, "wb_scene_create: %s", cudaGetErrorString(e));
  }
  *out = s;
  return WB_OK;
}

int32_t wb_scene_destroy(wb_scene* s) {
  if (!s) return WB_OK;
  cudaFree(s->d_state);
  cudaFree(s->d_collided);
  cudaFree(s->d_torques);
  delete s;
  return WB_OK;
}

int32_t wb_scene_set_stream(wb_scene* s, void* cuda_stream) {
  WB_REQUIRE(s, "scene is null");
  s->stream = (cudaStream_t)cuda_stream;
  return WB_OK;
}

int32_t wb_scene_state_floats(const wb_scene* s, int32_t* per_copy_out) {
  WB_REQUIRE(s && per_copy_out, "null argument");
  *per_copy_out = s->rows;
  return WB_OK;
}

int32_t wb_scene_set_state(wb_scene* s, const float* state_host, const int32_t* collided_host) {
  WB_REQUIRE(s && state_host, "null argument");
  WB_CUDA(cudaMemcpy2DAsync(s->d_state, sizeof(float) * s->n_pad, state_host, sizeof(float) * s->n, sizeof(float) * s->n, s->rows,
                            cudaMemcpyHostToDevice, s->stream));
  if (collided_host) WB_CUDA(cudaMemcpyAsync(s->d_collided, collided_host, sizeof(int32_t) * s->n, cudaMemcpyHostToDevice, s->stream));
  WB_CUDA(cudaStreamSynchronize(s->stream));
  return WB_OK;
}

int32_t wb_scene_get_state(wb_scene* s, float* state_host, int32_t* collided_host) {
  WB_REQUIRE(s && state_host, "null argument");
  WB_CUDA(cudaMemcpy2DAsync(state_host, sizeof(float) * s->n, s->d_state, sizeof(float) * s->n_pad, sizeof(float) * s->n, s->rows,
                            cudaMemcpyDeviceToHost, s->stream));
  if (collided_host) WB_CUDA(cudaMemcpyAsync(collided_host, s->d_collided, sizeof(int32_t) * s->n, cudaMemcpyDeviceToHost, s->stream));
  WB_CUDA(cudaStreamSynchronize(s->stream));
  return WB_OK;
}

static int32_t scene_launch(wb_scene* s, const float* d_torques, float dt, int iterations) {
  SceneParams p{};
  p.state = s->d_state;
  p.collided = s->d_collided;
  p.torques = d_torques;
  p.n = s->n;
  p.n_pad = s->n_pad;
  p.dt = dt;
  p.iterations = iterations;
  WB_CUDA(launch_scene(p, s->sc, s->stream));
  s->launches++;
  return WB_OK;
}

int32_t wb_scene_set_torques(wb_scene* s, const float* torques_host) {
  WB_REQUIRE(s && torques_host, "null argument");
  WB_REQUIRE(s->sc.n_joints > 0, "the scene has no joints");
  WB_CUDA(cudaMemcpyAsync(s->d_torques, torques_host, sizeof(float) * (size_t)s->n * s->sc.n_joints, cudaMemcpyHostToDevice, s->stream));
  if (int32_t rc = scene_launch(s, s->d_torques, 0.0f, 0)) return rc;
  WB_CUDA(cudaStreamSynchronize(s->stream));
  return WB_OK;
}

int32_t wb_scene_step_objects(wb_scene* s, float delta_time) {
  WB_REQUIRE(s, "scene is null");
  if (int32_t rc = scene_launch(s, nullptr, delta_time, s->iterations)) return rc;
  WB_CUDA(cudaStreamSynchronize(s->stream));
  return WB_OK;
}

}  // extern "C"
