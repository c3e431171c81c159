/*
 * walker_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A strict-binary32 restatement, in plain C, of the hot path of
 * De-Rosa/PPO-BipedalWalker (C#/MonoGame).  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * can check (and time) the CUDA path against the reference's arithmetic.
 * Nothing under ppo-bipedalwalker_b200/ may include, link or call it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and
 * no C#/.NET/MonoGame toolchain exists in the build image, so this restatement
 * cannot be checked against reference outputs.  It is pinned instead against
 * (i) hand-derived known answers from source constants (SURVEY.md Appendix D,
 * tests/golden/kat_appendix_d.json) and (ii) an independent NumPy-float32
 * restatement (oracle/np_oracle.py), bit for bit.
 *
 * Third-party arithmetic not vendored in the reference (MonoGame.Framework
 * Vector2/Matrix, version unpinned; .NET Math/MathF) is restated from the
 * published upstream formulas listed in SURVEY.md Appendix C.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#ifndef WALKER_ORACLE_H
#define WALKER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- canonical per-env state record (shared vocabulary with include/walker_b200.h) ----
 * 92 floats:
 *   [ 0..57]  vertices, body order LLL(6) LLU(6) Body(5) RLL(6) RLU(6), (x,y) interleaved
 *   [58..67]  cached centroids (x,y) x 5 bodies        (Skeleton._centroid, Skeleton.cs:14)
 *   [68..77]  linear velocities (x,y) x 5              (RigidBody._linearVelocity, RigidBody.cs:30)
 *   [78..82]  angular velocities x 5                   (RigidBody._angularVelocity, RigidBody.cs:31)
 *   [83..87]  tracked angles x 5                       (RigidBody._angle, RigidBody.cs:33)
 *   [88..91]  joint _currentTorque x 4                 (Joint.cs:18)
 * 2 int32:
 *   flags: bit0..4 Collided per body (RigidBody.cs:21), bit5 Walker.Terminal (Walker.cs:23),
 *          bit6 floor-first list order (after first Reset, Walker.cs:212-223)
 *   steps: Environment._steps (Environment.cs:29)
 */
#define WO_STATE_FLOATS 92
#define WO_STATE_INTS 2
#define WO_OBS 12
#define WO_ACT 4

enum { WO_LLL = 0, WO_LLU = 1, WO_BODY = 2, WO_RLL = 3, WO_RLU = 4, WO_FLOOR = 5 };
#define WO_FLAG_TERMINAL (1 << 5)
#define WO_FLAG_FLOOR_FIRST (1 << 6)

/* Materials/{Ice..SuperRubber}.cs:7-9 */
typedef struct {
  float inverse_mass, restitution, friction;
} wo_material;
enum { WO_ICE = 0, WO_WOOD, WO_PAPER, WO_TITANIUM, WO_CARPET, WO_RUBBER, WO_METAL, WO_SUPERRUBBER, WO_NUM_MATERIALS };
const wo_material* wo_builtin_material(int id);

/* one record per ordered candidate pair per substep (RigidBody.ResolveCollisions, RigidBody.cs:66-96).
 * slot = LLL:{0,1} LLU:{2,3} Body:{4} RLL:{5,6} RLU:{7,8}; second index = list order of the candidate. */
typedef struct {
  int32_t other;     /* body id of the candidate (-1: slot unused)                              */
  int32_t aabb;      /* Skeleton.IsColliding result                                              */
  int32_t sat;       /* SATCollision.IsColliding result (0 when aabb == 0)                       */
  int32_t axis;      /* index of the winning axis in [A edges..., B edges...] (-1 when !sat)     */
  float nx, ny;      /* final (oriented) normal                                                  */
  float depth;
  int32_t ncontacts; /* 0..2                                                                     */
  float c0x, c0y, c1x, c1y;
} wo_pair_trace;
#define WO_PAIR_SLOTS 9

typedef struct {
  int32_t active; /* gap >= 0.1f (Joint.cs:35) */
  float depth;
} wo_joint_trace;

typedef struct wo_env wo_env;

int wo_env_sizeof(void);
/* Environment ctor (Environment.cs:39-51): walker first, floor last. */
void wo_env_init(wo_env* e, wo_material floor, wo_material walker);
/* Environment.Reset (Environment.cs:167-173) + Walker.Reset (Walker.cs:212-223): floor first afterwards. */
void wo_env_reset(wo_env* e);
/* Matrix.Clip(action, 1, -1) + Walker.TakeActions (Environment.cs:78, Walker.cs:66-75, Joint.cs:56-61). */
void wo_env_take_actions(wo_env* e, const float* actions4);
/* Environment.StepObjects (Environment.cs:126-143). trace may be NULL; else
 * pair_trace[iterations][9], joint_trace[iterations][4]. */
void wo_env_step_objects(wo_env* e, float dt, int iterations, wo_pair_trace* pair_trace, wo_joint_trace* joint_trace);
/* _steps++ (Environment.cs:72) happens in wo_env_step; this is the tail of Environment.Step
 * (Environment.cs:101-121): Walker.Update, reward, terminal, GetState. */
void wo_env_observe(wo_env* e, int max_timesteps, float* obs12, float* reward, uint8_t* done);
/* Walker.GetState only (Walker.cs:132-152) */
void wo_env_get_obs(const wo_env* e, float* obs12);
/* one full Environment.Update body minus the policy: steps++, TakeActions, Step; auto_reset -> Reset + InitialState. */
void wo_env_step(wo_env* e, const float* actions4, float dt, int iterations, int max_timesteps, int auto_reset,
                 float* obs12, float* reward, uint8_t* done);
void wo_env_get_state(const wo_env* e, float* f92, int32_t* i2);
void wo_env_set_state(wo_env* e, const float* f92, const int32_t* i2);

/* batch helpers (OpenMP over independent envs; AoS array of wo_env of wo_env_sizeof() bytes each) */
void wo_batch_step(void* envs, int n, const float* actions /*[n][4]*/, float dt, int iterations, int max_timesteps,
                   int auto_reset, float* obs /*[n][12]*/, float* reward, uint8_t* done, int nthreads);

/* stand-alone pieces exposed for unit tests (polygons as interleaved xy, n <= 8) */
int wo_sat(const float* a, int na, const float* b, int nb, const float* ca, const float* cb, float* normal2, float* depth,
           int* axis);
int wo_contacts(const float* a, int na, const float* b, int nb, const float* normal2, float* pts4);
void wo_pole_from_size(float cx, float cy, float size, float* verts12, float* centroid2);
void wo_rotz(float radians, float* c, float* s);

/* ------------------------------------------------------------------ PPO ---- */
/* Layer kinds (PPOAgent.ParseLayers, PPOAgent.cs:96-143) */
enum { WO_DENSE = 0, WO_RELU = 1, WO_LEAKYRELU = 2, WO_TANH = 3 };

typedef struct {
  float alpha, beta1, beta2, adam_epsilon; /* Hyperparameters.cs:104-107 */
  float epsilon;                           /* clip, :120 */
  float log_std;                           /* :121 */
  float gamma, lambda;                     /* :116-117 */
  int32_t batch_size;                      /* :111 */
} wo_hyper;
void wo_hyper_defaults(wo_hyper* hp);

typedef struct wo_net wo_net;
/* kinds[n], sizes[n] (output size for dense, ignored otherwise) */
wo_net* wo_net_create(int input_size, const int32_t* kinds, const int32_t* sizes, int nlayers);
void wo_net_destroy(wo_net* net);
int wo_net_num_params(const wo_net* net);
int wo_net_output_size(const wo_net* net);
/* flat order: per dense layer, W[out][in] row-major then b[out] (DenseLayer.Save, DenseLayer.cs:73-79) */
void wo_net_set_params(wo_net* net, const float* flat);
void wo_net_get_params(const wo_net* net, float* flat);
void wo_net_get_grads(const wo_net* net, float* flat);
void wo_net_get_adam(const wo_net* net, float* m_flat, float* v_flat, int32_t* iters /*per dense layer*/);
void wo_net_set_adam(wo_net* net, const float* m_flat, const float* v_flat, const int32_t* iters);
/* NeuralNetwork.FeedForward (NeuralNetwork.cs:52-64) */
void wo_net_forward(wo_net* net, const float* x, float* y, int cache);
/* NeuralNetwork.FeedBack (:67-82), Zero (:179-185), Optimise (:85-91) */
void wo_net_feedback(wo_net* net, const float* grad_out);
void wo_net_zero(wo_net* net);
void wo_net_optimise(wo_net* net, const wo_hyper* hp);

/* NormalDistribution.LogProbabilityDensity (NormalDistribution.cs:24-32) */
float wo_log_prob(float mean, float std, float action);
/* NormalDistribution.BoxMullerTransform (:12-19) with the two uniforms injected */
float wo_box_muller(float mean, float std, float u1, float u2);
/* PPOAgent.SampleActions (PPOAgent.cs:381-398) with injected uniforms u[act][2] */
void wo_sample_actions(wo_net* actor, const wo_hyper* hp, const float* state, const float* u, float* action, float* logp,
                       float* mean);
/* PPOAgent.Train(Batch) (PPOAgent.cs:218-346): returns number of samples skipped; optimise!=0 runs Adam. */
int wo_ppo_train_batch(wo_net* actor, wo_net* critic, const wo_hyper* hp, int n, const float* states, const float* actions,
                       const float* old_logp, const float* advantages, const float* returns, int optimise, float* critic_loss,
                       float* actor_loss);
/* per-sample dL/dmu[act] and dL/dV (before backprop), for kernel unit tests. returns 0 if the sample is skipped. */
int wo_ppo_sample_grad(const wo_hyper* hp, int act, const float* mean, const float* action, const float* old_logp, float adv,
                       float value, float ret, float* g_mu, float* g_v);
/* PPOAgent.MonteCarloReturn/MonteCarloAdvantages (:475-498), GAE (:414-444), Normalize (:461-472) */
void wo_mc_returns(const float* rewards, const float* values, int n, float gamma, float* returns, float* advantages);
void wo_gae(const float* rewards, const float* values, int n, float gamma, float lambda, float* returns, float* advantages);
void wo_normalize(float* list, int n, float epsilon);

#ifdef __cplusplus
}
#endif
#endif
