/*
 * walker_oracle_ppo.c -- CPU ORACLE (test infrastructure, NOT the product). PARITY UNPINNED, see walker_oracle.h.
 *
 * Restatement of the reference's PPO arithmetic: the jagged-array Matrix library semantics
 * (Walker/PPO/Matrix.cs), the layer stack (Walker/PPO/Network/ files), the per-sample clipped-surrogate
 * gradient (Walker/PPO/PPOAgent.cs:218-346) and Adam (Walker/PPO/Network/DenseLayer.cs:125-159).
 * fp32 throughout, sums accumulated left to right exactly as Matrix.Multiply (Matrix.cs:604-616) does.
 * Transcendentals (MathF.Exp/Log/Tanh/Sin/Sqrt, Math.Pow) map to the platform libm, as in .NET.
 */
#include "walker_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LAYERS 16

typedef struct {
  int kind;
  int in, out;
  /* dense only */
  float *W, *b;   /* DenseLayer._weights [out][in], _biases [out] */
  float *dW, *db; /* _derivativeLossWrt* */
  float *mW, *mb, *vW, *vb;
  int iteration;
} layer_t;

struct wo_net {
  int input_size;
  int nlayers;
  layer_t layers[MAX_LAYERS];
  float* cache[MAX_LAYERS]; /* NeuralNetwork._cache: the INPUT of each layer */
  int cache_valid;
  int max_width;
};

void wo_hyper_defaults(wo_hyper* hp) { /* Hyperparameters.cs:104-121 */
  hp->alpha = 0.001f;
  hp->beta1 = 0.9f;
  hp->beta2 = 0.999f;
  hp->adam_epsilon = 1e-8f;
  hp->epsilon = 0.3f;
  hp->log_std = -1.0f;
  hp->gamma = 0.9f;
  hp->lambda = 0.95f;
  hp->batch_size = 64;
}

wo_net* wo_net_create(int input_size, const int32_t* kinds, const int32_t* sizes, int nlayers) {
  if (nlayers > MAX_LAYERS) return NULL;
  wo_net* net = (wo_net*)calloc(1, sizeof(wo_net));
  net->input_size = input_size;
  net->nlayers = nlayers;
  int width = input_size;
  net->max_width = width;
  for (int i = 0; i < nlayers; i++) {
    layer_t* L = &net->layers[i];
    L->kind = kinds[i];
    L->in = width;
    if (L->kind == WO_DENSE) {
      L->out = sizes[i];
      size_t nw = (size_t)L->in * L->out;
      L->W = (float*)calloc(nw, 4);
      L->dW = (float*)calloc(nw, 4);
      L->mW = (float*)calloc(nw, 4);
      L->vW = (float*)calloc(nw, 4);
      L->b = (float*)calloc(L->out, 4);
      L->db = (float*)calloc(L->out, 4);
      L->mb = (float*)calloc(L->out, 4);
      L->vb = (float*)calloc(L->out, 4);
      width = L->out;
    } else {
      L->out = width;
    }
    net->cache[i] = (float*)calloc(L->in, 4);
    if (width > net->max_width) net->max_width = width;
  }
  return net;
}

void wo_net_destroy(wo_net* net) {
  if (!net) return;
  for (int i = 0; i < net->nlayers; i++) {
    layer_t* L = &net->layers[i];
    free(L->W);
    free(L->b);
    free(L->dW);
    free(L->db);
    free(L->mW);
    free(L->mb);
    free(L->vW);
    free(L->vb);
    free(net->cache[i]);
  }
  free(net);
}

int wo_net_num_params(const wo_net* net) {
  int n = 0;
  for (int i = 0; i < net->nlayers; i++)
    if (net->layers[i].kind == WO_DENSE) n += net->layers[i].in * net->layers[i].out + net->layers[i].out;
  return n;
}

int wo_net_output_size(const wo_net* net) { return net->layers[net->nlayers - 1].out; }

#define FOR_DENSE(net, L) \
  for (int _i = 0; _i < (net)->nlayers; _i++) \
    for (layer_t* L = (layer_t*)&(net)->layers[_i]; L && L->kind == WO_DENSE; L = NULL)

void wo_net_set_params(wo_net* net, const float* flat) {
  FOR_DENSE(net, L) {
    memcpy(L->W, flat, (size_t)L->in * L->out * 4);
    flat += L->in * L->out;
    memcpy(L->b, flat, (size_t)L->out * 4);
    flat += L->out;
  }
}
void wo_net_get_params(const wo_net* net, float* flat) {
  FOR_DENSE(net, L) {
    memcpy(flat, L->W, (size_t)L->in * L->out * 4);
    flat += L->in * L->out;
    memcpy(flat, L->b, (size_t)L->out * 4);
    flat += L->out;
  }
}
void wo_net_get_grads(const wo_net* net, float* flat) {
  FOR_DENSE(net, L) {
    memcpy(flat, L->dW, (size_t)L->in * L->out * 4);
    flat += L->in * L->out;
    memcpy(flat, L->db, (size_t)L->out * 4);
    flat += L->out;
  }
}
void wo_net_get_adam(const wo_net* net, float* m, float* v, int32_t* iters) {
  FOR_DENSE(net, L) {
    size_t nw = (size_t)L->in * L->out;
    memcpy(m, L->mW, nw * 4);
    memcpy(v, L->vW, nw * 4);
    m += nw;
    v += nw;
    memcpy(m, L->mb, (size_t)L->out * 4);
    memcpy(v, L->vb, (size_t)L->out * 4);
    m += L->out;
    v += L->out;
    *iters++ = L->iteration;
  }
}
void wo_net_set_adam(wo_net* net, const float* m, const float* v, const int32_t* iters) {
  FOR_DENSE(net, L) {
    size_t nw = (size_t)L->in * L->out;
    memcpy(L->mW, m, nw * 4);
    memcpy(L->vW, v, nw * 4);
    m += nw;
    v += nw;
    memcpy(L->mb, m, (size_t)L->out * 4);
    memcpy(L->vb, v, (size_t)L->out * 4);
    m += L->out;
    v += L->out;
    L->iteration = *iters++;
  }
}

/* .NET MathF.Max: IEEE maximum */
static inline float mathf_max(float a, float b) {
  if (a != b) {
    if (!isnan(a)) return b < a ? a : b;
    return a;
  }
  return signbit(b) ? a : b;
}

/* ActivationLayer.cs:32-73 */
static float act_fwd(int kind, float x) {
  switch (kind) {
    case WO_RELU: return mathf_max(0.0f, x);
    case WO_LEAKYRELU: return mathf_max(0.2f * x, x);
    case WO_TANH: return tanhf(x);
  }
  return x;
}
static float act_bwd(int kind, float x) {
  switch (kind) {
    case WO_RELU: return x < 0.0f ? 0.0f : 1.0f;
    case WO_LEAKYRELU: return x < 0.0f ? 0.2f : 1.0f;
    case WO_TANH: return (1.0f - (tanhf(x) * tanhf(x)));
  }
  return 1.0f;
}

/* NeuralNetwork.FeedForward (NeuralNetwork.cs:52-64); DenseLayer.FeedForward (DenseLayer.cs:82-98):
 * result = W * x (Matrix.Multiply: sum starts at 0, adds left to right) then result += b. */
void wo_net_forward(wo_net* net, const float* x, float* y, int cache) {
  float bufA[1024], bufB[1024];
  float* cur = bufA;
  float* nxt = bufB;
  memcpy(cur, x, (size_t)net->input_size * 4);
  for (int li = 0; li < net->nlayers; li++) {
    layer_t* L = &net->layers[li];
    if (cache) memcpy(net->cache[li], cur, (size_t)L->in * 4);
    if (L->kind == WO_DENSE) {
      for (int o = 0; o < L->out; o++) {
        float sum = 0.0f;
        const float* w = L->W + (size_t)o * L->in;
        for (int i = 0; i < L->in; i++) sum += w[i] * cur[i];
        nxt[o] = sum + L->b[o];
      }
    } else {
      for (int o = 0; o < L->out; o++) nxt[o] = act_fwd(L->kind, cur[o]);
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  if (cache) net->cache_valid = 1;
  memcpy(y, cur, (size_t)net->layers[net->nlayers - 1].out * 4);
}

/* NeuralNetwork.FeedBack (:67-82); DenseLayer.FeedBack (DenseLayer.cs:103-120):
 *   db += Flatten(g)  -> db[o] = db[o] + (0 + g[o])
 *   dW += g * x^T     -> dW[o][i] = dW[o][i] + (0 + g[o]*x[i])
 *   g  <- W^T * g     -> sum over o, left to right, from 0
 * ActivationLayer.FeedBack (ActivationLayer.cs:18-21): g[o] * f'(cached input[o]) */
void wo_net_feedback(wo_net* net, const float* grad_out) {
  float bufA[1024], bufB[1024];
  float* g = bufA;
  float* ng = bufB;
  memcpy(g, grad_out, (size_t)net->layers[net->nlayers - 1].out * 4);
  for (int li = net->nlayers - 1; li >= 0; li--) {
    layer_t* L = &net->layers[li];
    const float* x = net->cache[li];
    if (L->kind == WO_DENSE) {
      for (int o = 0; o < L->out; o++) {
        float flat = 0.0f;
        flat += g[o];
        L->db[o] = L->db[o] + flat;
      }
      for (int o = 0; o < L->out; o++)
        for (int i = 0; i < L->in; i++) {
          float prod = 0.0f;
          prod += g[o] * x[i];
          L->dW[(size_t)o * L->in + i] = L->dW[(size_t)o * L->in + i] + prod;
        }
      for (int i = 0; i < L->in; i++) {
        float sum = 0.0f;
        for (int o = 0; o < L->out; o++) sum += L->W[(size_t)o * L->in + i] * g[o];
        ng[i] = sum;
      }
    } else {
      for (int o = 0; o < L->out; o++) ng[o] = g[o] * act_bwd(L->kind, x[o]);
    }
    float* t = g;
    g = ng;
    ng = t;
  }
}

void wo_net_zero(wo_net* net) {
  FOR_DENSE(net, L) {
    memset(L->dW, 0, (size_t)L->in * L->out * 4);
    memset(L->db, 0, (size_t)L->out * 4);
  }
}

/* DenseLayer.Adam, DenseLayer.cs:125-159 */
static void adam_array(float* w, const float* g, float* m, float* v, size_t n, const wo_hyper* hp, int iteration) {
  float one_m_b1 = 1.0f - hp->beta1;
  float one_m_b2 = 1.0f - hp->beta2;
  float corr1 = (float)(1.0 - pow((double)hp->beta1, (double)iteration));
  float corr2 = (float)(1.0 - pow((double)hp->beta2, (double)iteration));
  for (size_t i = 0; i < n; i++) {
    m[i] = (one_m_b1 * g[i]) + (hp->beta1 * m[i]);
    v[i] = (hp->beta2 * v[i]) + (one_m_b2 * (g[i] * g[i]));
    float mhat = m[i] / corr1;
    float vhat = v[i] / corr2;
    float denom = sqrtf(vhat) + hp->adam_epsilon;
    if (denom == 0.0f) return; /* HadamardDivision throws -> caught, layer left as is (DenseLayer.cs:154-157) */
    w[i] = w[i] - (hp->alpha * (mhat / denom));
  }
}

void wo_net_optimise(wo_net* net, const wo_hyper* hp) {
  FOR_DENSE(net, L) {
    L->iteration += 1;
    adam_array(L->W, L->dW, L->mW, L->vW, (size_t)L->in * L->out, hp, L->iteration);
    adam_array(L->b, L->db, L->mb, L->vb, (size_t)L->out, hp, L->iteration);
  }
}

/* NormalDistribution.LogProbabilityDensity, NormalDistribution.cs:24-32 */
float wo_log_prob(float mean, float std, float action) {
  float fraction = (action - mean) / std;
  fraction *= fraction;
  fraction /= 2.0f;
  return -logf(std) - logf(sqrtf(2.0f * 3.14159274f)) - fraction;
}

/* NormalDistribution.BoxMullerTransform, NormalDistribution.cs:12-19 */
float wo_box_muller(float mean, float std, float u1, float u2) {
  if (u1 == 0.0f) u1 = 1.0f;
  float standard_normal = sqrtf(-2.0f * logf(u1)) * sinf(2.0f * 3.14159274f * u2);
  return mean + (std * standard_normal);
}

/* PPOAgent.SampleActions + GetStandardDeviations, PPOAgent.cs:367-398 */
void wo_sample_actions(wo_net* actor, const wo_hyper* hp, const float* state, const float* u, float* action, float* logp,
                       float* mean_out) {
  int act = wo_net_output_size(actor);
  float mean[64];
  wo_net_forward(actor, state, mean, 0);
  float std = expf(hp->log_std);
  for (int k = 0; k < act; k++) {
    action[k] = wo_box_muller(mean[k], std, u[2 * k], u[2 * k + 1]);
    logp[k] = wo_log_prob(mean[k], std, action[k]);
    if (mean_out) mean_out[k] = mean[k];
  }
}

/* the per-sample part of PPOAgent.Train(Batch), PPOAgent.cs:248-326 */
int wo_ppo_sample_grad(const wo_hyper* hp, int act, const float* mean, const float* action, const float* old_logp, float adv,
                       float value, float ret, float* g_mu, float* g_v) {
  float std = expf(hp->log_std);
  float upper = 1.0f + hp->epsilon, lower = 1.0f - hp->epsilon;
  float bs = (float)hp->batch_size;
  float critic_loss = 2.0f * (value - ret);
  float tmp[64];
  for (int k = 0; k < act; k++) {
    float logp = wo_log_prob(mean[k], std, action[k]);
    float ratio = expf(logp - old_logp[k]);
    float clipped = ratio >= upper ? upper : (ratio <= lower ? lower : ratio); /* Matrix.Clip */
    float cra = clipped * adv;
    float ra = ratio * adv;
    float partA = (ra <= cra ? 1.0f : 0.0f) * adv;                          /* LessThan */
    float partB = (cra < ra ? 1.0f : 0.0f) * adv;                           /* LessThanNotEquals */
    float partC = (ratio >= lower && ratio <= upper) ? 1.0f : 0.0f;         /* InRange */
    float dclip = partA + (partB * partC);
    dclip = dclip * -1.0f;
    float pold = expf(old_logp[k]);
    if (pold == 0.0f) return 0; /* HadamardDivision throws -> sample skipped (PPOAgent.cs:286-290) */
    dclip = dclip / pold;
    float prob = expf(logp);
    float amm = action[k] - mean[k];
    float variance = std * std;
    if (variance == 0.0f) return 0;
    float fraction = amm / variance;
    float dmean = prob * fraction;
    tmp[k] = (dmean * dclip) / bs;
  }
  for (int k = 0; k < act; k++) g_mu[k] = tmp[k];
  *g_v = critic_loss / bs;
  return 1;
}

/* PPOAgent.Train(Batch), PPOAgent.cs:218-346 */
int wo_ppo_train_batch(wo_net* actor, wo_net* critic, const wo_hyper* hp, int n, const float* states, const float* actions,
                       const float* old_logp, const float* advantages, const float* returns, int optimise, float* critic_loss,
                       float* actor_loss) {
  int act = wo_net_output_size(actor);
  int sdim = actor->input_size;
  wo_net_zero(actor);
  wo_net_zero(critic);
  float avg_c = 0.0f, avg_a = 0.0f;
  int skipped = 0;
  for (int i = 0; i < n; i++) {
    const float* s = states + (size_t)i * sdim;
    float value;
    wo_net_forward(critic, s, &value, 1);
    float mean[64], g_mu[64], g_v;
    wo_net_forward(actor, s, mean, 1);
    if (!wo_ppo_sample_grad(hp, act, mean, actions + (size_t)i * act, old_logp + (size_t)i * act, advantages[i], value,
                            returns[i], g_mu, &g_v)) {
      skipped++;
      continue;
    }
    avg_c += g_v;
    float sum = 0.0f; /* Matrix.Average, Matrix.cs:588-602 */
    for (int k = 0; k < act; k++) sum += g_mu[k];
    avg_a += sum / (float)act;
    wo_net_feedback(critic, &g_v);
    wo_net_feedback(actor, g_mu);
  }
  if (optimise) {
    wo_net_optimise(critic, hp);
    wo_net_optimise(actor, hp);
  }
  if (critic_loss) *critic_loss = avg_c;
  if (actor_loss) *actor_loss = avg_a;
  return skipped;
}

/* PPOAgent.MonteCarloReturn + MonteCarloAdvantages, PPOAgent.cs:475-498 */
void wo_mc_returns(const float* rewards, const float* values, int n, float gamma, float* returns, float* advantages) {
  float g = 0.0f;
  for (int i = n - 1; i >= 0; i--) {
    g = rewards[i] + (g * gamma);
    returns[i] = g;
  }
  for (int i = 0; i < n; i++) advantages[i] = returns[i] - values[i];
}

/* PPOAgent.GeneralizedAdvantageEstimate + CalculateDelta, PPOAgent.cs:414-444 */
void wo_gae(const float* rewards, const float* values, int n, float gamma, float lambda, float* returns, float* advantages) {
  float next_gae = 0.0f, next_value = 0.0f;
  for (int i = n - 1; i >= 0; i--) {
    float cur = values[i];
    float delta = rewards[i] + (gamma * next_value) - cur;
    next_value = cur;
    /* NOTE: the reference never updates nextGae (PPOAgent.cs:422-431), so the lambda term is always 0 */
    float gae = delta + (gamma * lambda * next_gae);
    advantages[i] = gae;
    returns[i] = gae + values[i];
  }
}

/* PPOAgent.Normalize, PPOAgent.cs:461-472: List<float>.Average() accumulates in double; std via double pow/sqrt */
void wo_normalize(float* list, int n, float epsilon) {
  if (n == 0) return;
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc += (double)list[i];
  float mean = (float)(acc / (double)n);
  double ss = 0.0;
  for (int i = 0; i < n; i++) ss += pow((double)(list[i] - mean), 2.0);
  float std = (float)sqrt(ss / (double)n);
  for (int i = 0; i < n; i++) {
    list[i] -= mean;
    list[i] /= std + epsilon;
  }
}
