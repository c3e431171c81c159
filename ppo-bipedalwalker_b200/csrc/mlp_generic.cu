// mlp_generic.cu -- the PPO pipeline for ANY network the reference's DSL can describe (PPOAgent.ParseLayers, PPOAgent.cs:96-143):
// any sequence of dense layers (widths 1..128) and ReLU / LeakyReLU / TanH activations (ActivationLayer.cs:32-73), any state and
// action size.  The tensor-core kernel (mlp_tc.cu) and the register-blocked fp32 kernel (mlp.cu) are specialised for the
// reference's default topologies; a user-edited Hyperparameters.ActorNeuralNetwork / CriticNeuralNetwork lands here instead of
// falling off the GPU path.  General, not tuned: one CTA per tile of TS samples, the layer inputs the reference caches
// (NeuralNetwork._cache: the INPUT of every layer, NeuralNetwork.cs:52-64) live in shared memory, every dot product runs left to
// right like Matrix.Multiply (Matrix.cs:604-616) with the product and the sum rounded separately (no FMA: a pre-activation that
// lands within rounding distance of an activation's kink takes the same side as in the C# loop), the cross-sample sums of
// dW / db are formed per CTA in sample order and added to a per-CTA partial that the same thread always owns (deterministic),
// then reduced by reduce_partials_n.
//
//   forward    NeuralNetwork.FeedForward / DenseLayer.FeedForward (DenseLayer.cs:82-98)
//   sample     PPOAgent.SampleActions (PPOAgent.cs:381-398), Box-Muller sin branch (NormalDistribution.cs:12-19)
//   gradient   PPOAgent.Train(Batch) (PPOAgent.cs:218-342): per-dimension ratio / clip, HadamardDivision zero divisor => sample
//              skipped, losses; NeuralNetwork.FeedBack / DenseLayer.FeedBack (DenseLayer.cs:103-120) / ActivationLayer.FeedBack
#include "mlp.cuh"

namespace wb {

namespace {

__device__ __forceinline__ float gen_activate(int kind, float v) {
  if (kind == WB_RELU) return fmaxf(0.0f, v);             // MathF.Max(0, value)
  if (kind == WB_LEAKYRELU) return fmaxf(0.2f * v, v);    // MathF.Max(Alpha * value, value)
  return tanhf(v);                                        // MathF.Tanh
}
__device__ __forceinline__ float gen_derivative(int kind, float v) {
  if (kind == WB_RELU) return v < 0.0f ? 0.0f : 1.0f;
  if (kind == WB_LEAKYRELU) return v < 0.0f ? 0.2f : 1.0f;
  const float t = tanhf(v);
  return 1.0f - (t * t);
}

__device__ __forceinline__ uint4 gen_philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float gen_log_prob(float mean, float stdv, float action, float neg_log_std, float log_sqrt_2pi) {
  float fraction = (action - mean) / stdv;  // NormalDistribution.LogProbabilityDensity, NormalDistribution.cs:24-32
  fraction *= fraction;
  fraction /= 2.0f;
  return neg_log_std - log_sqrt_2pi - fraction;
}

// NeuralNetwork.FeedForward over a tile: cache row of sample s = cache + s * stride; layer l reads its input at L[l].cache_off and
// writes the next layer's input (the last layer writes at net.cache_floats: the network output)
__device__ void gen_forward(const GenNet& net, const float* __restrict__ params, float* cache, int stride, int ts) {
  for (int l = 0; l < net.n_layers; l++) {
    const GenLayer& L = net.L[l];
    const int out_off = (l + 1 < net.n_layers) ? net.L[l + 1].cache_off : net.cache_floats;
    if (L.kind == WB_DENSE) {
      const float* W = params + L.w_off;
      const float* b = params + L.b_off;
      for (int idx = threadIdx.x; idx < ts * L.out; idx += blockDim.x) {
        const int s = idx / L.out, o = idx - s * L.out;
        const float* x = cache + (size_t)s * stride + L.cache_off;
        const float* w = W + (size_t)o * L.in;
        float acc = 0.0f;
        for (int i = 0; i < L.in; i++) acc = __fadd_rn(acc, __fmul_rn(x[i], __ldg(w + i)));  // product and sum individually rounded, like the C# loop
        cache[(size_t)s * stride + out_off + o] = acc + __ldg(b + o);
      }
    } else {
      for (int idx = threadIdx.x; idx < ts * L.in; idx += blockDim.x) {
        const int s = idx / L.in, j = idx - s * L.in;
        cache[(size_t)s * stride + out_off + j] = gen_activate(L.kind, cache[(size_t)s * stride + L.cache_off + j]);
      }
    }
    __syncthreads();
  }
}

// NeuralNetwork.FeedBack over a tile: g0 holds dL/d(output) [ts][kGenMaxWidth]; g1 is the ping-pong partner.  The per-CTA partial
// sums of dW / db go to `partial` (flat parameter layout of this network).
__device__ void gen_backward(const GenNet& net, const float* __restrict__ params, const float* cache, int stride, int ts, float* g0, float* g1,
                             float* partial) {
  float* g = g0;
  float* gn = g1;
  for (int l = net.n_layers - 1; l >= 0; l--) {
    const GenLayer& L = net.L[l];
    if (L.kind == WB_DENSE) {
      const float* W = params + L.w_off;
      // db += g ; dW += g x^T  (DenseLayer.FeedBack, DenseLayer.cs:103-120); element (o, i), i == in: the bias
      for (int idx = threadIdx.x; idx < L.out * (L.in + 1); idx += blockDim.x) {
        const int o = idx / (L.in + 1), i = idx - o * (L.in + 1);
        float acc = 0.0f;
        for (int s = 0; s < ts; s++) {
          const float gs = g[s * kGenMaxWidth + o];
          acc = __fadd_rn(acc, (i < L.in) ? __fmul_rn(gs, cache[(size_t)s * stride + L.cache_off + i]) : gs);
        }
        float* dst = partial + (i < L.in ? L.w_off + o * L.in + i : L.b_off + o);
        *dst += acc;
      }
      // g <- W^T g (sum over the outputs from 0, Matrix.Multiply order); the reference also does this for the first layer, where
      // nothing reads the result
      if (l > 0) {
        for (int idx = threadIdx.x; idx < ts * L.in; idx += blockDim.x) {
          const int s = idx / L.in, i = idx - s * L.in;
          float acc = 0.0f;
          for (int o = 0; o < L.out; o++) acc = __fadd_rn(acc, __fmul_rn(__ldg(W + (size_t)o * L.in + i), g[s * kGenMaxWidth + o]));
          gn[s * kGenMaxWidth + i] = acc;
        }
        __syncthreads();
        float* t = g;
        g = gn;
        gn = t;
      } else {
        __syncthreads();
      }
    } else {
      for (int idx = threadIdx.x; idx < ts * L.in; idx += blockDim.x) {
        const int s = idx / L.in, j = idx - s * L.in;
        g[s * kGenMaxWidth + j] *= gen_derivative(L.kind, cache[(size_t)s * stride + L.cache_off + j]);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kGenThreads) ppo_generic_kernel(const GenParams p) {
  extern __shared__ __align__(16) float gsm[];
  const int ts_max = p.ts;
  const int strideA = p.actor.cache_floats + p.actor.output, strideC = p.critic.cache_floats + p.critic.output;
  float* cacheA = gsm;
  float* cacheC = cacheA + (size_t)ts_max * strideA;
  float* g0 = cacheC + (size_t)ts_max * strideC;
  float* g1 = g0 + (size_t)ts_max * kGenMaxWidth;
  float* gv = g1 + (size_t)ts_max * kGenMaxWidth;       // [ts] dL/dV per sample
  float* sums = gv + ts_max;                            // [4]: loss V, loss A, skipped
  const int tid = threadIdx.x;
  const int act = p.actor.output, sdim = p.actor.input;
  const float* paramsA = p.params;
  const float* paramsC = p.params + p.actor.n_params;
  const bool grad = p.mode == kModeGrad;
  float* partial = grad ? p.partials + (size_t)blockIdx.x * p.grad_floats : nullptr;
  if (grad) {
    for (int i = tid; i < p.grad_floats; i += blockDim.x) partial[i] = 0.0f;
    if (tid < 4) sums[tid] = 0.0f;
    __syncthreads();
  }
  const float stdv = expf(p.log_std);                   // PPOAgent.GetStandardDeviations, PPOAgent.cs:367-378
  const float neg_log_std = -logf(stdv), log_sqrt_2pi = logf(sqrtf(2.0f * 3.14159274f));
  const float upper = 1.0f + p.epsilon, lower = 1.0f - p.epsilon, variance = stdv * stdv;

  const int ntiles = (p.n + ts_max - 1) / ts_max;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int s0 = tile * ts_max;
    const int ts = min(ts_max, p.n - s0);
    for (int idx = tid; idx < ts * sdim; idx += blockDim.x) {
      const int s = idx / sdim, i = idx - s * sdim;
      const float v = p.states[(size_t)(s0 + s) * sdim + i];
      cacheA[(size_t)s * strideA + p.actor.L[0].cache_off + i] = v;
      cacheC[(size_t)s * strideC + p.critic.L[0].cache_off + i] = v;
    }
    __syncthreads();
    gen_forward(p.actor, paramsA, cacheA, strideA, ts);
    gen_forward(p.critic, paramsC, cacheC, strideC, ts);
    const float* mu_of = cacheA + p.actor.cache_floats;   // + s * strideA
    const float* v_of = cacheC + p.critic.cache_floats;   // + s * strideC

    if (!grad) {
      for (int idx = tid; idx < ts * act; idx += blockDim.x) {
        const int s = idx / act, k = idx - s * act;
        const size_t gs = (size_t)(s0 + s);
        const float mu = mu_of[(size_t)s * strideA + k];
        if (p.mean) p.mean[gs * act + k] = mu;
        if (p.mode == kModeSample || p.mode == kModeSamplePhilox) {
          float u1, u2;
          if (p.mode == kModeSample) {
            u1 = p.uniforms[(gs * act + k) * 2];
            u2 = p.uniforms[(gs * act + k) * 2 + 1];
          } else {
            const uint4 r = gen_philox4x32(make_uint4((uint32_t)gs, (uint32_t)k, (uint32_t)p.step, (uint32_t)(p.step >> 32)),
                                           make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
            u1 = (float)(r.x >> 8) * (1.0f / 16777216.0f);
            u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
          }
          if (u1 == 0.0f) u1 = 1.0f;  // NormalDistribution.BoxMullerTransform, NormalDistribution.cs:12-19
          const float z = sqrtf(-2.0f * logf(u1)) * sinf(2.0f * 3.14159274f * u2);
          const float a = mu + (stdv * z);
          p.out_actions[gs * act + k] = a;
          p.out_logp[gs * act + k] = gen_log_prob(mu, stdv, a, neg_log_std, log_sqrt_2pi);
        }
      }
      if (p.value)
        for (int s = tid; s < ts; s += blockDim.x) p.value[s0 + s] = v_of[(size_t)s * strideC];
      __syncthreads();
      continue;
    }

    // ---- per-sample clipped-surrogate gradient (PPOAgent.cs:232-326): one thread per sample, action dimensions in order
    for (int s = tid; s < ts; s += blockDim.x) {
      const size_t gs = (size_t)(s0 + s);
      const float adv = p.advantages[gs];
      const float value = v_of[(size_t)s * strideC];
      float gvs = (2.0f * (value - p.returns[gs])) / p.batch_size;
      bool skip = false;
      float sum = 0.0f;
      for (int k = 0; k < act; k++) {
        const float mu = mu_of[(size_t)s * strideA + k];
        const float a = p.actions[gs * act + k];
        const float lp_old = p.old_logp[gs * act + k];
        const float lp = gen_log_prob(mu, stdv, a, neg_log_std, log_sqrt_2pi);
        const float ratio = expf(lp - lp_old);
        const float clipped = ratio >= upper ? upper : (ratio <= lower ? lower : ratio);  // Matrix.Clip, Matrix.cs:377-405
        const float cra = clipped * adv, ra = ratio * adv;
        const float partA = (ra <= cra ? 1.0f : 0.0f) * adv;                              // Matrix.LessThan (<=)
        const float partB = (cra < ra ? 1.0f : 0.0f) * adv;                               // Matrix.LessThanNotEquals
        const float partC = (ratio >= lower && ratio <= upper) ? 1.0f : 0.0f;             // Matrix.InRange
        float dclip = (partA + partB * partC) * -1.0f;
        const float pold = expf(lp_old);
        if (pold == 0.0f) skip = true;                                                    // HadamardDivision throws (Matrix.cs:348-374)
        dclip = dclip / pold;
        const float dmean = expf(lp) * ((a - mu) / variance);
        const float gk = (dmean * dclip) / p.batch_size;
        g0[s * kGenMaxWidth + k] = gk;
        sum += gk;
      }
      if (skip) {  // the reference `continue`s: no feedback for this sample, no loss contribution
        for (int k = 0; k < act; k++) g0[s * kGenMaxWidth + k] = 0.0f;
        gvs = 0.0f;
      }
      gv[s] = gvs;
      // (losses and the skipped count: thread 0 adds them below in sample order)
      g1[s * kGenMaxWidth] = skip ? 1.0f : 0.0f;
      g1[s * kGenMaxWidth + 1] = skip ? 0.0f : sum / (float)act;  // Matrix.Average, PPOAgent.cs:332
    }
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < ts; s++) {
        sums[0] += gv[s];
        sums[1] += g1[s * kGenMaxWidth + 1];
        sums[2] += g1[s * kGenMaxWidth];
      }
    }
    __syncthreads();
    // actor: g0 holds dL/dmu
    gen_backward(p.actor, paramsA, cacheA, strideA, ts, g0, g1, partial);
    // critic: dL/dV into column 0 of g0
    for (int s = tid; s < ts; s += blockDim.x) g0[s * kGenMaxWidth] = gv[s];
    __syncthreads();
    gen_backward(p.critic, paramsC, cacheC, strideC, ts, g0, g1, partial + p.actor.n_params);
  }
  if (grad) {
    __syncthreads();
    const int total = p.actor.n_params + p.critic.n_params;
    if (tid < 3) partial[total + tid] = sums[tid];
  }
}

__global__ void reduce_partials_n_kernel(const float* __restrict__ partials, int nparts, float* __restrict__ grads, int n_floats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_floats) return;
  float sum = 0.0f;
  for (int c = 0; c < nparts; c++) sum += partials[(size_t)c * n_floats + i];
  grads[i] = sum;
}

// DenseLayer.Adam for one parameter of any topology (DenseLayer.cs:125-159): same operation order as adam_update in mlp.cu
__global__ void adam_generic_kernel(const GenAdamParams a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_params) return;
  int layer = 0;
  while (layer + 1 < a.n_dense && i >= a.layer_end[layer]) layer++;
  const float g = a.grads[i];
  const float m = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, a.beta1), g), __fmul_rn(a.beta1, a.m[i]));
  const float v = __fadd_rn(__fmul_rn(a.beta2, a.v[i]), __fmul_rn(__fsub_rn(1.0f, a.beta2), __fmul_rn(g, g)));
  a.m[i] = m;
  a.v[i] = v;
  const float mhat = __fdiv_rn(m, a.corr1[layer]);
  const float vhat = __fdiv_rn(v, a.corr2[layer]);
  const float denom = __fadd_rn(__fsqrt_rn(vhat), a.eps);
  a.params[i] = __fsub_rn(a.params[i], __fmul_rn(a.alpha, __fdiv_rn(mhat, denom)));
}

}  // namespace

size_t generic_smem_bytes(const GenNet& actor, const GenNet& critic, int ts) {
  const size_t strideA = (size_t)actor.cache_floats + actor.output, strideC = (size_t)critic.cache_floats + critic.output;
  return sizeof(float) * ((size_t)ts * (strideA + strideC + 2 * kGenMaxWidth + 1) + 4);
}

int generic_tile_samples(const GenNet& actor, const GenNet& critic) {
  for (int ts = 32; ts >= 1; ts >>= 1)
    if (generic_smem_bytes(actor, critic, ts) <= 200 * 1024) return ts;
  return 0;
}

int generic_grid_for(int n, int ts, int sm_count) {
  const int ntiles = (n + ts - 1) / ts;
  const int cap = sm_count * 2;
  return ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap;
}

cudaError_t launch_mlp_generic(const GenParams& p, int grid, cudaStream_t stream) {
  const size_t smem = generic_smem_bytes(p.actor, p.critic, p.ts);
  static size_t configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(ppo_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = 200 * 1024;
  }
  ppo_generic_kernel<<<grid, kGenThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials_n(const float* partials, int nparts, float* grads, int n_floats, cudaStream_t stream) {
  reduce_partials_n_kernel<<<(n_floats + 127) / 128, 128, 0, stream>>>(partials, nparts, grads, n_floats);
  return cudaGetLastError();
}

cudaError_t launch_adam_generic(const GenAdamParams& a, cudaStream_t stream) {
  adam_generic_kernel<<<(a.n_params + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace wb
