// mlp_tc.cu -- tensor-core (tcgen05 / TMEM) building blocks for the PPO MLP on sm_100a.
//
// Operand tiles live in shared memory in ONE physical layout, [feature/4][sample][4 floats] (16-byte units, no swizzle):
//   * read as a K-major operand    (rows = samples,  K = features): SBO = 128 B (8 rows x 16 B), LBO = rows*16 B
//   * read as an MN-major operand  (rows = features, K = samples) : SBO = samples*16 B,         LBO = 128 B
// so the same activation / gradient tile feeds the forward and dX GEMMs (K-major) and the weight-gradient GEMMs
// (MN-major, reduction over samples) without a transpose.  fp32 accuracy comes from a 3xTF32 split (hi*hi + lo*hi + hi*lo).
#include "mlp.cuh"
#include "tc.cuh"

namespace wb {

// Dual-use operand tile ("B32" layout = UMMA SWIZZLE_128B_BASE32B): element (sample s, feature f) of a tile with S samples
// lives at float offset ((f/32)*S + s)*32 + ((((f/8)&3) ^ (s&3)) * 8) + (f&7): rows of 32 features (128 B), 32-byte units
// XOR-swizzled by (s & 3).  Read K-major (rows = samples, K = features) or MN-major (rows = features, K = samples).
__device__ __forceinline__ int tile_off_b32(int s, int f, int S) {
  return ((f >> 5) * S + s) * 32 + ((((f >> 3) & 3) ^ (s & 3)) << 3) + (f & 7);
}
constexpr uint64_t kLayoutB32 = (uint64_t)1 << 61;  // layout_type = 1 (SWIZZLE_128B_BASE32B)

// ---------------------------------------------------------------- debug: one CTA, D[M x N] = A[M x K] * B[N x K]^T
// a_mn / b_mn: 0 = the operand is staged for K-major reading, 1 = for MN-major reading (see header comment).
// passes: 1 = plain TF32 (hi*hi), 3 = 3xTF32.
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                           int M, int N, int K, int a_mn, int b_mn, int passes) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // tile sizes in floats (B32 tiles are padded to whole 32-feature blocks; every tile starts 1024-B aligned)
  const int a_sz = (a_mn == 2) ? ((K + 31) / 32) * M * 32 : (a_mn == 3) ? ((M + 31) / 32) * K * 32 : M * K;
  const int b_sz = (b_mn == 2) ? ((K + 31) / 32) * N * 32 : (b_mn == 3) ? ((N + 31) / 32) * K * 32 : N * K;
  const int a_al = (a_sz + 255) & ~255, b_al = (b_sz + 255) & ~255;
  float* a_hi = reinterpret_cast<float*>(smem_raw);
  float* a_lo = a_hi + a_al;
  float* b_hi = a_lo + a_al;
  float* b_lo = b_hi + b_al;
  for (int idx = threadIdx.x; idx < 2 * a_al + 2 * b_al; idx += blockDim.x) a_hi[idx] = 0.0f;
  __syncthreads();
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tc::tmem_alloc(&tmem_base_slot, 128);
  if (tid == 0) {
    tc::mbar_init(&mbar, 1);
    tc::mbar_fence_init();
  }
  // stage operands: element (r, k) of a row-major [R x K] matrix
  for (int idx = tid; idx < M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    float hi, lo;
    tc::split_tf32(A[idx], hi, lo);
    const int off = (a_mn >= 2) ? (a_mn == 2 ? tile_off_b32(r, k, M) : tile_off_b32(k, r, K))
                                : (a_mn ? ((r >> 2) * K + k) * 4 + (r & 3) : ((k >> 2) * M + r) * 4 + (k & 3));
    a_hi[off] = hi;
    a_lo[off] = lo;
  }
  for (int idx = tid; idx < N * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    float hi, lo;
    tc::split_tf32(B[idx], hi, lo);
    const int off = (b_mn >= 2) ? (b_mn == 2 ? tile_off_b32(r, k, N) : tile_off_b32(k, r, K))
                                : (b_mn ? ((r >> 2) * K + k) * 4 + (r & 3) : ((k >> 2) * N + r) * 4 + (k & 3));
    b_hi[off] = hi;
    b_lo[off] = lo;
  }
  const int probe = passes >= 100 ? passes - 100 : -1;  // probe mode: fill the B tile with its own float offsets
  if (probe >= 0) {
    for (int idx = tid; idx < 12288; idx += blockDim.x) b_hi[idx] = (float)idx;  // 48 KB window after the B tile base
    passes = 1;
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(M, N, a_mn & 1, b_mn & 1);
    uint32_t a_lbo = a_mn ? 128u : (uint32_t)M * 16u, a_sbo = a_mn ? (uint32_t)K * 16u : 128u;
    uint32_t b_lbo = b_mn ? 128u : (uint32_t)N * 16u, b_sbo = b_mn ? (uint32_t)K * 16u : 128u;
    uint32_t a_step = a_mn ? 128u : 2u * (uint32_t)M * 16u;  // bytes per k-step of 8
    uint32_t b_step = b_mn ? 128u : 2u * (uint32_t)N * 16u;
    uint64_t a_layout_bits = 0, b_layout_bits = 0;
    if (a_mn >= 2) {  // B32 layout: K-major -> 32 B per k-step inside a 128-B row; MN-major -> 8 samples = 2 groups of 4 rows
      a_layout_bits = kLayoutB32;
      a_sbo = 512u;
      a_lbo = (a_mn == 2) ? 0u : (uint32_t)K * 128u;
      a_step = (a_mn == 2) ? 32u : 1024u;
    }
    if (probe == 1) {
      const uint32_t t = b_lbo;
      b_lbo = b_sbo;
      b_sbo = t;
    }
    if (probe >= 2) {
      b_layout_bits = (uint64_t)(probe - 1) << 61;
      b_lbo = 4096u;
      b_sbo = 1024u;
    }
    if (b_mn >= 2 && probe < 0) {
      b_layout_bits = kLayoutB32;
      b_sbo = 512u;
      b_lbo = (b_mn == 2) ? 0u : (uint32_t)K * 128u;
      b_step = (b_mn == 2) ? 32u : 1024u;
    }
    bool acc = false;
    for (int pass = 0; pass < passes; pass++) {
      const float* at = (pass == 1) ? a_lo : a_hi;
      const float* bt = (pass == 2) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 8; ks++) {
        // K-major B32 tiles: four k-steps per 32-feature block, blocks are rows*128 B apart
        const uint32_t a_off = (a_mn == 2) ? (uint32_t)(ks >> 2) * M * 128u + (ks & 3) * 32u : ks * a_step;
        const uint32_t b_off = (b_mn == 2) ? (uint32_t)(ks >> 2) * N * 128u + (ks & 3) * 32u : ks * b_step;
        const uint64_t da = tc::make_smem_desc(tc::smem_u32(at) + a_off, a_lbo, a_sbo) | a_layout_bits;
        const uint64_t db = tc::make_smem_desc(tc::smem_u32(bt) + b_off, b_lbo, b_sbo) | b_layout_bits;
        tc::mma_tf32(tmem_base, da, db, idesc, acc);
        acc = true;
      }
    }
    tc::mma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::fence_after_thread_sync();
  // accumulator rows: M = 128 -> row = TMEM lane; M = 64 -> row r sits in lane (r % 16) + 32 * (r / 16)
  const int row = (M == 128) ? tid : (lane < 16 ? warp * 16 + lane : -1);
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tc::tmem_ld_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    if (row >= 0 && row < M)
      for (int j = 0; j < 8; j++) D[row * N + c0 + j] = v[j];
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

cudaError_t launch_tc_gemm_test(const float* A, const float* B, float* D, int M, int N, int K, int a_mn, int b_mn, int passes,
                                cudaStream_t stream) {
  size_t smem = 160 * 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_gemm_test_kernel<<<1, 128, smem, stream>>>(A, B, D, M, N, K, a_mn, b_mn, passes);
  return cudaGetLastError();
}

}  // namespace wb
