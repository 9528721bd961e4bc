// Squeeze-and-Excitation fully connected layers on the pooled [N, C] vectors, sm_100a.
//
// Replaces, behind yamb_se_fc_fwd / yamb_se_fc_bwd (include/yamb200.h), the two 1x1 convolutions
// on [N, C, 1, 1] of SqueezeAndExcitation.forward and the sigmoid
// (reference models/mobilenet_base.py:110-113: se_reduce -> active_fn -> se_expand -> sigmoid) and
// their autograd backward, which round 1 left to ~12 ATen/cuBLAS launches per SE block:
//   forward : u = W_r s + b_r ;  v = act(u) ;  gate = sigmoid(W_e v + b_e)
//   backward: dt = dgate*gate*(1-gate) ;  du = (W_e^T dt) * act'(u) ;  dpool = (W_r^T du) / HW ;
//             dW_e += dt^T v, db_e += sum dt, dW_r += du^T s, db_r += sum du   (sums over samples)
// fp32 throughout (the reference keeps these tiny layers in fp32 too).  One CTA per sample for
// the per-sample chain; the parameter gradients are owned one-output-per-thread (no global
// atomics; the per-sample chain uses a handful of shared-memory float adds).  C <= 4096, R <= 1024.
#include <cuda_runtime.h>

#include "host_util.h"
#include "prims.cuh"

namespace yamb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) se_fc_fwd_kernel(const __grid_constant__ yamb_se_fc a) {
  extern __shared__ float sf[];
  float* s_s = sf;            // [C] pooled row
  float* s_v = sf + a.C;      // [R]
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < a.C; c += 256) s_s[c] = a.pooled[(size_t)n * a.C + c];
  __syncthreads();
  for (int j = warp; j < a.R; j += 8) {          // one warp per squeeze unit: coalesced W_r row
    const float* w = a.w_r + (size_t)j * a.C;
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) acc = fmaf(w[c], s_s[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float u = acc + a.b_r[j];
      const float v = act_fwd(u, a.act);
      a.u[(size_t)n * a.R + j] = u;
      a.v[(size_t)n * a.R + j] = v;
      s_v[j] = v;
    }
  }
  __syncthreads();
  for (int c = tid; c < a.C; c += 256) {
    const float* w = a.w_e + (size_t)c * a.R;
    float acc0 = a.b_e[c], acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int j = 0;
    for (; j + 3 < a.R; j += 4) {          // 4 independent chains: the loads pipeline
      acc0 = fmaf(w[j], s_v[j], acc0);
      acc1 = fmaf(w[j + 1], s_v[j + 1], acc1);
      acc2 = fmaf(w[j + 2], s_v[j + 2], acc2);
      acc3 = fmaf(w[j + 3], s_v[j + 3], acc3);
    }
    for (; j < a.R; ++j) acc0 = fmaf(w[j], s_v[j], acc0);
    const float acc = (acc0 + acc1) + (acc2 + acc3);
    a.gate[(size_t)n * a.C + c] = 1.f / (1.f + __expf(-acc));
  }
}

// per-sample chain of the backward, kSeS samples per CTA: every weight element fetched from L2 is
// used for kSeS samples, 16 weight loads in flight per thread (one sample per CTA re-read both weight matrices per sample — 256 MB of L2
// traffic and ~1000 dependent load batches per SE block of AtomNAS-C+, 225 us)
constexpr int kSeS = 2;
__global__ void __launch_bounds__(256) se_fc_bwd_sample_kernel(const __grid_constant__ yamb_se_fc_grad a) {
  extern __shared__ float sf[];
  float* s_dt = sf;                    // [kSeS][C]
  float* s_du = sf + kSeS * a.C;       // [kSeS][R]
  const int n0 = blockIdx.x * kSeS, tid = threadIdx.x;
  const int ns = min(kSeS, a.N - n0);
  for (int e = tid; e < kSeS * a.C; e += 256) {
    const int s = e / a.C, c = e - s * a.C;
    float dt = 0.f;
    if (s < ns) {
      const float g = a.gate[(size_t)(n0 + s) * a.C + c];
      dt = a.dgate[(size_t)(n0 + s) * a.C + c] * g * (1.f - g);
      a.dt[(size_t)(n0 + s) * a.C + c] = dt;
    }
    s_dt[e] = dt;
  }
  for (int e = tid; e < kSeS * a.R; e += 256) s_du[e] = 0.f;
  __syncthreads();
  // dv[s][j] = sum_c dt[s][c] * W_e[c][j]: consecutive threads own consecutive columns j (coalesced
  // rows of W_e); the C-long sum is split over the P = 256 / R' thread groups
  {
    int Rp = 32;
    while (Rp < a.R && Rp < 256) Rp <<= 1;          // columns per pass, power of two <= 256
    const int P = 256 / Rp, part = tid / Rp;
    for (int j0 = 0; j0 < a.R; j0 += Rp) {
      const int j = j0 + (tid % Rp);
      if (j < a.R) {
        float acc[kSeS];
#pragma unroll
        for (int s = 0; s < kSeS; ++s) acc[s] = 0.f;
#pragma unroll 16
        for (int c = part; c < a.C; c += P) {      // 16 independent weight loads in flight
          const float w = a.w_e[(size_t)c * a.R + j];
#pragma unroll
          for (int s = 0; s < kSeS; ++s) acc[s] = fmaf(s_dt[s * a.C + c], w, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < kSeS; ++s) atomicAdd(&s_du[s * a.R + j], acc[s]);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < kSeS * a.R; e += 256) {
    const int s = e / a.R, j = e - s * a.R;
    float du = 0.f;
    if (s < ns) {
      du = s_du[e] * act_bwd(a.u[(size_t)(n0 + s) * a.R + j], a.act);
      a.du[(size_t)(n0 + s) * a.R + j] = du;
    }
    s_du[e] = du;
  }
  __syncthreads();
  for (int c = tid; c < a.C; c += 256) {
    float acc[kSeS];
#pragma unroll
    for (int s = 0; s < kSeS; ++s) acc[s] = 0.f;
#pragma unroll 8
    for (int j = 0; j < a.R; ++j) {
      const float w = a.w_r[(size_t)j * a.C + c];
#pragma unroll
      for (int s = 0; s < kSeS; ++s) acc[s] = fmaf(s_du[s * a.R + j], w, acc[s]);
    }
    for (int s = 0; s < ns; ++s) a.dpool[(size_t)(n0 + s) * a.C + c] = acc[s] * a.inv_hw;
  }
}

// parameter gradients: thread (c, j) owns dW_e[c][j] and dW_r[j][c]; the bias sums ride along
__global__ void __launch_bounds__(256) se_fc_bwd_param_kernel(const __grid_constant__ yamb_se_fc_grad a) {
  const long long total = (long long)a.C * a.R;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total;
       i += (long long)gridDim.x * 256) {
    const int c = (int)(i / a.R), j = (int)(i % a.R);
    float we = 0.f, wr = 0.f, be = 0.f, br = 0.f;
#pragma unroll 4
    for (int n = 0; n < a.N; ++n) {
      const float dt = a.dt[(size_t)n * a.C + c], du = a.du[(size_t)n * a.R + j];
      we = fmaf(dt, a.v[(size_t)n * a.R + j], we);
      wr = fmaf(du, a.pooled[(size_t)n * a.C + c], wr);
      be += dt;
      br += du;
    }
    a.g_we[(size_t)c * a.R + j] += we;
    a.g_wr[(size_t)j * a.C + c] += wr;
    if (j == 0) a.g_be[c] += be;
    if (c == 0) a.g_br[j] += br;
  }
}

static int se_fc_check(int N, int C, int R) {
  if (N <= 0 || C <= 0 || R <= 0 || C > 4096 || R > 1024)
    return set_error(YAMB_EINVAL, "se_fc: N=%d C=%d R=%d", N, C, R);
  if (max_ctas() <= 0) return set_error(YAMB_ENODEV, "no CUDA device");
  return 0;
}

int se_fc_fwd_launch(const yamb_se_fc* a, cudaStream_t st) {
  if (!a) return set_error(YAMB_EINVAL, "null args");
  int rc = se_fc_check(a->N, a->C, a->R);
  if (rc) return rc;
  if (!a->pooled || !a->w_r || !a->b_r || !a->w_e || !a->b_e || !a->u || !a->v || !a->gate)
    return set_error(YAMB_EINVAL, "se_fc: null pointer");
  se_fc_fwd_kernel<<<a->N, 256, (a->C + a->R) * sizeof(float), st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "se_fc fwd: %s", cudaGetErrorString(e));
  return 0;
}

int se_fc_bwd_launch(const yamb_se_fc_grad* a, cudaStream_t st) {
  if (!a) return set_error(YAMB_EINVAL, "null args");
  int rc = se_fc_check(a->N, a->C, a->R);
  if (rc) return rc;
  if (!a->dgate || !a->gate || !a->u || !a->v || !a->pooled || !a->w_r || !a->w_e || !a->dpool ||
      !a->dt || !a->du || !a->g_wr || !a->g_br || !a->g_we || !a->g_be)
    return set_error(YAMB_EINVAL, "se_fc bwd: null pointer");
  const size_t smem = (size_t)kSeS * (a->C + a->R) * sizeof(float);
  static size_t attr = 0;   // process-wide: only ever raise the limit
  cudaError_t e;
  if (smem > attr && smem > 48 * 1024) {
    e = cudaFuncSetAttribute(se_fc_bwd_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(YAMB_ECUDA, "se_fc bwd attr: %s", cudaGetErrorString(e));
    attr = smem;
  }
  se_fc_bwd_sample_kernel<<<(a->N + kSeS - 1) / kSeS, 256, smem, st>>>(*a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "se_fc bwd: %s", cudaGetErrorString(e));
  const long long total = (long long)a->C * a->R;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)max_ctas() * 8;
  se_fc_bwd_param_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(*a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "se_fc bwd params: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace yamb
