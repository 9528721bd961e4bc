// Depthwise k x k convolution (k in {3,5,7}, stride 1/2) on NHWC bf16 activations, sm_100a.
//
// Replaces, behind yamb_depthwise_fwd / yamb_depthwise_bwd (include/yamb200.h), the
// nn.Conv2d(groups=C) + BatchNorm2d + activation library calls of the reference block
// (models/mobilenet_base.py:405-411 unfused, :275-281 fused) and their autograd backward.
//
// forward : y = dwconv(act(in_scale*x + in_shift))   — the producer's BatchNorm + activation is
//           applied on load in fp32 (the normalised tensor never exists in HBM); per-channel
//           sum / sum^2 of the bf16 output feed the next BatchNorm (last-CTA finalize).
// backward: dh = ca*dz + cb*h + cc (BatchNorm backward of the depthwise output, applied on load),
//           da = dwconv^T(dh, w), dwgt += sum dh * a  (fused wgrad), dx = da * act'(z) with the
//           statistics of the preceding BatchNorm's backward.
//
// HBM-bound integer-free stencil: channels are the contiguous dimension; a thread owns VEC
// consecutive channels (16/8/4-byte vector accesses), consecutive threads own consecutive channel
// groups and then consecutive pixels, so every warp access is a contiguous run of >= 128 bytes.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bn_finalize.cuh"
#include "host_util.h"
#include "prims.cuh"

namespace yamb {

struct DwFwdDev {
  int N, H, W, Ho, Wo, C, ldc;
  const __nv_bfloat16* x;
  const float* in_scale;
  const float* in_shift;
  int in_act;
  const float* w;
  __nv_bfloat16* y;
  int has_bn;
  yamb_bn_fwd bn;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  } else if constexpr (VEC == 4) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
  } else {
    uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    v[0] = bf16lo(u); v[1] = bf16hi(u);
  }
}
// round to bf16, store, and return the rounded values in v
template <int VEC>
__device__ __forceinline__ void store_vec_round(__nv_bfloat16* p, float (&v)[VEC]) {
  uint32_t u[VEC / 2];
#pragma unroll
  for (int i = 0; i < VEC / 2; ++i) {
    u[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    v[2 * i] = bf16lo(u[i]);
    v[2 * i + 1] = bf16hi(u[i]);
  }
  if constexpr (VEC == 8) *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  else if constexpr (VEC == 4) *reinterpret_cast<uint2*>(p) = make_uint2(u[0], u[1]);
  else *reinterpret_cast<uint32_t*>(p) = u[0];
}

template <int K, int VEC, int S>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const __grid_constant__ DwFwdDev p) {
  constexpr int TH = 4;                  // output rows per thread (vertical strip)
  constexpr int P = (K - 1) / 2;
  constexpr int IR = (TH - 1) * S + K;   // input rows touched by one strip
  extern __shared__ float s_part[];      // [2][C] per-CTA statistics
  const int CG = p.C / VEC;
  const int PX = 256 / CG;
  const int px = threadIdx.x / CG;
  const int cg = threadIdx.x % CG;
  const int c0 = cg * VEC;
  const bool active = px < PX;
  if (p.has_bn)
    for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) s_part[i] = 0.f;
  __syncthreads();

  float ssum[VEC], ssq[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) ssum[v] = ssq[v] = 0.f;

  if (active) {
    float w[K * K][VEC];
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
      for (int t = 0; t < K * K; ++t) w[t][v] = __ldg(p.w + (size_t)(c0 + v) * K * K + t);
      sc[v] = p.in_scale ? __ldg(p.in_scale + c0 + v) : 1.f;
      sh[v] = p.in_scale ? __ldg(p.in_shift + c0 + v) : 0.f;
    }
    const int act = p.in_scale ? p.in_act : ACT_NONE;
    const int nstrips = (p.Ho + TH - 1) / TH;
    const long long items = (long long)p.N * nstrips * p.Wo;
    for (long long item = (long long)blockIdx.x * PX + px; item < items;
         item += (long long)gridDim.x * PX) {
      const int xo = (int)(item % p.Wo);
      const long long t = item / p.Wo;
      const int strip = (int)(t % nstrips);
      const int n = (int)(t / nstrips);
      const int yo0 = strip * TH;
      float acc[TH][VEC];
#pragma unroll
      for (int j = 0; j < TH; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[j][v] = 0.f;
#pragma unroll
      for (int r = 0; r < IR; ++r) {
        const int yi = yo0 * S - P + r;
        if (yi < 0 || yi >= p.H) continue;
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const int xi = xo * S - P + dx;
          if (xi < 0 || xi >= p.W) continue;
          float a[VEC];
          load_vec<VEC>(p.x + ((size_t)((size_t)n * p.H + yi) * p.W + xi) * p.ldc + c0, a);
#pragma unroll
          for (int v = 0; v < VEC; ++v) a[v] = act_fwd(fmaf(sc[v], a[v], sh[v]), act);
#pragma unroll
          for (int j = 0; j < TH; ++j) {
            const int ky = r - j * S;  // compile-time after unrolling
            if (ky >= 0 && ky < K) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[j][v] = fmaf(w[ky * K + dx][v], a[v], acc[j][v]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TH; ++j) {
        const int yo = yo0 + j;
        if (yo < p.Ho) {
          store_vec_round<VEC>(p.y + ((size_t)((size_t)n * p.Ho + yo) * p.Wo + xo) * p.ldc + c0,
                               acc[j]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            ssum[v] += acc[j][v];
            ssq[v] = fmaf(acc[j][v], acc[j][v], ssq[v]);
          }
        }
      }
    }
  }
  if (p.has_bn) {
    if (active) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        atomicAdd(&s_part[c0 + v], ssum[v]);
        atomicAdd(&s_part[p.C + c0 + v], ssq[v]);
      }
    }
    __syncthreads();
    if (publish_partials(s_part, p.C, p.bn.partials, p.bn.counter)) {
      bn_fwd_finalize(p.bn, p.C, gridDim.x);
      __syncthreads();
      if (threadIdx.x == 0) *p.bn.counter = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct DwBwdDev {
  int N, H, W, Ho, Wo, C, ldc;
  const __nv_bfloat16* dz;  // [N,Ho,Wo,ldc]
  const __nv_bfloat16* h;   // [N,Ho,Wo,ldc]
  const float *ca, *cb, *cc;
  const float* w;           // [C][K][K]
  float* dw;                // [C][K][K] +=
  const __nv_bfloat16* x;   // [N,H,W,ldc] pre-BN input of the depthwise stage
  const float *in_scale, *in_shift;
  int in_act;
  __nv_bfloat16* dx;        // [N,H,W,ldc]
  const __nv_bfloat16* residual;
  int has_bn;
  yamb_bn_bwd bn;
};

template <int K, int VEC, int S>
__global__ void __launch_bounds__(256) dw_bwd_kernel(const __grid_constant__ DwBwdDev p) {
  constexpr int P = (K - 1) / 2;
  extern __shared__ float smem_f[];
  // layout: s_w[K*K][C] | s_gw[K*K][C] | s_part[2][C]
  float* s_w = smem_f;
  float* s_gw = s_w + K * K * p.C;
  float* s_part = s_gw + K * K * p.C;
  const int CG = p.C / VEC;
  const int PX = 256 / CG;
  const int px = threadIdx.x / CG;
  const int cg = threadIdx.x % CG;
  const int c0 = cg * VEC;
  const bool active = px < PX;
  for (int i = threadIdx.x; i < K * K * p.C; i += blockDim.x) {
    const int t = i / p.C, c = i % p.C;
    s_w[i] = __ldg(p.w + (size_t)c * K * K + t);
    s_gw[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) s_part[i] = 0.f;
  __syncthreads();

  float ssum[VEC], ssq[VEC];
  float gw[K * K][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ssum[v] = ssq[v] = 0.f;
#pragma unroll
    for (int t = 0; t < K * K; ++t) gw[t][v] = 0.f;
  }
  if (active) {
    float sc[VEC], sh[VEC], ca[VEC], cb[VEC], cc[VEC], mu[VEC], rs[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      sc[v] = p.in_scale ? __ldg(p.in_scale + c0 + v) : 1.f;
      sh[v] = p.in_scale ? __ldg(p.in_shift + c0 + v) : 0.f;
      ca[v] = __ldg(p.ca + c0 + v);
      cb[v] = __ldg(p.cb + c0 + v);
      cc[v] = __ldg(p.cc + c0 + v);
      mu[v] = p.has_bn ? __ldg(p.bn.mean + c0 + v) : 0.f;
      rs[v] = p.has_bn ? __ldg(p.bn.invstd + c0 + v) : 0.f;
    }
    const int act = p.in_scale ? p.in_act : ACT_NONE;
    const long long items = (long long)p.N * p.H * p.W;
    for (long long item = (long long)blockIdx.x * PX + px; item < items;
         item += (long long)gridDim.x * PX) {
      const int x = (int)(item % p.W);
      const long long t = item / p.W;
      const int y = (int)(t % p.H);
      const int n = (int)(t / p.H);
      const size_t in_off = (size_t)item * p.ldc + c0;
      float xv[VEC], a1[VEC], dact[VEC], da[VEC];
      load_vec<VEC>(p.x + in_off, xv);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float z = fmaf(sc[v], xv[v], sh[v]);
        a1[v] = act_fwd(z, act);
        dact[v] = act_bwd(z, act);
        da[v] = 0.f;
      }
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int yy = y + P - ky;
        if (yy < 0 || (S == 2 && (yy & 1))) continue;
        const int yo = yy / S;
        if (yo >= p.Ho) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int xx = x + P - kx;
          if (xx < 0 || (S == 2 && (xx & 1))) continue;
          const int xo = xx / S;
          if (xo >= p.Wo) continue;
          const size_t o = ((size_t)((size_t)n * p.Ho + yo) * p.Wo + xo) * p.ldc + c0;
          float dzv[VEC], hv[VEC];
          load_vec<VEC>(p.dz + o, dzv);
          load_vec<VEC>(p.h + o, hv);
          const float* wt = s_w + (ky * K + kx) * p.C + c0;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float dh = fmaf(ca[v], dzv[v], fmaf(cb[v], hv[v], cc[v]));
            da[v] = fmaf(dh, wt[v], da[v]);
            gw[ky * K + kx][v] = fmaf(dh, a1[v], gw[ky * K + kx][v]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) da[v] *= dact[v];
      if (p.residual) {
        float rv[VEC];
        load_vec<VEC>(p.residual + in_off, rv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) da[v] += rv[v];
      }
      store_vec_round<VEC>(p.dx + in_off, da);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        ssum[v] += da[v];
        ssq[v] = fmaf(da[v], (xv[v] - mu[v]) * rs[v], ssq[v]);
      }
    }
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
      for (int v = 0; v < VEC; ++v) atomicAdd(&s_gw[t * p.C + c0 + v], gw[t][v]);
    if (p.has_bn) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        atomicAdd(&s_part[c0 + v], ssum[v]);
        atomicAdd(&s_part[p.C + c0 + v], ssq[v]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * p.C; i += blockDim.x) {
    const int t = i / p.C, c = i % p.C;
    atomicAdd(p.dw + (size_t)c * K * K + t, s_gw[i]);
  }
  if (p.has_bn) {
    if (publish_partials(s_part, p.C, p.bn.partials, p.bn.counter)) {
      bn_bwd_finalize(p.bn, p.C, gridDim.x);
      __syncthreads();
      if (threadIdx.x == 0) *p.bn.counter = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int pick_vec(int C, int k, bool bwd) {
  int vec = k == 3 ? (bwd ? 4 : 8) : (k == 5 ? 4 : 2);
  while (vec > 2 && (C % vec)) vec >>= 1;
  while (vec < 8 && C / vec > 256) vec <<= 1;  // at most 256 channel groups per CTA
  return vec;
}

template <int K, int VEC>
static cudaError_t launch_fwd(const DwFwdDev& p, int stride, int grid, size_t smem,
                              cudaStream_t st) {
  if (stride == 1) dw_fwd_kernel<K, VEC, 1><<<grid, 256, smem, st>>>(p);
  else dw_fwd_kernel<K, VEC, 2><<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}
template <int K, int VEC>
static cudaError_t launch_bwd(const DwBwdDev& p, int stride, int grid, size_t smem,
                              cudaStream_t st) {
  cudaError_t e;
  if (stride == 1) {
    e = cudaFuncSetAttribute(dw_bwd_kernel<K, VEC, 1>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dw_bwd_kernel<K, VEC, 1><<<grid, 256, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(dw_bwd_kernel<K, VEC, 2>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dw_bwd_kernel<K, VEC, 2><<<grid, 256, smem, st>>>(p);
  }
  return cudaGetLastError();
}

#define YAMB_DISPATCH_KV(FN, ...)                                                        \
  do {                                                                                   \
    if (k == 3 && vec == 8) e = FN<3, 8>(__VA_ARGS__);                                   \
    else if (k == 3 && vec == 4) e = FN<3, 4>(__VA_ARGS__);                              \
    else if (k == 3 && vec == 2) e = FN<3, 2>(__VA_ARGS__);                              \
    else if (k == 5 && vec == 8) e = FN<5, 8>(__VA_ARGS__);                              \
    else if (k == 5 && vec == 4) e = FN<5, 4>(__VA_ARGS__);                              \
    else if (k == 5 && vec == 2) e = FN<5, 2>(__VA_ARGS__);                              \
    else if (k == 7 && vec == 4) e = FN<7, 4>(__VA_ARGS__);                              \
    else if (k == 7 && vec == 2) e = FN<7, 2>(__VA_ARGS__);                              \
    else return set_error(YAMB_EINVAL, "depthwise: unsupported k=%d vec=%d", k, vec);    \
  } while (0)

static int check_common(int N, int H, int W, int C, int ldc, int k, int stride) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0) return set_error(YAMB_EINVAL, "depthwise: bad shape");
  if (C % 2 || ldc % 8 || C > ldc) return set_error(YAMB_EINVAL, "depthwise: C=%d ldc=%d", C, ldc);
  if (k != 3 && k != 5 && k != 7) return set_error(YAMB_EINVAL, "depthwise: k=%d", k);
  if (stride != 1 && stride != 2) return set_error(YAMB_EINVAL, "depthwise: stride=%d", stride);
  if (C > 2048) return set_error(YAMB_EINVAL, "depthwise: C > 2048 per slice");
  return 0;
}

int dw_fwd_launch(const yamb_dw_fwd* a, cudaStream_t st) {
  if (!a) return set_error(YAMB_EINVAL, "null args");
  int rc = check_common(a->N, a->H, a->W, a->C, a->ldc, a->k, a->stride);
  if (rc) return rc;
  if (max_ctas() <= 0) return set_error(YAMB_ENODEV, "no CUDA device");
  const int k = a->k, pad = (k - 1) / 2;
  DwFwdDev p;
  p.N = a->N; p.H = a->H; p.W = a->W; p.C = a->C; p.ldc = a->ldc;
  p.Ho = (a->H + 2 * pad - k) / a->stride + 1;
  p.Wo = (a->W + 2 * pad - k) / a->stride + 1;
  p.x = (const __nv_bfloat16*)a->x; p.y = (__nv_bfloat16*)a->y;
  p.in_scale = a->in_scale; p.in_shift = a->in_shift; p.in_act = a->in_act;
  p.w = a->w;
  p.has_bn = a->bn ? 1 : 0;
  if (a->bn) p.bn = *a->bn;
  if ((((uintptr_t)a->x) | ((uintptr_t)a->y)) & 15)
    return set_error(YAMB_EINVAL, "depthwise: activations must be 16-byte aligned");
  const int vec = pick_vec(a->C, k, false);
  const int CG = a->C / vec, PX = 256 / CG;
  const long long items = (long long)a->N * ((p.Ho + 3) / 4) * p.Wo;
  long long want = (items + PX - 1) / PX;
  int grid = (int)(want < (long long)2 * max_ctas() ? want : 2 * max_ctas());
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)2 * a->C * sizeof(float);
  cudaError_t e;
  YAMB_DISPATCH_KV(launch_fwd, p, a->stride, grid, smem, st);
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "dw fwd launch: %s", cudaGetErrorString(e));
  return 0;
}

int dw_bwd_launch(const yamb_dw_bwd* a, cudaStream_t st) {
  if (!a) return set_error(YAMB_EINVAL, "null args");
  int rc = check_common(a->N, a->H, a->W, a->C, a->ldc, a->k, a->stride);
  if (rc) return rc;
  if (max_ctas() <= 0) return set_error(YAMB_ENODEV, "no CUDA device");
  if (!a->ca || !a->cb || !a->cc || !a->dz || !a->h || !a->x || !a->dx || !a->dw || !a->w)
    return set_error(YAMB_EINVAL, "depthwise bwd: null pointer");
  const int k = a->k, pad = (k - 1) / 2;
  DwBwdDev p;
  p.N = a->N; p.H = a->H; p.W = a->W; p.C = a->C; p.ldc = a->ldc;
  p.Ho = (a->H + 2 * pad - k) / a->stride + 1;
  p.Wo = (a->W + 2 * pad - k) / a->stride + 1;
  p.dz = (const __nv_bfloat16*)a->dz; p.h = (const __nv_bfloat16*)a->h;
  p.ca = a->ca; p.cb = a->cb; p.cc = a->cc;
  p.w = a->w; p.dw = a->dw;
  p.x = (const __nv_bfloat16*)a->x;
  p.in_scale = a->in_scale; p.in_shift = a->in_shift; p.in_act = a->in_act;
  p.dx = (__nv_bfloat16*)a->dx;
  p.residual = (const __nv_bfloat16*)a->residual;
  p.has_bn = a->bn ? 1 : 0;
  if (a->bn) p.bn = *a->bn;
  const int vec = pick_vec(a->C, k, true);
  const int CG = a->C / vec, PX = 256 / CG;
  const long long items = (long long)a->N * a->H * a->W;
  long long want = (items + PX - 1) / PX;
  int grid = (int)(want < (long long)2 * max_ctas() ? want : 2 * max_ctas());
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)(2 * k * k + 2) * a->C * sizeof(float);
  if (smem > 200 * 1024) return set_error(YAMB_EINVAL, "depthwise bwd: slice too wide for smem");
  cudaError_t e;
  YAMB_DISPATCH_KV(launch_bwd, p, a->stride, grid, smem, st);
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "dw bwd launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace yamb
