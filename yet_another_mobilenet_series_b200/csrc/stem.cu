// Stem convolution: 3x3, stride 2, pad 1, 3 input channels (RGB) -> Cout (<= 64, multiple of 8)
// on NHWC bf16, sm_100a.
//
// Replaces, behind yamb_stem_conv_fwd / yamb_stem_conv_wgrad (include/yamb200.h), the cuDNN
// implicit-GEMM forward / weight-gradient kernels (and the nhwcAddPadding copies cuDNN needs for a
// 3-channel NHWC tensor) that round 1 used for the first layer of the network
// (reference models/mobilenet_supernet.py:124-130: ConvBNReLU(3, input_channel, stride=2)).
// No input gradient: the images do not require one.
//
// forward : thread = 2 horizontally adjacent output pixels x all output channels (8 at a time);
//           the 3x5x3 input patch sits in registers, the weights are broadcast from shared memory.
// wgrad   : dW[co][ci][ky][kx] = sum_pixels dh[p][co] * x[2p+tap][ci]; thread = (pair of output
//           channels, 2 of the 27 (tap, ci) positions); tiles of 16 x 32 output pixels are staged in
//           shared memory; per-thread register accumulators over the CTA's tiles, one fp32
//           reduction per weight element per CTA at the end.
// 864 MACs per output pixel against 6 + 64 bytes of traffic: FMA/LDS bound, not HBM bound.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "host_util.h"
#include "prims.cuh"

namespace yamb {

constexpr int kStemTH = 16, kStemTW = 32;            // output tile
constexpr int kStemIH = 2 * kStemTH + 1, kStemIW = 2 * kStemTW + 1;   // input tile incl. halo

struct StemDev {
  int N, H, W, Ho, Wo, Cout;
  const __nv_bfloat16* x;     // [N][H][W][3]
  const float* w;             // [Cout][3][3][3]
  __nv_bfloat16* y;           // [N][Ho][Wo][Cout]        (forward)
  const __nv_bfloat16* dh;    // [N][Ho][Wo][Cout]        (wgrad)
  float* dw;                  // [Cout][3][3][3] +=       (wgrad)
  int tiles_h, tiles_w, num_tiles;
};

// stage the input tile of output tile (n, ty, tx) as bf16 [IH][IW][3] (zero outside the image)
__device__ __forceinline__ void stem_stage_input(const StemDev& p, int n, int ty, int tx,
                                                 __nv_bfloat16* s_in) {
  const int iy0 = ty * kStemTH * 2 - 1, ix0 = tx * kStemTW * 2 - 1;
  const __nv_bfloat16* img = p.x + (size_t)n * p.H * p.W * 3;
  for (int e = threadIdx.x; e < kStemIH * kStemIW * 3; e += 256) {
    const int r = e / (kStemIW * 3), rem = e % (kStemIW * 3);
    const int iy = iy0 + r, ix = ix0 + rem / 3;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
      v = img[((size_t)iy * p.W + ix) * 3 + rem % 3];
    s_in[e] = v;
  }
}

__global__ void __launch_bounds__(256) stem_fwd_kernel(const __grid_constant__ StemDev p) {
  __shared__ __align__(16) float s_w[27 * 64];                     // [tap*3+ci][co]
  __shared__ __align__(16) __nv_bfloat16 s_in[kStemIH * kStemIW * 3];
  for (int e = threadIdx.x; e < 27 * p.Cout; e += 256) {
    const int co = e / 27, j = e % 27;           // w[co][ci][ky][kx]: j = ci*9 + ky*3 + kx
    const int ci = j / 9, tap = j % 9;
    s_w[(tap * 3 + ci) * p.Cout + co] = p.w[e];
  }
  const int ry = threadIdx.x / 16, cp = threadIdx.x % 16;          // output row, column pair
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const int n = t / (p.tiles_h * p.tiles_w), r = t % (p.tiles_h * p.tiles_w);
    const int ty = r / p.tiles_w, tx = r % p.tiles_w;
    __syncthreads();
    stem_stage_input(p, n, ty, tx, s_in);
    __syncthreads();
    // 3 x 5 x 3 input patch of this thread's two output pixels
    float in[3][5][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          in[a][b][c] = __bfloat162float(s_in[((2 * ry + a) * kStemIW + 4 * cp + b) * 3 + c]);
    const int oy = ty * kStemTH + ry, ox = tx * kStemTW + 2 * cp;
    const bool ok0 = oy < p.Ho && ox < p.Wo, ok1 = oy < p.Ho && ox + 1 < p.Wo;
    __nv_bfloat16* yrow = p.y + (((size_t)n * p.Ho + oy) * p.Wo + ox) * p.Cout;
    for (int c0 = 0; c0 < p.Cout; c0 += 8) {
      float acc[2][8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            const float4 w0 = *reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * 3 + ci) * p.Cout + c0);
            const float4 w1 = *reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * 3 + ci) * p.Cout + c0 + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float a0 = in[ky][kx][ci], a1 = in[ky][kx + 2][ci];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              acc[0][e] = fmaf(a0, wv[e], acc[0][e]);
              acc[1][e] = fmaf(a1, wv[e], acc[1][e]);
            }
          }
      if (ok0)
        *reinterpret_cast<uint4*>(yrow + c0) =
            make_uint4(pack_bf16(acc[0][0], acc[0][1]), pack_bf16(acc[0][2], acc[0][3]),
                       pack_bf16(acc[0][4], acc[0][5]), pack_bf16(acc[0][6], acc[0][7]));
      if (ok1)
        *reinterpret_cast<uint4*>(yrow + p.Cout + c0) =
            make_uint4(pack_bf16(acc[1][0], acc[1][1]), pack_bf16(acc[1][2], acc[1][3]),
                       pack_bf16(acc[1][4], acc[1][5]), pack_bf16(acc[1][6], acc[1][7]));
    }
  }
}

// thread = (co pair q = tid % (Cout/2), position group jg = tid / (Cout/2)); positions j = jg + k*G
template <int kMaxJ>   // positions per thread = ceil(27 / G)
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const __grid_constant__ StemDev p) {
  extern __shared__ __align__(16) unsigned char stem_smem[];
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(stem_smem);          // [IH][IW][3]
  __nv_bfloat16* s_dh = s_in + ((kStemIH * kStemIW * 3 + 7) & ~7);            // [TH*TW][Cout]
  const int CP = p.Cout / 2;
  const int G = 256 / CP;                           // position groups (Cout=32: 16)
  const int q = threadIdx.x % CP, jg = threadIdx.x / CP;
  float acc[kMaxJ][2];
  int joff[kMaxJ];                                  // smem offset of (tap, ci) relative to a pixel
  bool jok[kMaxJ];
#pragma unroll
  for (int k = 0; k < kMaxJ; ++k) {
    acc[k][0] = acc[k][1] = 0.f;
    const int j = jg + k * G;                       // j = ci*9 + ky*3 + kx (the parameter's layout)
    jok[k] = jg < G && j < 27;
    const int jj = jok[k] ? j : 0;
    const int ci = jj / 9, ky = (jj % 9) / 3, kx = jj % 3;
    joff[k] = (ky * kStemIW + kx) * 3 + ci;
  }
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const int n = t / (p.tiles_h * p.tiles_w), r = t % (p.tiles_h * p.tiles_w);
    const int ty = r / p.tiles_w, tx = r % p.tiles_w;
    __syncthreads();
    stem_stage_input(p, n, ty, tx, s_in);
    // dh tile, zero outside the image: 16-byte vectors
    const int V = p.Cout / 8;
    for (int e = threadIdx.x; e < kStemTH * kStemTW * V; e += 256) {
      const int pix = e / V, v = e % V;
      const int oy = ty * kStemTH + pix / kStemTW, ox = tx * kStemTW + pix % kStemTW;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (oy < p.Ho && ox < p.Wo)
        val = __ldg(reinterpret_cast<const uint4*>(p.dh + (((size_t)n * p.Ho + oy) * p.Wo + ox) * p.Cout) + v);
      reinterpret_cast<uint4*>(s_dh)[e] = val;
    }
    __syncthreads();
    if (jg < G) {
#pragma unroll 2
      for (int pix = 0; pix < kStemTH * kStemTW; ++pix) {
        const uint32_t d = reinterpret_cast<const uint32_t*>(s_dh + (size_t)pix * p.Cout)[q];
        const float d0 = bf16lo(d), d1 = bf16hi(d);
        const int base = ((2 * (pix / kStemTW)) * kStemIW + 2 * (pix % kStemTW)) * 3;
#pragma unroll
        for (int k = 0; k < kMaxJ; ++k) {
          const float xv = __bfloat162float(s_in[base + joff[k]]);
          acc[k][0] = fmaf(d0, xv, acc[k][0]);
          acc[k][1] = fmaf(d1, xv, acc[k][1]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxJ; ++k)
    if (jok[k]) {
      const int j = jg + k * G;
      atomicAdd(p.dw + (size_t)(2 * q) * 27 + j, acc[k][0]);
      atomicAdd(p.dw + (size_t)(2 * q + 1) * 27 + j, acc[k][1]);
    }
}

static int stem_fill(const yamb_stem_conv* a, StemDev& p) {
  if (!a || a->N <= 0 || a->H <= 0 || a->W <= 0 || !a->x)
    return set_error(YAMB_EINVAL, "stem conv: bad arguments");
  if (a->Cout <= 0 || a->Cout > 64 || (a->Cout % 8))
    return set_error(YAMB_EINVAL, "stem conv: Cout=%d must be a multiple of 8, <= 64", a->Cout);
  if (max_ctas() <= 0) return set_error(YAMB_ENODEV, "no CUDA device");
  p.N = a->N; p.H = a->H; p.W = a->W; p.Cout = a->Cout;
  p.Ho = (a->H - 1) / 2 + 1; p.Wo = (a->W - 1) / 2 + 1;
  p.x = (const __nv_bfloat16*)a->x; p.w = a->w; p.y = (__nv_bfloat16*)a->y;
  p.dh = (const __nv_bfloat16*)a->dh; p.dw = a->dw;
  p.tiles_h = (p.Ho + kStemTH - 1) / kStemTH;
  p.tiles_w = (p.Wo + kStemTW - 1) / kStemTW;
  long long nt = (long long)a->N * p.tiles_h * p.tiles_w;
  if (nt > 0x7fffffffLL) return set_error(YAMB_EINVAL, "stem conv: too many tiles");
  p.num_tiles = (int)nt;
  return 0;
}

int stem_conv_fwd_launch(const yamb_stem_conv* a, cudaStream_t st) {
  StemDev p;
  int rc = stem_fill(a, p);
  if (rc) return rc;
  if (!a->w || !a->y || (reinterpret_cast<uintptr_t>(a->y) & 15))
    return set_error(YAMB_EINVAL, "stem conv fwd: w / y (16-byte aligned) required");
  const int cap = 4 * max_ctas();
  stem_fwd_kernel<<<p.num_tiles < cap ? p.num_tiles : cap, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "stem conv fwd: %s", cudaGetErrorString(e));
  return 0;
}

int stem_conv_wgrad_launch(const yamb_stem_conv* a, cudaStream_t st) {
  StemDev p;
  int rc = stem_fill(a, p);
  if (rc) return rc;
  if (!a->dh || !a->dw || (reinterpret_cast<uintptr_t>(a->dh) & 15))
    return set_error(YAMB_EINVAL, "stem conv wgrad: dh (16-byte aligned) / dw required");
  if (256 % (a->Cout / 2)) return set_error(YAMB_EINVAL, "stem conv wgrad: Cout/2 must divide 256");
  const size_t smem = (size_t)((kStemIH * kStemIW * 3 + 7) & ~7) * 2 +
                      (size_t)kStemTH * kStemTW * a->Cout * 2;
  const int G = 256 / (a->Cout / 2);
  const int kj = (27 + G - 1) / G;              // 1 (Cout <= 16), 2 (32), 4 (64)
  const int cap = 2 * max_ctas();
  const int grid = p.num_tiles < cap ? p.num_tiles : cap;
  cudaError_t e = cudaSuccess;
#define YAMB_STEM_WGRAD(KJ)                                                                        \
  do {                                                                                            \
    static size_t attr = 0; /* process-wide: only ever raise the limit */                         \
    if (smem > attr) {                                                                            \
      e = cudaFuncSetAttribute(stem_wgrad_kernel<KJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               (int)smem);                                                        \
      if (e != cudaSuccess) return set_error(YAMB_ECUDA, "stem wgrad attr: %s", cudaGetErrorString(e)); \
      attr = smem;                                                                                \
    }                                                                                             \
    stem_wgrad_kernel<KJ><<<grid, 256, smem, st>>>(p);                                            \
  } while (0)
  if (kj <= 1) YAMB_STEM_WGRAD(1);
  else if (kj == 2) YAMB_STEM_WGRAD(2);
  else YAMB_STEM_WGRAD(4);
#undef YAMB_STEM_WGRAD
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "stem conv wgrad: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace yamb
