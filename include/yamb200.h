/* yamb200 — C ABI of the B200-native inverted-residual training path.
 *
 * This is the drop-in boundary of the hot path of meijieru/yet_another_mobilenet_series
 * (SURVEY.md §8b).  The reference's boundary is a Python nn.Module registry
 * (models/mobilenet_base.py:484-489 `get_block`) plus two YAML plugin hooks
 * (common.py:129-130 `FLAGS.model`, utils/optim.py:277-279 `FLAGS.optimizer`); the Python package
 * `yet_another_mobilenet_series_b200` mirrors those and calls ONLY the entry points below
 * (through ctypes).  No torch types cross this boundary: plain device pointers, sizes and a
 * CUDA stream handle.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - tensors are BORROWED: the library never allocates, frees or retains them past the call;
 *   - activations are NHWC bf16 (a [pixels, channels] row-major matrix, pixels = N*H*W,
 *     channels % 8 == 0 so every row is 16-byte aligned); parameters, statistics and
 *     gradients of parameters are fp32 in the reference's own layouts;
 *   - every call enqueues work on `stream` and returns without synchronising; it is
 *     CUDA-graph capturable (no allocation, no host sync);
 *   - return value: 0 on success, a negative YAMB_E* code otherwise; `yamb_last_error()` returns
 *     a thread-local message.  There is NO CPU fallback: without a CUDA device every compute
 *     entry point returns YAMB_ENODEV.
 */
#ifndef YAMB200_H_
#define YAMB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YAMB_OK 0
#define YAMB_EINVAL (-1)
#define YAMB_ENODEV (-2)
#define YAMB_ECUDA (-3)

/* activation codes; reference: models/mobilenet_base.py:461-469 get_active_fn,
 * :70-78 Swish, :81-88 HSwish */
#define YAMB_ACT_NONE 0
#define YAMB_ACT_RELU 1
#define YAMB_ACT_RELU6 2
#define YAMB_ACT_SWISH 3
#define YAMB_ACT_HSWISH 4

typedef void* yamb_stream_t; /* cudaStream_t */

/* ---- BatchNorm bookkeeping attached to a producer kernel ---------------------------------------
 * Forward (train mode), nn.BatchNorm2d semantics (reference: models/mobilenet_base.py:203,417;
 * momentum/eps from apps/mobilenet/mobilenet_v2_mnas.yml:10-11):
 *   scale = gamma*invstd, shift = beta - mean*scale  (what the consumer applies: z = scale*h+shift)
 *   running = (1-m)*running + m*batch  (running_var uses the UNBIASED batch variance)
 *   momentum < 0  => momentum=None in PyTorch: cumulative average with factor 1/num_batches_tracked
 *   (utils/common.py:175-187 bn_calibration). */
typedef struct yamb_bn_fwd {
  float* partials;              /* workspace, >= grid*2*C floats (see yamb_max_ctas) */
  uint32_t* counter;            /* one zero-initialised word, self-resetting */
  const float* gamma;           /* [C] or NULL (=1) */
  const float* beta;            /* [C] or NULL (=0) */
  float eps;
  float momentum;
  float* running_mean;          /* [C] or NULL */
  float* running_var;           /* [C] or NULL */
  int64_t* num_batches_tracked; /* scalar or NULL */
  float* scale;                 /* out [C] */
  float* shift;                 /* out [C] */
  float* mean;                  /* out [C] or NULL (saved for backward) */
  float* invstd;                /* out [C] or NULL (saved for backward) */
  int64_t count;                /* elements per channel (N*H*W) */
} yamb_bn_fwd;

/* Backward of the same BN: given sum(dz), sum(dz*xhat) produce dgamma, dbeta (ACCUMULATED into the
 * gradient buffers) and the affine form of the input gradient  dh = ca*dz + cb*h + cc. */
typedef struct yamb_bn_bwd {
  float* partials;
  uint32_t* counter;
  const float* gamma;  /* [C] or NULL */
  const float* mean;   /* [C] saved by forward */
  const float* invstd; /* [C] saved by forward */
  float* dgamma;       /* [C] += , or NULL */
  float* dbeta;        /* [C] += , or NULL */
  float* ca;           /* out [C] */
  float* cb;           /* out [C] */
  float* cc;           /* out [C] */
  int64_t count;
} yamb_bn_bwd;

/* ---- pointwise (1x1) convolution = GEMM on tcgen05 tensor cores ----------------------------------
 * Replaces nn.Conv2d(k=1) forward, dgrad and wgrad inside the block
 * (reference: models/mobilenet_base.py:391-395 expand, :413 project, :253-257/:284-285 fused).
 *
 *   D[M,N] = A'[M,K] * B'[N,K]^T        fp32 accumulation in TMEM
 *
 * a_mn_major = 0: A is a row-major [M][lda] array (K contiguous); 1: a row-major [K][lda] array
 * (M contiguous) — same for B with N.  With pixels on M this covers
 *   forward : A = activations [pix][Cin],  B = weight [Cout][Cin]           (0,0)
 *   dgrad   : A = dY [pix][Cout],          B = weight [Cout][Cin] as [K][N] (0,1)
 *   wgrad   : A = dY [pix][Cout] as [K][M],B = X [pix][Cin] as [K][N]       (1,1), split-K, epi=2
 * Operand transforms (applied to the tile in shared memory before the MMA):
 *   x' = act(scale[c]*x + shift[c])                     (xform = 1; BN-apply + activation)
 *   x' = scale[c]*x + shift2[c]*x2 + shift[c]           (xform = 2; BN-backward; x2 second tensor)
 * where c indexes the operand's contiguous (channel) dimension.
 * Epilogues:
 *   epi 0: D bf16 = acc (+ residual), optional per-column BN forward statistics (bn_fwd)
 *   epi 1: D bf16 = acc * act'(h_scale*H + h_shift), statistics sum(dz), sum(dz*xhat) (bn_bwd)
 *   epi 2: D fp32 += acc (atomic; split-K partial sums) */
typedef struct yamb_gemm {
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* D; int64_t ldd;
  int32_t epi;
  int32_t a_xform; int32_t a_act;
  const float* a_scale; const float* a_shift; const float* a_scale2;
  const void* A2; int64_t lda2;
  int32_t b_xform; int32_t b_act;
  const float* b_scale; const float* b_shift; const float* b_scale2;
  const void* B2; int64_t ldb2;
  const void* residual; int64_t ldr;   /* epi 0: bf16 [M][ldr], or NULL */
  const yamb_bn_fwd* bn_fwd;           /* epi 0: or NULL */
  const void* H; int64_t ldh;          /* epi 1: bf16 [M][ldh] pre-BN activations */
  const float* h_scale; const float* h_shift; int32_t h_act;
  const yamb_bn_bwd* bn_bwd;           /* epi 1 */
  int32_t max_ctas;                    /* 0 = one CTA per SM */
} yamb_gemm;

int yamb_pointwise_gemm(const yamb_gemm* args, yamb_stream_t stream);

/* number of CTAs the persistent kernels launch at most (sizes the `partials` workspaces) */
int yamb_max_ctas(void);

/* sizeof() of the ABI structs (0 bn_fwd, 1 bn_bwd, 2 gemm, ...) so bindings can self-check */
int yamb_struct_size(int which);

const char* yamb_last_error(void);
int yamb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* YAMB200_H_ */
