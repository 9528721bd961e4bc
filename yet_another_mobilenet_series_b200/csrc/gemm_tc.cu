// Pointwise-convolution GEMM for sm_100a: operand loaders -> shared memory -> tcgen05.mma -> TMEM ->
// epilogue.
//
// One persistent CTA (512 threads) per SM, warp-specialised:
//   warp 0      TMA producer for WIDE TRANSFORMED operands only (128-byte box rows)
//   warp 1      MMA issuer   (one elected lane, tcgen05.mma.cta_group::1.kind::f16, M=128)
//   warp 2      TMEM allocator / deallocator
//   warp 3      idle
//   warps 4-11  epilogue, two warpgroups that alternate tiles (one per TMEM accumulator stage) so
//               the latency chain of one tile hides behind the other; every WARP is independent:
//               tcgen05.ld (its 32 TMEM lanes) -> registers -> fused math -> warp-private swizzled
//               smem -> its own TMA store (box 64 x 32) -> per-column BatchNorm statistics read back
//               from the staged tile.  No CTA-wide barrier in the steady state.
//   warps 12-15 (and 8-11 when the epilogue is light) operand loaders: one warp per pipeline stage;
//               plain operands by cp.async into the swizzled layout, narrow transformed operands
//               through registers, wide transformed operands rewritten in place after the TMA
//               landed them (BN-apply + activation, or the two-source BN-backward affine)
// (TMA tile loads cost 7-16 cycles per box row whatever its width: with the 32-96-byte rows of the
// narrow operands the TMA unit bounded every GEMM — measured with the YAMB_GEMM_DEBUG=512 timers.)
//
// Replaces, behind yamb_pointwise_gemm (include/yamb200.h), the nn.Conv2d(kernel_size=1) forward /
// dgrad / wgrad library calls of the reference block (models/mobilenet_base.py:391-395, :413,
// :253-257, :284-285) together with the BatchNorm statistics passes (:203, :417).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bn_finalize.cuh"
#include "host_util.h"
#include "prims.cuh"

namespace yamb {
// phase timers of the roles (YAMB_GEMM_DEBUG=512) are compiled in only with -DYAMB_GEMM_TIMERS
// (YAMB_GEMM_TIMERS=1 python -c 'import __graft_entry__ as g; g.build(force=True)'): the clock
// reads are scheduling barriers in the hot loops
#ifdef YAMB_GEMM_TIMERS
#define YCLK() clock64()
#else
#define YCLK() 0LL
#endif
#define MBAR_WAIT(bar, par) do { if (p.dbg & 256) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)


constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kPanelBytes64 = 8192;  // 64 rows x 128 B
constexpr int kABytes = 16384;       // 128 x 64 bf16
constexpr int kWarpOutBytes = 4096;  // 32 rows x 64 cols bf16: one warp's staging sub-tile
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // two accumulator stages of 256 columns

struct GemmDev {
  int M, N, K;
  int block_n, m_blocks, n_blocks, num_k_blocks, ksplit, kb_per_split, num_work;
  int a_mn, b_mn;
  int num_stages;
  int a_bytes;      // bytes of the A region of one stage (8 KB when an MN-major A has <= 64 rows)
  int b_bytes;      // bytes of the B region of one stage
  int a2_off;       // offset of A2 region inside a stage (0 = none)
  int b2_off;
  int stage_bytes;
  int a_xform, a_act, b_xform, b_act;
  const float *a_scale, *a_shift, *a_scale2;
  const float *b_scale, *b_shift, *b_scale2;
  int epi;
  int has_residual;
  void* D;
  long long ldd;
  const __nv_bfloat16* side;  // residual (epi 0) or H (epi 1), row-major [M][lds]
  long long lds;
  int out_bufs;               // staging buffers per epilogue warp (1 or 2)
  const float *a_gate, *b_gate;  // optional per-(sample, channel) gates [n][C] (SE)
  long long gate_rps;            // pixels per sample
  int wg2x;                   // 1: warps 8-11 transform (light epilogue), 0: they are epilogue WG 1
  int dbg;                    // YAMB_GEMM_DEBUG bits: 1 skip transform math, 2 skip proxy fence
  unsigned long long* dbg_buf;  // bit 512: per-phase cycle sums of the epilogue warps
  // operand tensors in global memory ([pixels|rows][channels], bf16) for the cp.async loaders
  const __nv_bfloat16 *gA, *gA2, *gB, *gB2;
  long long lda, lda2, ldb, ldb2;
  int a_tma, b_tma;           // wide transformed operand: TMA load + in-place smem transform
  int lgroup;                 // loader warps that share one stage (1, 2 or 4)
  int red_vec;                // epi 2: D rows are 16-byte aligned -> vector reductions
  yamb_bn_fwd bnf;
  int has_bnf;
  const float *h_scale, *h_shift;
  int h_act;
  yamb_bn_bwd bnb;
  int has_bnb;
  // smem offsets (bytes from the 1024-aligned base)
  int off_out, off_hside, off_coef, off_stats, off_bars;
};

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t xdone[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_n() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- explicit shared-space vector accesses (32-bit shared addresses, never generic) ----
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w) : "memory");
}

// Register-staged load + transform of one operand panel by ONE warp: every lane owns the same
// 8-channel column (lc = lane & 7) for the whole panel, so its BatchNorm / affine coefficients sit
// in registers; 8 rows (16-byte vectors) per lane are in flight from global memory at a time, are
// transformed in registers and stored straight into the SWIZZLE_128B layout (no shared-memory
// round trip, no cross-warp synchronisation).
//   mode 1: v = act(s[c]*v + b[c]) [* gate]      mode 2: v = s[c]*v + s2[c]*v2 + b[c]
// Vectors outside the tensor (row >= rlimit or channel >= C) are stored as zero.
template <int MODE, bool SMEM>
__device__ __forceinline__ void gxform_panel(uint32_t panel, uint32_t panel2, const __nv_bfloat16* g1, long long ld1,
                                             const __nv_bfloat16* g2, long long ld2, int rbeg, int rows,
                                             int rlimit, int ln, int cs, const ActParam& ap,
                                             uint32_t tab_s, uint32_t tab_b, uint32_t tab_s2,
                                             int col0, int C, const float* gate, unsigned pixbase,
                                             unsigned rps) {
  // rows in flight per lane (16 B each, x2 sources in mode 2).  From global memory every batch
  // exposes one DRAM latency to this warp: keep the batch as large as the registers allow.
  constexpr int NB = (MODE == 2 && SMEM) ? 4 : 8;
  // 2^cs lanes per row (8 for 64-channel panels; 4 / 2 for narrow operands so that no lane idles)
  const int lc = ln & ((1 << cs) - 1);
  const int RS = 32 >> cs;                // rows covered by one warp-wide access
  const int c0 = col0 + lc * 8;
  const bool cok = c0 < C;
  float2 sc[4], sh[4], s2[4];
  {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0, t0 = a0, t1 = a0;
    if (cok) {
      a0 = lds128f(tab_s + c0 * 4); a1 = lds128f(tab_s + c0 * 4 + 16);
      b0 = lds128f(tab_b + c0 * 4); b1 = lds128f(tab_b + c0 * 4 + 16);
      if (MODE == 2) { t0 = lds128f(tab_s2 + c0 * 4); t1 = lds128f(tab_s2 + c0 * 4 + 16); }
    }
    sc[0] = make_float2(a0.x, a0.y); sc[1] = make_float2(a0.z, a0.w);
    sc[2] = make_float2(a1.x, a1.y); sc[3] = make_float2(a1.z, a1.w);
    sh[0] = make_float2(b0.x, b0.y); sh[1] = make_float2(b0.z, b0.w);
    sh[2] = make_float2(b1.x, b1.y); sh[3] = make_float2(b1.z, b1.w);
    s2[0] = make_float2(t0.x, t0.y); s2[1] = make_float2(t0.z, t0.w);
    s2[2] = make_float2(t1.x, t1.y); s2[3] = make_float2(t1.z, t1.w);
  }
  const uint32_t lo2 = pack_bf16(ap.lo, ap.lo), hi2 = pack_bf16(ap.hi, ap.hi);
  const __nv_bfloat16* p1 = g1 + c0;
  const __nv_bfloat16* p2 = g2 + c0;
  // Straight-line batches (no per-chunk branches: the NB chunk bodies interleave, one warp has
  // NB independent dependency chains in flight; with a branch per chunk a loader warp ran at
  // ~0.1 IPC and the loaders bounded the GEMM).  `lean`: clamp-type activation, no SE gate.
  const bool lean = ap.kind == 0 && gate == nullptr;
  const uint32_t swz = (uint32_t)(lc << 4);
  const float inv_rps = pixbase < (1u << 24) - 1024u ? 1.0f / (float)rps : 0.f;   // 0: divide
#pragma unroll 1
  for (int rb = rbeg + (ln >> cs); rb < rows; rb += RS * NB) {
    uint4 v[NB], w[MODE == 2 ? NB : 1];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int r = rb + RS * j;
      const bool ok = cok && r < rlimit;
      if (SMEM) {   // the TMA producer put the raw tile(s) there: rewrite in place
        const uint32_t off = (uint32_t)(r * 128) + (swz ^ (uint32_t)((r & 7) << 4));
        const bool in = r < rows;   // (always a valid shared address; value unused otherwise)
        v[j] = in ? lds128(panel + off) : make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 2) w[j] = in ? lds128(panel2 + off) : make_uint4(0u, 0u, 0u, 0u);
      } else {
        v[j] = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 2) w[j] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          v[j] = __ldg(reinterpret_cast<const uint4*>(p1 + (long long)r * ld1));
          if (MODE == 2) w[j] = __ldg(reinterpret_cast<const uint4*>(p2 + (long long)r * ld2));
        }
      }
    }
    uint32_t o[NB][4];
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const uint32_t xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        const uint32_t yv[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t1 = ffma2(s2[e], make_float2(bf16lo(yv[e]), bf16hi(yv[e])), sh[e]);
          const float2 d = ffma2(sc[e], make_float2(bf16lo(xv[e]), bf16hi(xv[e])), t1);
          o[j][e] = pack_bf16(d.x, d.y);
        }
      }
    } else if (lean) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const uint32_t xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 d = ffma2(sc[e], make_float2(bf16lo(xv[e]), bf16hi(xv[e])), sh[e]);
          o[j][e] = clamp_bf16x2(pack_bf16(d.x, d.y), lo2, hi2);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) {   // swish / h-swish / SE gate
        const int r = rb + RS * j;
        const uint32_t xv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        float x[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 d = ffma2(sc[e], make_float2(bf16lo(xv[e]), bf16hi(xv[e])), sh[e]);
          x[2 * e] = d.x; x[2 * e + 1] = d.y;
        }
        act_vec<8>(x, ap);
        if (gate != nullptr && cok && r < rlimit) {
          // SE: the gate multiplies the bf16-rounded activation (oracle rounding points)
          // sample index = pixel / pixels-per-sample without an integer division: float reciprocal
          // (pixel indices < 2^24 are exact in fp32) and a one-step correction either way
          const unsigned pix = pixbase + (unsigned)r;
          unsigned smp;
          if (inv_rps != 0.f) {
            smp = (unsigned)__fmul_rz((float)pix, inv_rps);
            if (smp * rps > pix) --smp;
            else if ((smp + 1u) * rps <= pix) ++smp;
          } else {
            smp = pix / rps;
          }
          const float* gr = gate + (size_t)smp * C + c0;
          const float4 q0 = __ldg(reinterpret_cast<const float4*>(gr));
          const float4 q1 = __ldg(reinterpret_cast<const float4*>(gr + 4));
          const float gq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = round_bf16(x[e]) * gq[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = pack_bf16(x[2 * e], x[2 * e + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int r = rb + RS * j;
      const bool ok = cok && r < rlimit;
      if (r < rows)
        sts128(panel + (uint32_t)(r * 128) + (swz ^ (uint32_t)((r & 7) << 4)),
               make_uint4(ok ? o[j][0] : 0u, ok ? o[j][1] : 0u, ok ? o[j][2] : 0u, ok ? o[j][3] : 0u));
    }
  }
}

template <bool kXform, int kEpi>
__global__ void __launch_bounds__(512, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ GemmDev p) {
  // 1024-byte alignment (SWIZZLE_128B atoms) comes from the declaration, not from pointer
  // arithmetic, so the compiler keeps every access in the shared address space (LDS/STS, not
  // generic LD/ST).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  Bars* bars = reinterpret_cast<Bars*>(smem + p.off_bars);
  stat_t* s_stats = reinterpret_cast<stat_t*>(smem + p.off_stats);  // [2][N], double
  float* s_coef = reinterpret_cast<float*>(smem + p.off_coef);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.num_stages;

  // ---- one-time setup ------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    if (kEpi != 2) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&bars->full[i], p.lgroup);   // the loader warps that own the stage arrive
      mbar_init(&bars->xdone[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  // zero the per-CTA statistics accumulators
  if (p.has_bnf || p.has_bnb) {
    for (int i = threadIdx.x; i < 2 * p.N; i += blockDim.x) s_stats[i] = 0.0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const bool use_x = kXform && (p.a_xform != 0 || p.b_xform != 0);

  // 512-thread variant: every thread starts with 128 registers; the control warpgroup hands its
  // surplus to the two epilogue warpgroups (64*4 + 160*8 + 128*4 warps x 32 lanes = 64 Ki regs).
  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 0) {
    // ============ TMA producer: wide transformed operands only (128-byte box rows) ============
    if (lane == 0 && (p.a_tma || p.b_tma)) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0, phase = 0;
      for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
        const int mn = w / p.ksplit, slab = w % p.ksplit;
        const int m_blk = mn / p.n_blocks, n_blk = mn % p.n_blocks;
        const int kb0 = slab * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          MBAR_WAIT(&bars->empty[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sB = sA + p.a_bytes;
          uint32_t tx = 0;
          const int a_panels = p.a_mn ? 2 : 1;
          int a_issue[2] = {0, 0};
          const int b_panels = (p.block_n + 63) / 64;
          if (p.a_tma) {
            if (!p.a_mn) {
              tx += (uint32_t)kABytes;
            } else {
              for (int q = 0; q < 2; ++q)
                if (m_blk * kBlockM + q * 64 < p.M) { a_issue[q] = 1; tx += kPanelBytes64; }
            }
            if (p.a_xform == 2) tx += p.a_mn ? (a_issue[0] + a_issue[1]) * kPanelBytes64 : (uint32_t)kABytes;
          }
          if (p.b_tma) {
            uint32_t tb = 0;
            if (!p.b_mn) {
              tb = (uint32_t)p.block_n * 128u;
            } else {
              for (int q = 0; q < b_panels; ++q)
                if (n_blk * p.block_n + q * 64 < p.N) tb += kPanelBytes64;
            }
            tx += p.b_xform == 2 ? 2 * tb : tb;
          }
          mbar_arrive_expect_tx(&bars->xdone[stage], tx);
          if (p.a_tma) {
            if (!p.a_mn) {
              tma_load_2d(&tmA, &bars->xdone[stage], sA, kb * kBlockK, m_blk * kBlockM);
              if (p.a_xform == 2)
                tma_load_2d(&tmA2, &bars->xdone[stage], sA + p.a2_off, kb * kBlockK, m_blk * kBlockM);
            } else {
              for (int q = 0; q < a_panels; ++q)
                if (a_issue[q]) {
                  tma_load_2d(&tmA, &bars->xdone[stage], sA + q * kPanelBytes64,
                              m_blk * kBlockM + q * 64, kb * kBlockK);
                  if (p.a_xform == 2)
                    tma_load_2d(&tmA2, &bars->xdone[stage], sA + p.a2_off + q * kPanelBytes64,
                                m_blk * kBlockM + q * 64, kb * kBlockK);
                }
            }
          }
          if (p.b_tma) {
            if (!p.b_mn) {
              tma_load_2d(&tmB, &bars->xdone[stage], sB, kb * kBlockK, n_blk * p.block_n);
              if (p.b_xform == 2)
                tma_load_2d(&tmB2, &bars->xdone[stage], sA + p.b2_off, kb * kBlockK,
                            n_blk * p.block_n);
            } else {
              for (int q = 0; q < b_panels; ++q)
                if (n_blk * p.block_n + q * 64 < p.N) {
                  tma_load_2d(&tmB, &bars->xdone[stage], sB + q * kPanelBytes64,
                              n_blk * p.block_n + q * 64, kb * kBlockK);
                  if (p.b_xform == 2)
                    tma_load_2d(&tmB2, &bars->xdone[stage], sA + p.b2_off + q * kPanelBytes64,
                                n_blk * p.block_n + q * 64, kb * kBlockK);
                }
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    const uint32_t idesc = umma_idesc_bf16(kBlockM, p.block_n, p.a_mn, p.b_mn);
    int stage = 0, phase = 0, it = 0;
    long long dbg_mma_full = 0, dbg_mma_empty = 0;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++it) {
      const int slab = w % p.ksplit;
      const int kb0 = slab * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      const int as = it & 1;
      const long long te0 = YCLK();
      MBAR_WAIT(&bars->tmem_empty[as], ((it >> 1) & 1) ^ 1);
      dbg_mma_empty += YCLK() - te0;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccStride);
      for (int kb = kb0; kb < kb1; ++kb) {
        const long long tm0 = YCLK();
        MBAR_WAIT(&bars->full[stage], phase);
        dbg_mma_full += YCLK() - tm0;
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t sB = sA + p.a_bytes;
          const int krem = p.K - kb * kBlockK;
          const int nk = krem >= kBlockK ? 4 : (krem + 15) / 16;
          for (int kk = 0; kk < nk; ++kk) {
            const uint64_t ad = p.a_mn ? umma_smem_desc(sA + kk * 2048, kPanelBytes64, 1024)
                                       : umma_smem_desc(sA + kk * 32, 16, 1024);
            const uint64_t bd = p.b_mn ? umma_smem_desc(sB + kk * 2048, kPanelBytes64, 1024)
                                       : umma_smem_desc(sB + kk * 32, 16, 1024);
            umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          if (kb == kb1 - 1) umma_commit(&bars->tmem_full[as]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    if ((p.dbg & 512) && lane == 0) {
      atomicAdd(p.dbg_buf + 10, (unsigned long long)dbg_mma_full);
      atomicAdd(p.dbg_buf + 11, (unsigned long long)dbg_mma_empty);
    }
   }
  } else if (warp < 8 || (warp < 12 && !(kXform && p.wg2x))) {
    // ======================================= epilogue =======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ew = warp - 4;           // 0..7
    const int wg = ew >> 2;            // epilogue warpgroup = TMEM accumulator stage it serves
    const int q = warp & 3;            // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;     // row inside the 128-row tile
    uint8_t* s_out = smem + p.off_out + (size_t)ew * p.out_bufs * kWarpOutBytes;
    uint8_t* s_h = smem + p.off_hside + (size_t)ew * kWarpOutBytes;  // epi 1: raw H rows
    // per-column coefficient tables for epi 1:  [h_scale | h_shift | mean | invstd] x N
    const float* cz_s = s_coef;
    const float* cz_t = s_coef + p.N;
    const float* cz_m = s_coef + 2 * p.N;
    const float* cz_r = s_coef + 3 * p.N;
    if (kEpi == 1) {
      float* wr = s_coef;
      const int n_epi_thr = (kXform && p.wg2x) ? 128 : 256;
      for (int i = threadIdx.x - 128; i < p.N; i += n_epi_thr) {
        wr[i] = p.h_scale[i];
        wr[p.N + i] = p.h_shift[i];
        wr[2 * p.N + i] = p.bnb.mean[i];
        wr[3 * p.N + i] = p.bnb.invstd[i];
      }
      named_bar_sync(1, n_epi_thr);
    }
    const int n_epi_wg = (kXform && p.wg2x) ? 1 : 2;
    uint32_t sub_count = 0;   // running sub-tile counter of this warp -> staging buffer parity
    const bool side_in = (kEpi == 1) || p.has_residual;
    const ActParam hap = make_act(kEpi == 1 ? p.h_act : ACT_NONE);
    // Column statistics: with a single n-block every lane owns the same 2 columns of sub-tile j in
    // every tile, so the sums live in registers for the whole kernel (shared-memory float atomics
    // are CAS loops and serialise the 8 epilogue warps); flushed once at the end.
    const bool reg_stats = (p.has_bnf || p.has_bnb) && p.n_blocks == 1;
    float racc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) racc[j][e] = 0.f;
    int it = 0;
    long long dbg_t[6] = {0, 0, 0, 0, 0, 0};
    const long long dbg_start = YCLK();
    // Side operand (residual / H) rows come straight from global memory (each thread owns one row:
    // 8 x 16 B per 64-column sub-tile).  They are requested ONE SUB-TILE AHEAD, into the registers
    // the previous sub-tile has just finished with: issued at the top of a sub-tile their ~2 us
    // latency was fully exposed (the "convert" phase of the dgrad epilogue was 4000 cycles).
    uint4 sv[8];
    auto load_side = [&](int w2, int sub2) {
      const int mn2 = w2 / p.ksplit;
      const int m2 = mn2 / p.n_blocks, n2 = mn2 % p.n_blocks;
      const int grow2 = m2 * kBlockM + row;
      const int col2 = n2 * p.block_n + sub2 * 64;
      const __nv_bfloat16* srow = p.side + (size_t)grow2 * p.lds + col2;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        sv[ch] = make_uint4(0u, 0u, 0u, 0u);
        if (grow2 < p.M && col2 + ch * 8 < p.N)
          sv[ch] = __ldg(reinterpret_cast<const uint4*>(srow + ch * 8));
      }
    };
    const int w_step = (n_epi_wg == 2 ? 2 : 1) * (int)gridDim.x;
    if (side_in) {
      const int w0 = blockIdx.x + ((n_epi_wg == 2 && wg == 1) ? (int)gridDim.x : 0);
      if (w0 < p.num_work) load_side(w0, 0);
    }
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++it) {
      if (n_epi_wg == 2 && (it & 1) != wg) continue;  // the other warpgroup owns this tile
      const int mn = w / p.ksplit;
      const int m_blk = mn / p.n_blocks, n_blk = mn % p.n_blocks;
      const int as = it & 1;
      const int grow = m_blk * kBlockM + row;
      const bool row_ok = grow < p.M;
      // sub-tiles of 64 columns; skip the ones that lie entirely beyond N (last n-block)
      const int n_sub = min((p.block_n + 63) / 64, (p.N - n_blk * p.block_n + 63) / 64);
      long long tq0 = YCLK();
      if (p.dbg & 256) mbar_wait_spin(&bars->tmem_full[as], (it >> 1) & 1);
      else mbar_wait_relaxed(&bars->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      dbg_t[0] += YCLK() - tq0;
      // NOT unrolled: one copy of the sub-tile body keeps the epilogue inside the instruction
      // cache (4 unrolled copies x 3 epilogue kinds were ~20k instructions; "no instruction" was
      // the top stall reason and a sub-tile took ~4000 cycles)
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        const int col0 = n_blk * p.block_n + sub * 64;  // global column of this sub-tile
        const uint32_t taddr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccStride + sub * 64);
        uint32_t acc[2][32];
        long long tq1 = YCLK();
        tmem_ld_32x32(taddr, acc[0]);
        tmem_ld_32x32(taddr + 32, acc[1]);
        tmem_ld_wait();
        dbg_t[1] += YCLK() - tq1;
        tq1 = YCLK();
        if (sub == n_sub - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          mbar_arrive(&bars->tmem_empty[as]);
        }
        if (kEpi == 2) {
          // split-K partial sums: fp32 atomic accumulate into D[M][ldd]
          // A thread owns one output row: 4 consecutive columns go out as ONE 16-byte vector
          // reduction (red.global.add.v4.f32, sm_90+).  Scalar reds were 32 L2 atomic transactions
          // per warp instruction (lanes = rows, 4 KB apart) and the 7x7 / 14x14 wgrads spent ~80 %
          // of their time draining them (r2 timers: 17 us of CTA activity in a 100 us kernel).
          if (row_ok) {
            float* drow = reinterpret_cast<float*>(p.D) + (size_t)grow * p.ldd;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int c = col0 + h * 32 + j;
                if (c < p.N && (sub * 64 + h * 32 + j) < p.block_n) {   // N, block_n: multiples of 8
                  if (p.red_vec) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c),
                                 "f"(__uint_as_float(acc[h][j])), "f"(__uint_as_float(acc[h][j + 1])),
                                 "f"(__uint_as_float(acc[h][j + 2])), "f"(__uint_as_float(acc[h][j + 3]))
                                 : "memory");
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(drow + c + e),
                                   "f"(__uint_as_float(acc[h][j + e]))
                                   : "memory");
                  }
                }
              }
          }
          continue;
        }
        uint8_t* sO = s_out + (p.out_bufs == 2 ? (sub_count & 1) : 0) * kWarpOutBytes;
        // the TMA store that last read this staging buffer must have drained
        if (lane == 0) {
          if (p.out_bufs == 2) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        __syncwarp();
        dbg_t[2] += YCLK() - tq1;
        tq1 = YCLK();
        // Straight-line conversion of the 8 chunks (8 columns each): the mode decisions are taken
        // ONCE per sub-tile, outside the unrolled chunk loop, so the chunk bodies interleave
        // (with a branch per chunk this phase took ~4700 cycles per sub-tile in the dgrad epilogue).
        auto stage_chunk = [&](int ch, const float (&v)[8]) {
          const int pc = ch ^ (lane & 7);
          *reinterpret_cast<uint4*>(sO + lane * 128 + (pc << 4)) =
              make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                         pack_bf16(v[6], v[7]));
        };
        if (kEpi == 1) {
          auto dz_chunks = [&](const bool clampk) {   // inlined twice, `clampk` a constant in each
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[ch >> 2][(ch & 3) * 8 + e]);
              const uint32_t sw[4] = {sv[ch].x, sv[ch].y, sv[ch].z, sv[ch].w};
              // columns past N: read the last valid coefficients (the values are never stored)
              const int cb = min(col0 + ch * 8, p.N - 8);
              const float4 s0 = *reinterpret_cast<const float4*>(cz_s + cb);
              const float4 s1 = *reinterpret_cast<const float4*>(cz_s + cb + 4);
              const float4 t0 = *reinterpret_cast<const float4*>(cz_t + cb);
              const float4 t1 = *reinterpret_cast<const float4*>(cz_t + cb + 4);
              const float zs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float zt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
              float zz[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                zz[2 * e] = fmaf(zs[2 * e], bf16lo(sw[e]), zt[2 * e]);
                zz[2 * e + 1] = fmaf(zs[2 * e + 1], bf16hi(sw[e]), zt[2 * e + 1]);
              }
              if (clampk) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (zz[e] > hap.lo && zz[e] < hap.hi) ? v[e] : 0.f;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= act_bwd(zz[e], p.h_act);
              }
              *reinterpret_cast<uint4*>(s_h + lane * 128 + ((ch ^ (lane & 7)) << 4)) = sv[ch];
              stage_chunk(ch, v);
            }
          };
          if (hap.kind == 0) dz_chunks(true);
          else dz_chunks(false);
        } else if (side_in) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[ch >> 2][(ch & 3) * 8 + e]);
            const uint32_t sw[4] = {sv[ch].x, sv[ch].y, sv[ch].z, sv[ch].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] += bf16lo(sw[e]);
              v[2 * e + 1] += bf16hi(sw[e]);
            }
            stage_chunk(ch, v);
          }
        } else {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[ch >> 2][(ch & 3) * 8 + e]);
            stage_chunk(ch, v);
          }
        }
        dbg_t[3] += YCLK() - tq1;
        tq1 = YCLK();
        if (p.dbg & 128) {
          // experiment: coalesced st.global from the staged tile instead of a TMA store
          __syncwarp();
          __nv_bfloat16* Dg = reinterpret_cast<__nv_bfloat16*>(p.D);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (lane >> 3) + 4 * i;
            const int gr = m_blk * kBlockM + q * 32 + r;
            const int gc = col0 + (lane & 7) * 8;
            const uint4 v = *reinterpret_cast<const uint4*>(sO + r * 128 + (((lane & 7) ^ (r & 7)) << 4));
            if (gr < p.M && gc < p.N && (sub * 64 + (lane & 7) * 8) < p.block_n)
              *reinterpret_cast<uint4*>(Dg + (size_t)gr * p.ldd + gc) = v;
          }
        } else {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(p.dbg & 64)) {
            tma_store_2d(&tmD, sO, col0, m_blk * kBlockM + q * 32);
            tma_store_commit();
          }
        }
        if (side_in) {
          // sv is dead and the store's proxy fence (a MEMBAR that would wait for these loads) is
          // behind us: request the next sub-tile's side rows now; they land during the
          // statistics pass and the next accumulator load
          if (sub + 1 < n_sub) load_side(w, sub + 1);
          else if (w + w_step < p.num_work) load_side(w + w_step, 0);
        }
        dbg_t[4] += YCLK() - tq1;
        tq1 = YCLK();
        // ---- per-column statistics of this warp's 32 rows of the bf16-rounded output ----
        if (p.has_bnf || p.has_bnb) {
          const int c = col0 + 2 * lane;  // column pair owned by this lane
          if (c < p.N && (sub * 64 + 2 * lane) < p.block_n) {
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            const int rmax = min(32, p.M - (m_blk * kBlockM + q * 32));
            if (p.has_bnf) {
              if (rmax == 32) {
                // full sub-tile: 4 independent accumulator sets (a 32-deep dependent chain of
                // LDS -> FADD/FFMA was ~1500 cycles on the epilogue's critical path per sub-tile);
                // packed fp32x2: one FFMA2 adds the column pair, one squares-and-adds it
                float2 sA = make_float2(0.f, 0.f), sB = sA, sC = sA, sD = sA;
                float2 qA = sA, qB = sA, qC = sA, qD = sA;
                const float2 one2 = make_float2(1.f, 1.f);
                auto rowf = [&](int r, float2& S, float2& Q) {
                  const uint32_t u = *reinterpret_cast<const uint32_t*>(
                      sO + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
                  const float2 v = make_float2(bf16lo(u), bf16hi(u));
                  S = ffma2(one2, v, S);
                  Q = ffma2(v, v, Q);
                };
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                  rowf(r, sA, qA);
                  rowf(r + 1, sB, qB);
                  rowf(r + 2, sC, qC);
                  rowf(r + 3, sD, qD);
                }
                s0 = (sA.x + sB.x) + (sC.x + sD.x); s1 = (sA.y + sB.y) + (sC.y + sD.y);
                q0 = (qA.x + qB.x) + (qC.x + qD.x); q1 = (qA.y + qB.y) + (qC.y + qD.y);
              } else {
                for (int r = 0; r < rmax; ++r) {
                  const uint32_t u = *reinterpret_cast<const uint32_t*>(
                      sO + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
                  const float a = bf16lo(u), b = bf16hi(u);
                  s0 += a; s1 += b;
                  q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
                }
              }
            } else {
              // sum(dz), sum(dz * xhat) with xhat = h * r - m * r: three packed FFMA2 per row
              const float2 r2 = make_float2(cz_r[c], cz_r[c + 1]);
              const float2 nmr2 = make_float2(-cz_m[c] * r2.x, -cz_m[c + 1] * r2.y);
              const float2 one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
              float2 sA = zero2, sB = zero2, sC = zero2, sD = zero2;
              float2 qA = zero2, qB = zero2, qC = zero2, qD = zero2;
              auto rowacc = [&](int r, float2& S, float2& Q) {
                const int off = r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2);
                const uint32_t u = *reinterpret_cast<const uint32_t*>(sO + off);
                const uint32_t hh = *reinterpret_cast<const uint32_t*>(s_h + off);
                const float2 a = make_float2(bf16lo(u), bf16hi(u));
                const float2 xh = ffma2(make_float2(bf16lo(hh), bf16hi(hh)), r2, nmr2);
                S = ffma2(one2, a, S);
                Q = ffma2(a, xh, Q);
              };
              int r = 0;
#pragma unroll 2
              for (; r + 3 < rmax; r += 4) {   // four independent accumulator sets
                rowacc(r, sA, qA);
                rowacc(r + 1, sB, qB);
                rowacc(r + 2, sC, qC);
                rowacc(r + 3, sD, qD);
              }
              for (; r < rmax; ++r) rowacc(r, sA, qA);
              s0 = (sA.x + sB.x) + (sC.x + sD.x); s1 = (sA.y + sB.y) + (sC.y + sD.y);
              q0 = (qA.x + qB.x) + (qC.x + qD.x); q1 = (qA.y + qB.y) + (qC.y + qD.y);
            }
            if (reg_stats) {
              switch (sub) {  // constant register indices in every case
                case 0: racc[0][0] += s0; racc[0][1] += s1; racc[0][2] += q0; racc[0][3] += q1; break;
                case 1: racc[1][0] += s0; racc[1][1] += s1; racc[1][2] += q0; racc[1][3] += q1; break;
                case 2: racc[2][0] += s0; racc[2][1] += s1; racc[2][2] += q0; racc[2][3] += q1; break;
                default: racc[3][0] += s0; racc[3][1] += s1; racc[3][2] += q0; racc[3][3] += q1; break;
              }
            } else {
              atomicAdd(&s_stats[c], (stat_t)s0);
              atomicAdd(&s_stats[c + 1], (stat_t)s1);
              atomicAdd(&s_stats[p.N + c], (stat_t)q0);
              atomicAdd(&s_stats[p.N + c + 1], (stat_t)q1);
            }
          }
          __syncwarp();  // s_h / sO reads done before the next sub-tile overwrites them
        }
        dbg_t[5] += YCLK() - tq1;
        ++sub_count;
      }
    }
    if ((p.dbg & 512) && lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(p.dbg_buf + i, (unsigned long long)dbg_t[i]);
      atomicAdd(p.dbg_buf + 6, (unsigned long long)(YCLK() - dbg_start));
      atomicAdd(p.dbg_buf + 7, 1ull);
    }
    if (reg_stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j * 64 + 2 * lane;
        if (c < p.N) {
          atomicAdd(&s_stats[c], (stat_t)racc[j][0]);
          atomicAdd(&s_stats[c + 1], (stat_t)racc[j][1]);
          atomicAdd(&s_stats[p.N + c], (stat_t)racc[j][2]);
          atomicAdd(&s_stats[p.N + c + 1], (stat_t)racc[j][3]);
        }
      }
    }
    if (lane == 0 && kEpi != 2) tma_store_wait_all<0>();
  } else {
    // ============================ operand loaders (+ transform) ============================
    // Warps 12-15 (t 0..127) plus, when wg2x, warps 8-11 (t 128..255) bring the operand tiles into
    // shared memory: plain operands with cp.async (16-byte chunks written straight into the
    // SWIZZLE_128B layout the UMMA descriptors expect, zero-filled outside the tensors), narrow
    // transformed operands through registers, wide transformed operands by rewriting the tile the
    // TMA producer (warp 0) landed; then they publish the stage to the MMA warp.  (TMA tile loads cost ~7-16 cycles per box ROW whatever its width:
    // with 32-96-byte rows of the narrow operands the TMA unit, not HBM, bounded every GEMM.)
    // registers: 4 control warps x 56 + (epilogue + loader) warps x 152 = 64 Ki / 32
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int nxt = p.wg2x ? 256 : 128;
    const int t = warp >= 12 ? threadIdx.x - 384 : threadIdx.x - 256 + 128;
    const int G = p.lgroup;                       // warps sharing a stage (rows split G ways)
    const int NW = (nxt >> 5) / G, lw = (t >> 5) / G, ls = (t >> 5) % G, ln = t & 31;
    const int Ca = (kXform && p.a_xform) ? (p.a_mn ? p.M : p.K) : 0;
    const int Cb = (kXform && p.b_xform) ? (p.b_mn ? p.N : p.K) : 0;
    float* xa = s_coef + (kEpi == 1 ? 4 * p.N : 0);
    float* xb = xa + 3 * Ca;
    if (use_x) {
      // coefficient tables in smem: A: [scale|shift|scale2] x Ca, then B likewise
      for (int i = t; i < Ca; i += nxt) {
        xa[i] = p.a_scale[i];
        xa[Ca + i] = p.a_shift[i];
        xa[2 * Ca + i] = p.a_xform == 2 ? p.a_scale2[i] : 0.f;
      }
      for (int i = t; i < Cb; i += nxt) {
        xb[i] = p.b_scale[i];
        xb[Cb + i] = p.b_shift[i];
        xb[2 * Cb + i] = p.b_xform == 2 ? p.b_scale2[i] : 0.f;
      }
      named_bar_sync(2, nxt);
    }
    long long dbg_l[4] = {0, 0, 0, 0};
    const ActParam apa = make_act((kXform && p.a_xform == 1) ? p.a_act : ACT_NONE);
    const ActParam apb = make_act((kXform && p.b_xform == 1) ? p.b_act : ACT_NONE);
    const int na = p.a_mn ? 2 : 1;                                         // panels of A
    const int nb = p.b_mn ? (p.block_n + 63) / 64 : (p.block_n + 127) / 128;  // panels of B

    // (tile, k-block) cursor over this CTA's work
    struct Cur { int w, kb, kb1, m_blk, n_blk; };
    const bool simple_work = p.ksplit == 1 && p.n_blocks == 1;   // one work item = one m-block
    auto cur_set = [&](Cur& c) {
      if (c.w < p.num_work) {
        if (simple_work) {
          c.m_blk = c.w; c.n_blk = 0; c.kb = 0; c.kb1 = p.num_k_blocks;
        } else {
          const int mn = c.w / p.ksplit, slab = c.w % p.ksplit;
          c.m_blk = mn / p.n_blocks; c.n_blk = mn % p.n_blocks;
          c.kb = slab * p.kb_per_split;
          c.kb1 = min(c.kb + p.kb_per_split, p.num_k_blocks);
        }
      }
    };
    auto cur_next = [&](Cur& c) {
      if (++c.kb >= c.kb1) { c.w += gridDim.x; cur_set(c); }
    };
    // geometry of panel pi (A panels first, then B) of the k-block at cursor c
    struct Pan { uint32_t off; int rows, row0, col0, rlimit, climit, logR, cshift; bool isA, mn; };
    auto panel_of = [&](const Cur& c, int pi) {
      Pan g;
      g.isA = pi < na;
      const int q = g.isA ? pi : pi - na;
      g.mn = g.isA ? (p.a_mn != 0) : (p.b_mn != 0);
      g.off = (g.isA ? 0u : (uint32_t)p.a_bytes) + (uint32_t)q * (g.mn ? kPanelBytes64 : 128 * 128);
      g.logR = g.mn ? 6 : 7;
      if (!g.mn) {  // K-major: rows are M (A) or N (B), channels along K
        g.row0 = g.isA ? c.m_blk * kBlockM : c.n_blk * p.block_n + q * 128;
        g.col0 = c.kb * kBlockK; g.climit = p.K;
        g.rlimit = g.isA ? min(128, p.M - g.row0)
                         : min(128, min(p.block_n - q * 128, p.N - g.row0));
        g.rows = g.rlimit;     // rows past the limit only feed outputs that are never stored
      } else {      // MN-major: rows are K (pixels), 64 channels of M/N per panel
        g.row0 = c.kb * kBlockK;
        g.col0 = (g.isA ? c.m_blk * kBlockM : c.n_blk * p.block_n) + q * 64;
        g.climit = g.isA ? p.M : p.N;
        g.rlimit = g.col0 < g.climit ? min(64, p.K - g.row0) : 0;
        // rows past K must read as zero (they are accumulated); a panel wholly past M / N only
        // feeds outputs that are never stored: leave it alone
        g.rows = g.col0 < g.climit ? 64 : 0;
      }
      if (g.rlimit < 0) g.rlimit = 0;
      if (g.rows < 0) g.rows = 0;
      // 16-byte chunk slots per row that the MMA can read: narrow operands (K or C <= 32) take
      // 2 or 4 lanes per row instead of 8
      int ncs = (g.climit - g.col0 + 7) >> 3;
      if (!g.mn) ncs = (ncs + 1) & ~1;   // K-steps of 16 columns: the odd chunk must read as zero
      g.cshift = ncs <= 2 ? 1 : (ncs <= 4 ? 2 : 3);
      return g;
    };
    // cp.async one panel (coalesced: 8 consecutive threads fetch the 128 bytes of one row)
    auto load_panel = [&](uint32_t dst, const __nv_bfloat16* src, long long ld, const Pan& g) {
      const int rpw = ((1 << g.logR) / G);          // rows of the panel per warp of the group
      const int rend = min(g.rows, (ls + 1) * rpw);
      const int total = rend << g.cshift;
      const __nv_bfloat16* base = src + (long long)g.row0 * ld + g.col0;
#pragma unroll 4
      for (int i = ((ls * rpw) << g.cshift) + ln; i < total; i += 32) {
        const int r = i >> g.cshift, lc = i & ((1 << g.cshift) - 1);
        const bool ok = r < g.rlimit && g.col0 + lc * 8 < g.climit;
        const __nv_bfloat16* sp = ok ? base + (long long)r * ld + lc * 8 : src;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                         dst + (uint32_t)(r * 128 + ((lc ^ (r & 7)) << 4))),
                     "l"(sp), "r"(ok ? 16 : 0)
                     : "memory");
      }
    };
    auto issue = [&](const Cur& c, int stage, int rnd) {
      const long long tl0 = YCLK();
      MBAR_WAIT(&bars->empty[stage], (rnd & 1) ^ 1);
      dbg_l[0] += YCLK() - tl0;
      const uint32_t sA = smem_u32(smem + (size_t)stage * p.stage_bytes);
      for (int pi = 0; pi < na + nb; ++pi) {
        const Pan g = panel_of(c, pi);
        if (kXform && (g.isA ? p.a_xform : p.b_xform) != 0) continue;   // register path (consume)
        load_panel(sA + g.off, g.isA ? p.gA : p.gB, g.isA ? p.lda : p.ldb, g);
      }
    };
    auto consume = [&](const Cur& c, int stage, int rnd) {
      const long long tl3 = YCLK();
      if (kXform && use_x) {
        const uint32_t sA = smem_u32(smem + (size_t)stage * p.stage_bytes);
        // narrow transformed operands: registers (few bytes per row, latency hidden by the batch)
#pragma unroll 1
        for (int pi = 0; pi < na + nb; ++pi) {
          const Pan g = panel_of(c, pi);
          const int mode = g.isA ? p.a_xform : p.b_xform;
          if (mode == 0 || (g.isA ? p.a_tma : p.b_tma)) continue;
          const ActParam& ap = g.isA ? apa : apb;
          const float* tab = g.isA ? xa : xb;
          const int Ct = g.isA ? Ca : Cb;
          const uint32_t t_s = smem_u32(tab), t_b = smem_u32(tab + Ct), t_s2 = smem_u32(tab + 2 * Ct);
          const float* gate = (mode == 1 && (g.mn || g.isA)) ? (g.isA ? p.a_gate : p.b_gate) : nullptr;
          const __nv_bfloat16* g1 = (g.isA ? p.gA : p.gB) + (long long)g.row0 * (g.isA ? p.lda : p.ldb);
          const __nv_bfloat16* g2 = mode == 2
              ? (g.isA ? p.gA2 : p.gB2) + (long long)g.row0 * (g.isA ? p.lda2 : p.ldb2) : g1;
          if (mode == 2)
            gxform_panel<2, false>(sA + g.off, 0u, g1, g.isA ? p.lda : p.ldb, g2, g.isA ? p.lda2 : p.ldb2,
                                   ls * ((1 << g.logR) / G), min(g.rows, (ls + 1) * ((1 << g.logR) / G)), g.rlimit, ln, g.cshift, ap, t_s, t_b, t_s2, g.col0, g.climit, gate,
                                   (unsigned)g.row0, (unsigned)p.gate_rps);
          else
            gxform_panel<1, false>(sA + g.off, 0u, g1, g.isA ? p.lda : p.ldb, g2, g.isA ? p.lda2 : p.ldb2,
                                   ls * ((1 << g.logR) / G), min(g.rows, (ls + 1) * ((1 << g.logR) / G)), g.rlimit, ln, g.cshift, ap, t_s, t_b, t_s2, g.col0, g.climit, gate,
                                   (unsigned)g.row0, (unsigned)p.gate_rps);
        }
        // wide transformed operands: the TMA producer loaded them; rewrite the tile in place
        if (p.a_tma || p.b_tma) {
          const long long tl1 = YCLK();
          MBAR_WAIT(&bars->xdone[stage], rnd & 1);
          dbg_l[1] += YCLK() - tl1;
          if (!(p.dbg & 1)) {
#pragma unroll 1
            for (int pi = 0; pi < na + nb; ++pi) {
              const Pan g = panel_of(c, pi);
              const int mode = g.isA ? p.a_xform : p.b_xform;
              if (mode == 0 || !(g.isA ? p.a_tma : p.b_tma)) continue;
              const ActParam& ap = g.isA ? apa : apb;
              const float* tab = g.isA ? xa : xb;
              const int Ct = g.isA ? Ca : Cb;
              const uint32_t t_s = smem_u32(tab), t_b = smem_u32(tab + Ct), t_s2 = smem_u32(tab + 2 * Ct);
              const float* gate = (mode == 1 && (g.mn || g.isA)) ? (g.isA ? p.a_gate : p.b_gate) : nullptr;
              const uint32_t op2 = sA + (g.isA ? p.a2_off : p.b2_off) +
                                   (g.off - (g.isA ? 0u : (uint32_t)p.a_bytes));
              // rows the TMA zero-filled (outside the tensor) must stay zero: limit = rlimit
              if (mode == 2)
                gxform_panel<2, true>(sA + g.off, op2, nullptr, 0, nullptr, 0, ls * ((1 << g.logR) / G),
                                      min(g.rlimit, (ls + 1) * ((1 << g.logR) / G)), g.rlimit, ln, 3, ap,
                                      t_s, t_b, t_s2, g.col0, g.climit, gate, (unsigned)g.row0,
                                      (unsigned)p.gate_rps);
              else
                gxform_panel<1, true>(sA + g.off, op2, nullptr, 0, nullptr, 0, ls * ((1 << g.logR) / G),
                                      min(g.rlimit, (ls + 1) * ((1 << g.logR) / G)), g.rlimit, ln, 3, ap,
                                      t_s, t_b, t_s2, g.col0, g.climit, gate, (unsigned)g.row0,
                                      (unsigned)p.gate_rps);
            }
          }
        }
      }
      const long long tl2 = YCLK();
      dbg_l[2] += tl2 - tl3;
      cp_async_wait_n<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (ln == 0) mbar_arrive(&bars->full[stage]);
      dbg_l[3] += YCLK() - tl2;
    };

    // One loader WARP fills a whole stage, and every stage has ONE owner warp (stage s belongs to
    // warp s mod NW): the owner walks the rounds of its stage in order, so a parity wait on
    // empty[s] can never alias a phase two rounds away.  Each warp has a single stage in flight, so
    // the proxy fence before publishing it (a MEMBAR that waits for ALL of the thread's
    // outstanding cp.async) never waits for younger copies; the CTA has min(NW, S) stages in flight.
    Cur c;
    c.w = blockIdx.x;
    cur_set(c);
    const long long dbg_l0 = YCLK();
    // n = running k-block index, st = n % S, rnd = n / S (no divisions in the loop)
    int st = 0, rnd = 0;
    for (; c.w < p.num_work; cur_next(c)) {
      if (st % NW == lw) {
        issue(c, st, rnd);
        cp_async_commit();
        consume(c, st, rnd);
      }
      if (++st == S) { st = 0; ++rnd; }
    }
    if ((p.dbg & 512) && ln == 0) {
      atomicAdd(p.dbg_buf + 12, (unsigned long long)dbg_l[0]);
      atomicAdd(p.dbg_buf + 13, (unsigned long long)dbg_l[1]);
      atomicAdd(p.dbg_buf + 14, (unsigned long long)dbg_l[2]);
      atomicAdd(p.dbg_buf + 15, (unsigned long long)dbg_l[3]);
      atomicAdd(p.dbg_buf + 8, (unsigned long long)(YCLK() - dbg_l0));
      atomicAdd(p.dbg_buf + 9, 1ull);
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.has_bnf) {
    if (publish_partials(s_stats, p.N, p.bnf.partials, p.bnf.counter)) {
      bn_fwd_finalize(p.bnf, p.N);
      __syncthreads();
      if (threadIdx.x == 0) *p.bnf.counter = 0;
    }
  } else if (p.has_bnb) {
    if (publish_partials(s_stats, p.N, p.bnb.partials, p.bnb.counter)) {
      bn_bwd_finalize(p.bnb, p.N);
      __syncthreads();
      if (threadIdx.x == 0) *p.bnb.counter = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int make_map_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer,
                       uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (!encode) return set_error(YAMB_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15))
    return set_error(YAMB_EINVAL, "TMA operand must be 16-byte aligned with a 16-byte row pitch");
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(YAMB_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int pick_block_n(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  // multi-block: block_n must be a multiple of 64 so staging sub-tiles never straddle blocks
  int best = 256, best_waste = 1 << 30;
  for (int bn = 256; bn >= 128; bn -= 64) {
    int blocks = (N + bn - 1) / bn;
    int waste = blocks * bn - N;
    if (waste < best_waste) { best_waste = waste; best = bn; }
  }
  return best;
}

int gemm_launch(const yamb_gemm* a, cudaStream_t stream) {
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(YAMB_EINVAL, "bad GEMM shape");
  if ((a->N % 8) || (a->a_mn_major ? (a->M % 8) : (a->K % 8)) || (a->b_mn_major ? 0 : (a->K % 8)))
    return set_error(YAMB_EINVAL, "GEMM channel dims must be multiples of 8 (M=%d N=%d K=%d)",
                     a->M, a->N, a->K);
  if (a->epi < 0 || a->epi > 2) return set_error(YAMB_EINVAL, "bad epilogue");
  int dev_ctas = max_ctas();
  if (dev_ctas <= 0) return set_error(YAMB_ENODEV, "no CUDA device");

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.block_n = pick_block_n(a->N);
  p.m_blocks = (a->M + kBlockM - 1) / kBlockM;
  p.n_blocks = (a->N + p.block_n - 1) / p.block_n;
  p.num_k_blocks = (a->K + kBlockK - 1) / kBlockK;
  p.epi = a->epi;
  p.a_xform = a->a_xform; p.a_act = a->a_act;
  p.b_xform = a->b_xform; p.b_act = a->b_act;
  p.a_scale = a->a_scale; p.a_shift = a->a_shift; p.a_scale2 = a->a_scale2;
  p.b_scale = a->b_scale; p.b_shift = a->b_shift; p.b_scale2 = a->b_scale2;
  if ((p.a_xform && (!a->a_scale || !a->a_shift)) || (p.b_xform && (!a->b_scale || !a->b_shift)))
    return set_error(YAMB_EINVAL, "operand transform without coefficients");
  if ((p.a_xform == 2 && (!a->A2 || !a->a_scale2)) || (p.b_xform == 2 && (!a->B2 || !a->b_scale2)))
    return set_error(YAMB_EINVAL, "two-source transform without second tensor");
  p.has_residual = (a->epi == 0 && a->residual) ? 1 : 0;
  static const int env_dbg = [] { const char* d = getenv("YAMB_GEMM_DEBUG"); return d ? atoi(d) : 0; }();
  p.dbg = env_dbg;
  p.dbg_buf = nullptr;
  if (p.dbg & 512) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 16 * sizeof(unsigned long long), stream);
    p.dbg_buf = dbuf;
  }
  // transform-heavy / epilogue-light launches give warps 8-11 to the transform
  // warps 8-11 load/transform instead of running a second epilogue warpgroup when the epilogue
  // is light relative to the operand work: split-K wgrad (no store at all) or a forward/dgrad
  // store of <= 192 columns fed by more than one k-block per tile
  p.wg2x = ((a->a_xform || a->b_xform) &&
            (a->epi == 2 || (a->epi == 0 && p.block_n <= 192 && p.num_k_blocks >= 2))) ? 1 : 0;
  if (p.dbg & 16) p.wg2x = 0;
  if (p.dbg & 32) p.wg2x = (a->a_xform || a->b_xform) ? 1 : 0;
  p.a_gate = a->a_xform == 1 ? a->a_gate : nullptr;
  p.b_gate = a->b_xform == 1 ? a->b_gate : nullptr;
  p.gate_rps = a->gate_rows_per_sample > 0 ? a->gate_rows_per_sample : 1;
  if ((p.a_gate || p.b_gate) && a->gate_rows_per_sample <= 0)
    return set_error(YAMB_EINVAL, "gate without gate_rows_per_sample");
  if (p.b_gate && !p.b_mn) return set_error(YAMB_EINVAL, "b_gate needs an MN-major B (rows = pixels)");
  p.D = a->D; p.ldd = a->ldd;
  p.red_vec = (a->epi == 2 && !(reinterpret_cast<uintptr_t>(a->D) & 15) && (a->ldd % 4) == 0) ? 1 : 0;
  if (p.dbg & 4) p.red_vec = 0;
  if (a->epi == 0 && a->bn_fwd) { p.bnf = *a->bn_fwd; p.has_bnf = 1; }
  if (a->epi == 1) {
    if (!a->H || !a->h_scale || !a->h_shift || !a->bn_bwd)
      return set_error(YAMB_EINVAL, "dz epilogue needs H, h_scale, h_shift, bn_bwd");
    p.bnb = *a->bn_bwd; p.has_bnb = 1;
    p.h_scale = a->h_scale; p.h_shift = a->h_shift; p.h_act = a->h_act;
  }
  // split-K only for the atomic epilogue
  const int mn_tiles = p.m_blocks * p.n_blocks;
  int ctas = dev_ctas;
  if (a->max_ctas > 0 && a->max_ctas < ctas) ctas = a->max_ctas;
  p.ksplit = 1;
  if (a->epi == 2) {
    int want = (2 * ctas + mn_tiles - 1) / mn_tiles;
    int maxsplit = (p.num_k_blocks + 3) / 4;
    p.ksplit = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
    if (p.ksplit < 1) p.ksplit = 1;
  }
  p.kb_per_split = (p.num_k_blocks + p.ksplit - 1) / p.ksplit;
  p.ksplit = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty slabs
  p.num_work = mn_tiles * p.ksplit;

  // ---- shared memory plan ----
  const int b_panels = (p.block_n + 63) / 64;
  p.b_bytes = p.b_mn ? b_panels * kPanelBytes64 : ((p.block_n * 128 + 1023) / 1024) * 1024;
  // an MN-major A of <= 64 channels has a single live panel: the second one (rows 64..127 of the
  // MMA, outputs that are never stored) may alias the B region, the stage shrinks by 8 KB
  p.a_bytes = (p.a_mn && a->M <= 64) ? kPanelBytes64 : kABytes;
  int stage = p.a_bytes + p.b_bytes;
  // wide transformed operands come in by TMA (128-byte box rows) and are rewritten in place, a
  // second source needs its own region; narrow ones are combined in registers on their way in
  p.a_tma = (p.a_xform != 0 && (p.a_mn ? a->M : a->K) >= 64) ? 1 : 0;
  p.b_tma = (p.b_xform != 0 && (p.b_mn ? a->N : a->K) >= 64) ? 1 : 0;
  const int dbg_env = env_dbg;
  if (dbg_env & 1024) p.a_tma = p.b_tma = 0;
  // transformed operands: 4 (2) loader warps share a stage so that its transform latency is short;
  // plain GEMMs are bound by load latency: one warp per stage, as many stages in flight as warps
  p.a2_off = p.b2_off = 0;
  if (p.a_tma && p.a_xform == 2) { p.a2_off = stage; stage += kABytes; }
  if (p.b_tma && p.b_xform == 2) { p.b2_off = stage; stage += p.b_bytes; }
  p.stage_bytes = stage;
  const bool xf = p.a_xform || p.b_xform;
  const int Ca = p.a_xform ? (p.a_mn ? p.M : p.K) : 0;
  const int Cb = p.b_xform ? (p.b_mn ? p.N : p.K) : 0;
  int fixed = 0;
  // every epilogue warp stages its own 32 x 64 sub-tile: 2 buffers each unless smem is short
  const int side_bytes = (a->epi == 1) ? kEpiWarps * kWarpOutBytes : 0;  // raw H rows (epi 1)
  const int coef_bytes = ((a->epi == 1 ? 4 * p.N : 0) + 3 * Ca + 3 * Cb) * 4;
  const int stats_bytes = (p.has_bnf || p.has_bnb) ? 2 * p.N * (int)sizeof(stat_t) : 0;
  const int budget = 232448 - 2048;  // 227 KB minus the kernel's static shared memory
  int stages = 0, out_bytes = 0;
  for (p.out_bufs = 2; p.out_bufs >= 1; --p.out_bufs) {
    out_bytes = (a->epi == 2) ? 0 : p.out_bufs * kEpiWarps * kWarpOutBytes;
    fixed = out_bytes + side_bytes + ((coef_bytes + 15) & ~15) + ((stats_bytes + 15) & ~15) +
            (int)sizeof(Bars) + 64;
    stages = (budget - fixed) / p.stage_bytes;
    // one 64-column sub-tile per tile: a warp's next staging use is two tile periods away (the
    // warpgroups alternate tiles), a single buffer never waits — spend the 32 KB on stages
    const bool single_sub = p.block_n <= 64 && a->epi != 2;
    if (single_sub && p.out_bufs == 2 && stages < kMaxStages) continue;
    if (stages >= 3 || p.out_bufs == 1) break;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(YAMB_EINVAL, "GEMM tile does not fit shared memory");
  p.num_stages = stages;
  // Throughput-bound launches (many k-blocks per CTA): one loader warp per stage, all warps on
  // different stages (measured on b3: 2 / 4 warps per stage are 10 / 40 % slower).  Latency-bound
  // launches (a handful of k-blocks per CTA: 7x7 / 14x14 layers, split-K tails): the spare warps
  // share a stage so that its load/transform latency, which then IS the kernel time, shrinks.
  {
    const int grid_est = p.num_work < ctas ? p.num_work : ctas;
    const long long iters = ((long long)(p.num_work + grid_est - 1) / grid_est) * p.kb_per_split;
    const int nlw = p.wg2x ? 8 : 4;
    p.lgroup = 1;
    while (p.lgroup < 4 && (long long)(nlw / (p.lgroup * 2)) >= iters) p.lgroup *= 2;
    // fewer stages than loader warps: the surplus warps would idle — let them share stages
    while (p.lgroup < 4 && nlw / (p.lgroup * 2) >= stages) p.lgroup *= 2;
  }
  if (dbg_env & 2048) p.lgroup = 4;
  if (dbg_env & 4096) p.lgroup = 2;
  if (dbg_env & 8192) p.lgroup = 1;
  int off = stages * p.stage_bytes;
  p.off_out = off; off += out_bytes;
  p.off_hside = off; off += side_bytes;
  p.off_coef = off; off += (coef_bytes + 15) & ~15;
  p.off_stats = off; off += (stats_bytes + 15) & ~15;
  p.off_bars = (off + 15) & ~15; off = p.off_bars + (int)sizeof(Bars);
  const int smem_total = off;  // the extern array is declared 1024-aligned: no slack needed

  // ---- tensor maps ----
  CUtensorMap tmA, tmB, tmA2, tmB2, tmD;
  memset(&tmA, 0, sizeof(tmA)); memset(&tmB, 0, sizeof(tmB));
  memset(&tmA2, 0, sizeof(tmA2)); memset(&tmB2, 0, sizeof(tmB2));
  memset(&tmD, 0, sizeof(tmD));
  int rc;
  if (p.a_tma) {
    if (!p.a_mn) rc = make_map_2d(&tmA, a->A, a->K, a->M, a->lda, 64, 128);
    else rc = make_map_2d(&tmA, a->A, a->M, a->K, a->lda, 64, 64);
    if (rc) return rc;
    if (p.a_xform == 2) {
      if (!p.a_mn) rc = make_map_2d(&tmA2, a->A2, a->K, a->M, a->lda2, 64, 128);
      else rc = make_map_2d(&tmA2, a->A2, a->M, a->K, a->lda2, 64, 64);
      if (rc) return rc;
    }
  }
  if (p.b_tma) {
    if (!p.b_mn) rc = make_map_2d(&tmB, a->B, a->K, a->N, a->ldb, 64, p.block_n);
    else rc = make_map_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, 64);
    if (rc) return rc;
    if (p.b_xform == 2) {
      if (!p.b_mn) rc = make_map_2d(&tmB2, a->B2, a->K, a->N, a->ldb2, 64, p.block_n);
      else rc = make_map_2d(&tmB2, a->B2, a->N, a->K, a->ldb2, 64, 64);
      if (rc) return rc;
    }
  }
  // operands are fetched with 16-byte cp.async: pointers 16-byte aligned, leading dims % 8 == 0
  {
    const void* ptrs[4] = {a->A, a->B, p.a_xform == 2 ? a->A2 : a->A, p.b_xform == 2 ? a->B2 : a->B};
    const long long lds[4] = {a->lda, a->ldb, p.a_xform == 2 ? a->lda2 : a->lda,
                              p.b_xform == 2 ? a->ldb2 : a->ldb};
    for (int i = 0; i < 4; ++i)
      if (!ptrs[i] || (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) || (lds[i] % 8) || lds[i] <= 0)
        return set_error(YAMB_EINVAL, "GEMM operand %d: pointer must be 16-byte aligned, ld %% 8 == 0", i);
  }
  p.gA = (const __nv_bfloat16*)a->A; p.lda = a->lda;
  p.gB = (const __nv_bfloat16*)a->B; p.ldb = a->ldb;
  p.gA2 = (const __nv_bfloat16*)(p.a_xform == 2 ? a->A2 : nullptr); p.lda2 = a->lda2;
  p.gB2 = (const __nv_bfloat16*)(p.b_xform == 2 ? a->B2 : nullptr); p.ldb2 = a->ldb2;
  if (a->epi != 2) {
    rc = make_map_2d(&tmD, a->D, a->N, a->M, a->ldd, 64, 32);  // one epilogue warp's rows
    if (rc) return rc;
  }
  if (a->epi == 1) { p.side = (const __nv_bfloat16*)a->H; p.lds = a->ldh; }
  else if (p.has_residual) { p.side = (const __nv_bfloat16*)a->residual; p.lds = a->ldr; }
  if (p.side && ((reinterpret_cast<uintptr_t>(p.side) & 15) || (p.lds % 8)))
    return set_error(YAMB_EINVAL, "side operand must be 16-byte aligned with lds % 8 == 0");

  const int grid = p.num_work < ctas ? p.num_work : ctas;
  cudaError_t e;
#define YAMB_GEMM_LAUNCH(XF, EP, THREADS)                                                         \
  do {                                                                                           \
    static int attr_smem = 0; /* per instantiation, process-wide: only ever RAISE the limit */     \
    if (attr_smem < smem_total) {                                                                \
      e = cudaFuncSetAttribute(gemm_tc_kernel<XF, EP>,                                           \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);         \
      if (e != cudaSuccess) return set_error(YAMB_ECUDA, "smem attr: %s", cudaGetErrorString(e)); \
      attr_smem = smem_total;                                                                    \
    }                                                                                            \
    gemm_tc_kernel<XF, EP><<<grid, THREADS, smem_total, stream>>>(tmA, tmB, tmA2, tmB2, tmD, p); \
  } while (0)
  if (xf) {
    if (a->epi == 0) YAMB_GEMM_LAUNCH(true, 0, 512);
    else if (a->epi == 1) YAMB_GEMM_LAUNCH(true, 1, 512);
    else YAMB_GEMM_LAUNCH(true, 2, 512);
  } else {
    if (a->epi == 0) YAMB_GEMM_LAUNCH(false, 0, 512);
    else if (a->epi == 1) YAMB_GEMM_LAUNCH(false, 1, 512);
    else YAMB_GEMM_LAUNCH(false, 2, 512);
  }
#undef YAMB_GEMM_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(YAMB_ECUDA, "gemm launch: %s", cudaGetErrorString(e));
  if (p.dbg & 512) {
    unsigned long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = h[7] ? (double)h[7] : 1.0;
    fprintf(stderr, "gemm dbg (cycles per epilogue warp): wait_full %.0f tmem_ld %.0f wait_store %.0f "
            "convert %.0f store %.0f stats %.0f total %.0f warps %.0f\n", h[0] / n, h[1] / n, h[2] / n,
            h[3] / n, h[4] / n, h[5] / n, h[6] / n, n);
    const double lw_n = h[9] ? (double)h[9] : 1.0;
    fprintf(stderr, "   per loader warp: wait_empty %.0f wait_tma %.0f xform(incl wait_tma) %.0f publish %.0f total %.0f "
            "(warps %.0f); per CTA: mma wait_full %.0f wait_tmem_empty %.0f; stages %d out_bufs %d block_n %d "
            "tma %d%d wg2x %d\n", h[12] / lw_n, h[13] / lw_n, h[14] / lw_n, h[15] / lw_n, h[8] / lw_n, lw_n,
            h[10] / (double)grid, h[11] / (double)grid, p.num_stages, p.out_bufs, p.block_n, p.a_tma, p.b_tma,
            p.wg2x);
  }
  return 0;
}

}  // namespace yamb
