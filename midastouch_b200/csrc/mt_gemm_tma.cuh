// Batched codebook query on the 5th-generation tensor cores, TMA-fed:
//     out[q][m] = cos(Q_q, E_m)      for nq tactile codes against all M codebook rows
// (the Q x D . (M x D)^T form of get_similarity: eval/single_touch_test.py:35-73 computes it in batches of 5000
// queries; the per-frame heat map of filter.py:213-215 is its nq = 1 case and stays on the HBM-bound k_codebook_query).
//
// Accuracy first: TF32 keeps 10 mantissa bits (~1e-3), the bar is 1e-5.  Every operand is split x = big + small
// (big = x with the low 13 mantissa bits cleared, exactly representable in TF32) and three products
// big*big + big*small + small*big are accumulated in the float32 TMEM accumulator ("3xTF32", ~1e-6 relative).
// The split planes of the codebook are built ONCE per upload (float32, also for a float64 codebook), the planes of the
// queries by a small kernel per call -- the GEMM kernel itself never converts or splits.
//
// Kernel: persistent, warp-specialised, one CTA per SM (192 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor.2d of the four operand tiles of a K step (128 rows x 32 floats each,
//             SWIZZLE_128B) into a 3-stage shared-memory ring, completion counted on the stage's "full" mbarrier
//   warp 1    MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 128, K = 8), twelve per
//             stage, from UMMA shared-memory descriptors (K-major, SWIZZLE_128B); tcgen05.commit releases the stage
//             ("empty" mbarrier) and, after the last K step, hands the accumulator to the epilogue ("tmem_full")
//   warps 2-5 epilogue: tcgen05.ld of the 128 x 128 float32 accumulator (TMEM is double-buffered: 2 x 128 columns, so the
//             MMA of the next tile overlaps the epilogue of this one), scaling by the cached row norms and the query
//             norms, 128-byte coalesced stores of out[q][m0 .. m0+127]
// Tile order: the query tiles of one row tile are consecutive, so the row tile's operand planes are read from HBM once
// and served from L2 to the other query tiles.
// Included by midas_b200.cu (same translation unit).
#pragma once
#include <cuda.h>

#define G2_BM 128
#define G2_BN 128
#define G2_BK 32  // floats per K step = 128 bytes = one SWIZZLE_128B atom row
#define G2_STAGES 3
#define G2_TILE_BYTES (128 * 128)
#define G2_STAGE_BYTES (4 * G2_TILE_BYTES)
#define G2_THREADS 192
#define G2_SMEM_BYTES (G2_STAGES * G2_STAGE_BYTES + 1024 + 256)

__device__ __forceinline__ uint32_t g2_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g2_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g2_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void g2_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g2_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g2_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void g2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 | leading byte
// offset (unused for a swizzled K-major tile one atom wide: 1) | stride byte offset = 8 rows x 128 B = 1024 >> 4 |
// version 1 | layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t g2_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void g2_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// x -> (big, small) planes, float32; optional per-row 1 / max(|row|, 1e-8) (queries)
template <typename T>
__global__ void __launch_bounds__(256) k_split_planes(const T* __restrict__ x, long long rows, int D, float* __restrict__ big,
                                                     float* __restrict__ small, float* __restrict__ rinv) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float nn = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = (float)x[row * D + k];
    const float b = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    big[row * D + k] = b;
    small[row * D + k] = v - b;
    nn = fmaf(v, v, nn);
  }
  if (rinv) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    if (lane == 0) rinv[row] = 1.0f / fmaxf(sqrtf(nn), 1e-8f);
  }
}

__global__ void __launch_bounds__(G2_THREADS, 1)
k_codebook_gemm_tma(const __grid_constant__ CUtensorMap tmA_big, const __grid_constant__ CUtensorMap tmA_small,
                    const __grid_constant__ CUtensorMap tmB_big, const __grid_constant__ CUtensorMap tmB_small,
                    const double* __restrict__ rnorm, const float* __restrict__ qinv, int M, int D, int nq, float* __restrict__ out) {
  extern __shared__ unsigned char g2_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)g2_smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte aligned tiles
  unsigned long long* bars = (unsigned long long*)(smem + G2_STAGES * G2_STAGE_BYTES);
  // bars[0..2] full, [3..5] empty, [6..7] tmem_full, [8..9] tmem_empty; then the TMEM base address
  uint32_t* tmem_slot = (uint32_t*)(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = g2_u32(smem), bar0 = g2_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (G2_STAGES + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * G2_STAGES + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * G2_STAGES + 2 + a); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < G2_STAGES; ++s) g2_mbar_init(full(s), 1), g2_mbar_init(empty(s), 1);
    for (int a = 0; a < 2; ++a) g2_mbar_init(tfull(a), 1), g2_mbar_init(tempty(a), 4);
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_big));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_small));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_big));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_small));
  }
  if (warp == 1) {  // TMEM: two float32 accumulators of 128 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g2_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;

  const int nqt = (nq + G2_BN - 1) / G2_BN, nmt = (M + G2_BM - 1) / G2_BM;
  const int ntiles = nqt * nmt, nk = (D + G2_BK - 1) / G2_BK;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile / nqt) * G2_BM, q0 = (tile % nqt) * G2_BN;
        for (int kb = 0; kb < nk; ++kb) {
          g2_mbar_wait(empty(stage), phase ^ 1);
          g2_mbar_expect_tx(full(stage), G2_STAGE_BYTES);
          const uint32_t dst = smem0 + stage * G2_STAGE_BYTES;
          g2_tma_load_2d(dst, &tmA_big, kb * G2_BK, m0, full(stage));
          g2_tma_load_2d(dst + G2_TILE_BYTES, &tmA_small, kb * G2_BK, m0, full(stage));
          g2_tma_load_2d(dst + 2 * G2_TILE_BYTES, &tmB_big, kb * G2_BK, q0, full(stage));
          g2_tma_load_2d(dst + 3 * G2_TILE_BYTES, &tmB_small, kb * G2_BK, q0, full(stage));
          if (++stage == G2_STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(G2_BN >> 3) << 17) | ((uint32_t)(G2_BM >> 4) << 24);
      int stage = 0, t = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        g2_mbar_wait(tempty(acc), acc_phase ^ 1);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t d = tmem + (uint32_t)(acc * G2_BN);
        for (int kb = 0; kb < nk; ++kb) {
          g2_mbar_wait(full(stage), phase);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t base = smem0 + stage * G2_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < G2_BK / 8; ++kk) {
            const uint32_t adv = kk * 32;  // 8 TF32 = 32 bytes along K inside the swizzle atom
            const uint64_t dAb = g2_desc(base + adv), dAs = g2_desc(base + G2_TILE_BYTES + adv);
            const uint64_t dBb = g2_desc(base + 2 * G2_TILE_BYTES + adv), dBs = g2_desc(base + 3 * G2_TILE_BYTES + adv);
            g2_mma_tf32(d, dAb, dBb, idesc, (kb | kk) ? 1u : 0u);
            g2_mma_tf32(d, dAb, dBs, idesc, 1u);
            g2_mma_tf32(d, dAs, dBb, idesc, 1u);
          }
          g2_commit(empty(stage));  // the stage is free once these MMAs have read it
          if (++stage == G2_STAGES) stage = 0, phase ^= 1;
        }
        g2_commit(tfull(acc));  // accumulator complete
      }
    }
  } else {  // ===== epilogue warps 2..5: TMEM lane quadrant = warp & 3
    const int quad = warp & 3;
    int t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      const int m0 = (tile / nqt) * G2_BM, q0 = (tile % nqt) * G2_BN;
      const int m = m0 + 32 * quad + lane;
      const float rinv = (m < M) ? (float)(1.0 / rnorm[m]) : 0.f;
      g2_mbar_wait(tfull(acc), acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int c0 = 0; c0 < G2_BN && q0 + c0 < nq; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(acc * G2_BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = q0 + c0 + j;
            if (q < nq) __stcs(out + (size_t)q * M + m, __uint_as_float(v[j]) * (rinv * __ldg(qinv + q)));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) g2_mbar_arrive(tempty(acc));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

// host: 2-D tensor map over a (rows, D) float32 plane, box = 32 floats x 128 rows, SWIZZLE_128B, zero fill out of bounds
typedef CUresult (*g2_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static g2_encode_fn g2_get_encode() {
  static g2_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (g2_encode_fn)p;
  }
  return fn;
}
static bool g2_make_map(CUtensorMap* map, const float* plane, long long rows, int D) {
  g2_encode_fn enc = g2_get_encode();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)D * sizeof(float)};
  const cuuint32_t box[2] = {G2_BK, G2_BM};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)plane, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
