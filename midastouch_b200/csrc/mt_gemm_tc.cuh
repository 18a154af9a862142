// Batched codebook query on the 5th-generation tensor cores (tcgen05 + TMEM, sm_100a):
//     out[q][m] = cos(Q_q, E_m)      for nq tactile codes against all M codebook rows
// (the Q x D . (M x D)^T form of get_similarity: eval/single_touch_test.py:35-73 computes it in
// batches of 5000 queries; live_demo.py:107-109 and the heat map of filter.py:213-215 are its nq = 1
// case, which stays on the HBM-bound k_codebook_query).
//
// One CTA (128 threads) owns a 128 (codebook rows) x 128 (queries) output tile whose float32
// accumulator lives in TMEM (128 lanes x 128 columns).  Per K-step of 32 the threads stage the two
// operand tiles in shared memory in the canonical K-major / no-swizzle UMMA layout (8-row x 16-byte
// core matrices), one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8), and
// tcgen05.commit signals an mbarrier when the tensor core has consumed the stage.
// Accuracy: TF32 keeps 10 mantissa bits, which would cost ~1e-3; every operand is therefore split
// x = big + small (big = x with the low 13 mantissa bits cleared) and three products
// big*big + big*small + small*big are accumulated ("3xTF32"), giving ~1e-6 relative -- inside the
// 1e-5 bar of the float64 reference.  The epilogue reads the accumulator back with tcgen05.ld and
// divides by the cached row norms and the query norms.
//
// Included by midas_b200.cu (same translation unit).
#pragma once

#define TC_BM 128
#define TC_BN 128
#define TC_KT 32
#define TC_TILE_FLOATS (128 * TC_KT)

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (between the two core matrices of
//   one K=8 step) | [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// element (row r, k) of a 128 x TC_KT tile: 8 core matrices along K per 8-row group
__device__ __forceinline__ int tc_tile_off_floats(int r, int kc) { return (r >> 3) * (TC_KT / 4) * 32 + kc * 32 + (r & 7) * 4; }

__device__ __forceinline__ void tc_split(float x, float& big, float& small) {
  big = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  small = x - big;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

template <typename TE>
__global__ void __launch_bounds__(128) k_codebook_gemm_tc(const TE* __restrict__ E, const double* __restrict__ rnorm, int M, int D,
                                                          const float* __restrict__ Q, int nq, float* __restrict__ out) {
  extern __shared__ __align__(128) float tc_smem[];
  float* sA_big = tc_smem;
  float* sA_small = sA_big + TC_TILE_FLOATS;
  float* sB_big = sA_small + TC_TILE_FLOATS;
  float* sB_small = sB_big + TC_TILE_FLOATS;
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_qn[TC_BN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM, q0 = blockIdx.y * TC_BN;
  const uint32_t mbar = tc_smem_u32(&s_mbar);

  if (warp == 0) {  // TMEM: 128 columns of float32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"((uint32_t)TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint32_t lbo = 128, sbo = (TC_KT / 4) * 128;
  const uint32_t aB = tc_smem_u32(sA_big), aS = tc_smem_u32(sA_small), bB = tc_smem_u32(sB_big), bS = tc_smem_u32(sB_small);

  const int arow = min(m0 + tid, M - 1);  // rows past the end are computed on a clamped row and never stored
  const TE* __restrict__ erow = E + (size_t)arow * D;
  const bool qok = q0 + tid < nq;
  const float* __restrict__ qrow = Q + (size_t)(qok ? q0 + tid : 0) * D;
  float qn2 = 0.f;
  uint32_t phase = 0;
  for (int k0 = 0; k0 < D; k0 += TC_KT) {
    if (k0) {  // the tensor core must be done with the previous stage before its operands are overwritten
      uint32_t done;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
      } while (!done);
      phase ^= 1;
    }
#pragma unroll
    for (int kc = 0; kc < TC_KT / 4; ++kc) {
      const int k = k0 + 4 * kc;
      float a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = (k + j < D) ? (float)erow[k + j] : 0.f;
        b[j] = (qok && k + j < D) ? qrow[k + j] : 0.f;
        qn2 = fmaf(b[j], b[j], qn2);
      }
      float4 ab, as, bb, bs;
      tc_split(a[0], ab.x, as.x), tc_split(a[1], ab.y, as.y), tc_split(a[2], ab.z, as.z), tc_split(a[3], ab.w, as.w);
      tc_split(b[0], bb.x, bs.x), tc_split(b[1], bb.y, bs.y), tc_split(b[2], bb.z, bs.z), tc_split(b[3], bb.w, bs.w);
      const int off = tc_tile_off_floats(tid, kc);
      *reinterpret_cast<float4*>(sA_big + off) = ab;
      *reinterpret_cast<float4*>(sA_small + off) = as;
      *reinterpret_cast<float4*>(sB_big + off) = bb;
      *reinterpret_cast<float4*>(sB_small + off) = bs;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int kk = 0; kk < TC_KT / 8; ++kk) {
        const uint32_t adv = kk * 2 * lbo;  // two core matrices per K = 8 step
        const uint64_t dAb = tc_desc(aB + adv, lbo, sbo), dAs = tc_desc(aS + adv, lbo, sbo);
        const uint64_t dBb = tc_desc(bB + adv, lbo, sbo), dBs = tc_desc(bS + adv, lbo, sbo);
        tc_mma_tf32(tmem, dAb, dBb, idesc, (k0 | kk) ? 1u : 0u);
        tc_mma_tf32(tmem, dAb, dBs, idesc, 1u);
        tc_mma_tf32(tmem, dAs, dBb, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
    }
  }
  s_qn[tid] = fmaxf(sqrtf(qn2), 1e-8f);
  {  // last stage done -> accumulator complete
    uint32_t done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
    } while (!done);
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  __syncthreads();
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= codebook rows), 8 columns (= queries) at a time
  const int m = m0 + 32 * warp + lane;
  const float rn = (m < M) ? (float)rnorm[m] : 1.f;
  for (int c0 = 0; c0 < TC_BN; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = q0 + c0 + j;
      if (q < nq && m < M) out[(size_t)q * M + m] = __uint_as_float(v[j]) / (rn * s_qn[c0 + j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_BN));
}
