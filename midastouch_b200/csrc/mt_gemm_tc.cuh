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
// shared-memory tile of 128 rows x TC_KT floats in the canonical K-major layout: 8-row x 16-byte core
// matrices; the K-adjacent core matrices of a row group sit TC_LBO bytes apart (128 + 16 bytes of padding:
// the staging stores of a warp then spread over all banks), row groups TC_SBO bytes apart
#define TC_LBO 144
#define TC_SBO ((TC_KT / 4) * TC_LBO)
#define TC_TILE_FLOATS (16 * TC_SBO / 4)

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (between the two core matrices of
//   one K=8 step) | [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// float offset of (row r, 16-byte K chunk kc) inside a tile
__device__ __forceinline__ int tc_tile_off_floats(int r, int kc) { return (r >> 3) * (TC_SBO / 4) + kc * (TC_LBO / 4) + (r & 7) * 4; }

__device__ __forceinline__ void tc_split4(const float4& x, float4& big, float4& small) {
  big.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), small.x = x.x - big.x;
  big.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u), small.y = x.y - big.y;
  big.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), small.z = x.z - big.z;
  big.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u), small.w = x.w - big.w;
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

// 4 consecutive K values of one row as float4 (zero beyond D / for absent rows)
template <typename T>
__device__ __forceinline__ float4 tc_load4(const T* __restrict__ row, int k, int D, bool ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!ok) return v;
  if (k + 3 < D) {
    if (sizeof(T) == 4) {
      v = __ldg(reinterpret_cast<const float4*>(row + k));
    } else {
      const double2 a = __ldg(reinterpret_cast<const double2*>(row + k)), b = __ldg(reinterpret_cast<const double2*>(row + k + 2));
      v = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    }
  } else {
    if (k < D) v.x = (float)row[k];
    if (k + 1 < D) v.y = (float)row[k + 1];
    if (k + 2 < D) v.z = (float)row[k + 2];
  }
  return v;
}

__device__ __forceinline__ void tc_mbar_wait(uint32_t mbar, uint32_t phase) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
  } while (!done);
}

// Thread mapping of the staging loads: warp w owns rows 32w..32w+31 of both tiles; lane l reads the
// 16-byte chunk kc = l & 7 of row 32w + 4i + (l >> 3), i = 0..7 -- eight lanes cover 128 contiguous
// bytes of a row, so every load instruction is four full lines.  The next stage's chunks are fetched
// into registers while the tensor core works on the current one.  D must be a multiple of 4 (float32
// rows: 16-byte alignment) -- checked by the host wrapper.
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
  __shared__ float s_qinv[TC_BN];

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
  const uint32_t aB = tc_smem_u32(sA_big), aS = tc_smem_u32(sA_small), bB = tc_smem_u32(sB_big), bS = tc_smem_u32(sB_small);

  const int kc = lane & 7, rsub = lane >> 3;
  const TE* arow[8];
  const float* brow[8];
  bool bok[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 32 * warp + 4 * i + rsub;
    arow[i] = E + (size_t)min(m0 + r, M - 1) * D;  // rows past the end: computed on a clamped row, never stored
    bok[i] = q0 + r < nq;
    brow[i] = Q + (size_t)(bok[i] ? q0 + r : 0) * D;
  }
  float4 ra[8], rb[8];
  float qn2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ra[i] = tc_load4(arow[i], 4 * kc, D, true);
    rb[i] = tc_load4(brow[i], 4 * kc, D, bok[i]);
    qn2[i] = 0.f;
  }
  uint32_t phase = 0;
  for (int k0 = 0; k0 < D; k0 += TC_KT) {
    if (k0) {  // the tensor core must be done with the previous stage before its operands are overwritten
      tc_mbar_wait(mbar, phase);
      phase ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int off = tc_tile_off_floats(32 * warp + 4 * i + rsub, kc);
      float4 big, small;
      tc_split4(ra[i], big, small);
      *reinterpret_cast<float4*>(sA_big + off) = big;
      *reinterpret_cast<float4*>(sA_small + off) = small;
      qn2[i] += rb[i].x * rb[i].x + rb[i].y * rb[i].y + rb[i].z * rb[i].z + rb[i].w * rb[i].w;
      tc_split4(rb[i], big, small);
      *reinterpret_cast<float4*>(sB_big + off) = big;
      *reinterpret_cast<float4*>(sB_small + off) = small;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int kk = 0; kk < TC_KT / 8; ++kk) {
        const uint32_t adv = kk * 2 * TC_LBO;  // two core matrices per K = 8 step
        const uint64_t dAb = tc_desc(aB + adv, TC_LBO, TC_SBO), dAs = tc_desc(aS + adv, TC_LBO, TC_SBO);
        const uint64_t dBb = tc_desc(bB + adv, TC_LBO, TC_SBO), dBs = tc_desc(bS + adv, TC_LBO, TC_SBO);
        tc_mma_tf32(tmem, dAb, dBb, idesc, (k0 | kk) ? 1u : 0u);
        tc_mma_tf32(tmem, dAb, dBs, idesc, 1u);
        tc_mma_tf32(tmem, dAs, dBb, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
    }
    if (k0 + TC_KT < D) {  // next stage's operands: in flight while the tensor core runs
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ra[i] = tc_load4(arow[i], k0 + TC_KT + 4 * kc, D, true);
        rb[i] = tc_load4(brow[i], k0 + TC_KT + 4 * kc, D, bok[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // |Q_q|: the eight lanes that share a row hold its partial sums
    float s = qn2[i];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (kc == 0) s_qinv[32 * warp + 4 * i + rsub] = 1.0f / fmaxf(sqrtf(s), 1e-8f);
  }
  tc_mbar_wait(mbar, phase);  // last stage done -> accumulator complete
  asm volatile("tcgen05.fence::after_thread_sync;");
  __syncthreads();
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= codebook rows), 32 columns (= queries) at a time
  const int m = m0 + 32 * warp + lane;
  const float rinv = (m < M) ? (float)(1.0 / rnorm[m]) : 0.f;
  for (int c0 = 0; c0 < TC_BN && q0 + c0 < nq; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
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
        if (q < nq) out[(size_t)q * M + m] = __uint_as_float(v[j]) * (rinv * s_qinv[c0 + j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_BN));
}
