// Per-particle arithmetic shared by every kernel of libmidas_b200 (sm_100a).
//
// All functions are MT_HD (host+device) so the same source is unit-tested on the build
// box by compiling tests/host_math_harness.cpp with g++ -- the CUDA library itself has
// no host compute path.  Citations are file:line under the reference tree.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MT_HD __host__ __device__ __forceinline__
#else
#define MT_HD inline
#endif

// float ops that must not be contracted into FMAs (bit-stable distance ordering).
#if defined(__CUDA_ARCH__)
#define MT_FSUB(a, b) __fsub_rn((a), (b))
#define MT_FMUL(a, b) __fmul_rn((a), (b))
#define MT_FADD(a, b) __fadd_rn((a), (b))
#else
static inline float mt_nofma_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float mt_nofma_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float mt_nofma_add(float a, float b) { volatile float r = a + b; return r; }
#define MT_FSUB(a, b) mt_nofma_sub((a), (b))
#define MT_FMUL(a, b) mt_nofma_mul((a), (b))
#define MT_FADD(a, b) mt_nofma_add((a), (b))
#endif

// ---------------------------------------------------------------- L2 residency hints
// One filter step streams ~0.5 GB (poses read + written twice, the embeddings once) through a 126 MB L2, while the
// tables every particle consults (codebook keys, neighbour lists, drift-test voxels and vertices, weight tables)
// total a few tens of MB of hot lines.  Without hints the stream evicts the tables and every dependent table
// access of the next kernel pays DRAM latency (measured: 45 % of the L2 requests of k_step_a missed).  Table loads
// therefore carry an L2 evict_last policy, the streams evict_first.
#ifndef MT_L2_HINTS
#define MT_L2_HINTS 1
#endif
#if defined(__CUDACC__)
// MT_POL_CONST: the two policies as immediate descriptors instead of a createpolicy instruction per use (what
// createpolicy.fractional returns for fraction 1.0; the same 64-bit encodings CUTLASS passes as TMA cache hints:
// cute/arch/copy_sm90_desc.hpp, CacheHintSm90::EVICT_FIRST / EVICT_LAST)
#ifndef MT_POL_CONST
#define MT_POL_CONST 1
#endif
__device__ __forceinline__ unsigned long long mt_pol_keep() {
#if MT_POL_CONST
  return 0x14F0000000000000ull;
#else
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
#endif
}
__device__ __forceinline__ unsigned long long mt_pol_stream() {
#if MT_POL_CONST
  return 0x12F0000000000000ull;
#else
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
#endif
}
// single-instruction square root (MUFU.SQRT, relative error <= 2^-22, subnormal inputs flush to zero): only for
// quantities that enter inflated bounds or drawn noise, never for a compared distance or a key
__device__ __forceinline__ float mt_sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#if MT_L2_HINTS
// table loads (read-only data path, L2 evict_last)
__device__ __forceinline__ float4 mt_ldk(const float4* a) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(mt_pol_keep()));
  return v;
}
__device__ __forceinline__ int mt_ldk(const int* a) {
  int v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(mt_pol_keep()));
  return v;
}
__device__ __forceinline__ unsigned mt_ldk(const unsigned* a) {
  unsigned v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(mt_pol_keep()));
  return v;
}
__device__ __forceinline__ float mt_ldk(const float* a) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(mt_pol_keep()));
  return v;
}
__device__ __forceinline__ double mt_ldk(const double* a) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(mt_pol_keep()));
  return v;
}
__device__ __forceinline__ void mt_prefetch_keep(const void* a) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a)); }
// streamed particle arrays (L2 evict_first)
__device__ __forceinline__ float4 mt_lds(const float4* a) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(mt_pol_stream()));
  return v;
}
__device__ __forceinline__ int mt_lds(const int* a) {
  int v;
  asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(mt_pol_stream()));
  return v;
}
__device__ __forceinline__ void mt_sts(float4* a, float4 v) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(mt_pol_stream()) : "memory");
}
__device__ __forceinline__ void mt_sts(int* a, int v) {
  asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(mt_pol_stream()) : "memory");
}
#else
template <typename T>
__device__ __forceinline__ T mt_ldk(const T* a) { return __ldg(a); }
__device__ __forceinline__ void mt_prefetch_keep(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }
template <typename T>
__device__ __forceinline__ T mt_lds(const T* a) { return *a; }
template <typename T>
__device__ __forceinline__ void mt_sts(T* a, T v) { *a = v; }
#endif
#endif

struct mt_pose {  // rows of the 3x4 [R|t]; the constant bottom row (0,0,0,1) is implicit
  float r[3][4];
};

// odometry pre-multiplied form: G = odom (3x4 affine)
struct mt_affine {
  float m[3][4];
};

// ---------------------------------------------------------------- SE(3) helpers
// C = A * B for 3x4 affines with implicit bottom row; same association as
// torch.matmul on (4,4) operands (particle_filter.py:345,374): each entry is a 4-term
// dot product whose products with the constant row are exact.
MT_HD void mt_compose(const float A[3][4], const float B[3][4], float C[3][4]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = A[i][0] * B[0][j];
      s = fmaf(A[i][1], B[1][j], s);
      s = fmaf(A[i][2], B[2][j], s);
      if (j == 3) s += A[i][3];
      C[i][j] = s;
    }
  }
}

// Rn = Rz(a0) Ry(a1) Rx(a2), angles = deg2rad(rot_deg) in float32
// (euler_angles_to_matrix(.., "ZYX"), pose.py:215-269; deg2rad at particle_filter.py:336).
// fast != 0 (device, in-kernel noise only): SFU sine / cosine -- the angles are drawn noise of a
// fraction of a degree, where __sincosf is accurate to ~1e-7 absolute.
MT_HD void mt_noise_affine(const float tn[3], const float rot_deg[3], float Tn[3][4], int fast = 0) {
  const float d2r = 0.017453292519943295f;  // torch.deg2rad: x * (pi/180) in float32
  float a0 = rot_deg[0] * d2r, a1 = rot_deg[1] * d2r, a2 = rot_deg[2] * d2r;
  float cz, sz, cy, sy, cx, sx;
#if defined(__CUDA_ARCH__)
  if (fast) {
    __sincosf(a0, &sz, &cz);
    __sincosf(a1, &sy, &cy);
    __sincosf(a2, &sx, &cx);
  } else {
    sincosf(a0, &sz, &cz);
    sincosf(a1, &sy, &cy);
    sincosf(a2, &sx, &cx);
  }
#else
  (void)fast;
  sz = sinf(a0); cz = cosf(a0); sy = sinf(a1); cy = cosf(a1); sx = sinf(a2); cx = cosf(a2);
#endif
  // (Rz Ry) first, then times Rx -- matrices[0] @ matrices[1] @ matrices[2]
  float zy[3][3] = {{cz * cy, -sz, cz * sy}, {sz * cy, cz, sz * sy}, {-sy, 0.f, cy}};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Tn[i][0] = zy[i][0];
    Tn[i][1] = zy[i][1] * cx + zy[i][2] * sx;
    Tn[i][2] = zy[i][2] * cx - zy[i][1] * sx;
    Tn[i][3] = tn[i];
  }
}

// scipy Rotation.from_euler("zyx", rot, degrees=True) -- lowercase = extrinsic axes, i.e.
// Rn = Rx(a2) Ry(a1) Rz(a0) -- as used by init_filter (particle_filter.py:139-141).
MT_HD void mt_noise_affine_extrinsic(const float tn[3], const float rot_deg[3], float Tn[3][4]) {
  const float d2r = 0.017453292519943295f;
  float a0 = rot_deg[0] * d2r, a1 = rot_deg[1] * d2r, a2 = rot_deg[2] * d2r;
  float sz = sinf(a0), cz = cosf(a0), sy = sinf(a1), cy = cosf(a1), sx = sinf(a2), cx = cosf(a2);
  float xy[3][3] = {{cy, 0.f, sy}, {sx * sy, cx, -sx * cy}, {-cx * sy, sx, cx * cy}};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Tn[i][0] = xy[i][0] * cz + xy[i][1] * sz;
    Tn[i][1] = xy[i][1] * cz - xy[i][0] * sz;
    Tn[i][2] = xy[i][2];
    Tn[i][3] = tn[i];
  }
}

// theseus SO3.log_map restated in float32 (call site pose.py:19-23); thresholds are the
// float32 entries of theseus/constants.py (near-zero 5e-3, near-pi 1e-2).
MT_HD void mt_so3_log(const float R[3][4], float out[3]) {
  float sa0 = 0.5f * (R[2][1] - R[1][2]);
  float sa1 = 0.5f * (R[0][2] - R[2][0]);
  float sa2 = 0.5f * (R[1][0] - R[0][1]);
  float cosine = 0.5f * ((R[0][0] + R[1][1] + R[2][2]) - 1.f);
  float sine = sqrtf(sa0 * sa0 + sa1 * sa1 + sa2 * sa2);
  float theta = atan2f(sine, cosine);
  bool near_zero = theta < 5e-3f;
  bool near_pi = (1.f + cosine) <= 1e-2f;
  if (!near_pi) {
    float scale = near_zero ? (1.f + sine * sine / 6.f) : (theta / sine);
    out[0] = sa0 * scale; out[1] = sa1 * scale; out[2] = sa2 * scale;
    return;
  }
  float d0 = R[0][0], d1 = R[1][1], d2 = R[2][2];
  int major = ((d1 > d0) && (d1 > d2) ? 1 : 0) + 2 * ((d2 > d0) && (d2 > d1) ? 1 : 0);
  float sel[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = (major == 0 ? R[0][k] : (major == 1 ? R[1][k] : R[2][k]));  // row `major`
    float b = (major == 0 ? R[k][0] : (major == 1 ? R[k][1] : R[k][2]));  // column `major`
    sel[k] = 0.5f * (a + b);
  }
  if (major == 0) sel[0] -= cosine; else if (major == 1) sel[1] -= cosine; else sel[2] -= cosine;
  float nrm = sqrtf(sel[0] * sel[0] + sel[1] * sel[1] + sel[2] * sel[2]);
  float sgn_src = (major == 0 ? sa0 : (major == 1 ? sa1 : sa2));
  float sign = (sgn_src < 0.f) ? -1.f : 1.f;  // sign(0) -> +1
  float k = theta * sign;
  out[0] = sel[0] / nrm * k; out[1] = sel[1] / nrm * k; out[2] = sel[2] / nrm * k;
}

// R3_SE3 (tactile_tree.py:73-77): key = [(1-w) t, w Log(R)], w = 0.01
MT_HD void mt_se3_key(const float P[3][4], float key[6]) {
  const float w = 0.01f;
  float lg[3];
  mt_so3_log(P, lg);
  key[0] = (1.0f - w) * P[0][3]; key[1] = (1.0f - w) * P[1][3]; key[2] = (1.0f - w) * P[2][3];
  key[3] = w * lg[0]; key[4] = w * lg[1]; key[5] = w * lg[2];
}

// squared L2 over the 6-D key with a fixed evaluation order (matches oracle.l2_sq_f32) so that argmin
// ties are decided identically everywhere.  Two interleaved fused chains,
//   lo = fma(d4, d4, fma(d2, d2, d0 * d0)),  hi = fma(d5, d5, fma(d3, d3, d1 * d1)),  result = lo + hi,
// d_k = a_k - b_k, every operation correctly rounded in float32: on sm_100a this is three packed
// subtractions + one packed multiply + two packed FMAs (FADD2 / FMUL2 / FFMA2, two float32 per issue slot)
// + one add -- 7 instructions instead of 17 for the unfused sum.  The host side uses fmaf().
#if defined(__CUDA_ARCH__) && !defined(MT_DIST_SCALAR)
__device__ __forceinline__ unsigned long long mt_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float mt_key_dist(const float a[6], const float b[6]) {
  unsigned long long d01, d23, d45, acc;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(mt_pack2(a[0], a[1])), "l"(mt_pack2(b[0], b[1])));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(mt_pack2(a[2], a[3])), "l"(mt_pack2(b[2], b[3])));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d45) : "l"(mt_pack2(a[4], a[5])), "l"(mt_pack2(b[4], b[5])));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(acc) : "l"(d01));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(acc) : "l"(d23), "l"(acc));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(acc) : "l"(d45), "l"(acc));
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
  return __fadd_rn(lo, hi);
}
#else
MT_HD float mt_key_dist(const float a[6], const float b[6]) {
  const float d0 = MT_FSUB(a[0], b[0]), d1 = MT_FSUB(a[1], b[1]), d2 = MT_FSUB(a[2], b[2]);
  const float d3 = MT_FSUB(a[3], b[3]), d4 = MT_FSUB(a[4], b[4]), d5 = MT_FSUB(a[5], b[5]);
  const float lo = fmaf(d4, d4, fmaf(d2, d2, MT_FMUL(d0, d0)));
  const float hi = fmaf(d5, d5, fmaf(d3, d3, MT_FMUL(d1, d1)));
  return MT_FADD(lo, hi);
}
#endif

// check_quats (particle_filter.py:347-357): a pose is pruned when its quaternion norm is 0
// or NaN.  For finite inputs theseus' to_quaternion never has zero norm (w = 0.5 sqrt(1+tr)
// and the near-pi branch returns a unit axis), so the test reduces to non-finite entries.
MT_HD bool mt_pose_invalid(const float P[3][4]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) s += P[i][j] * 0.f;  // NaN/Inf -> NaN
  return !(s == 0.f);
}

// rot2euler of R_gt R_n^T (particle_filter.py:484, pose.py:201-208) + nan_to_num + wrap
MT_HD float mt_rot_err_deg(const float G[3][4], const float P[3][4]) {
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) tr += G[i][0] * P[i][0] + G[i][1] * P[i][1] + G[i][2] * P[i][2];
  float a = acosf((tr - 1.0f) * 0.5f) * 57.29577951308232f;
  if (a != a) a = 0.f;
  if (a > 180.f) a -= 360.f;
  if (a < -180.f) a += 360.f;
  return a;
}

// ---------------------------------------------------------------- systematic draw
// sample location of slot k exactly as particle_filter.py:254-261: fl64(k/N) + off with
// off = float32(u)/N promoted to float64, then remainder 1.
MT_HD double mt_loc(long long k, double dN, double off) {
  double l = (double)k / dN + off;
  if (l >= 1.0) l -= 1.0;  // torch.remainder(x, 1) for x in [1, 2)
  return l;
}

// fl(k / dN), correctly rounded, without a division in the loop: with rN = fl(1 / dN), q = fl(k rN),
// r = k - q dN (exact in one fma), fl(q + r rN) is the correctly rounded quotient (Markstein's correction step; it
// needs rN within half an ulp of 1/dN, which the IEEE division gives, and fails only for divisors whose 53-bit
// significand is all ones -- not an integer particle count).  Checked against the division itself in
// tests/test_host_math.py.
MT_HD double mt_div_rn(double k, double dN, double rN) {
  const double q = k * rN;
  const double r = fma(-q, dN, k);
  return fma(r, rN, q);
}

// loc with the two-pointer semantics of the reference loop (particle_filter.py:295-302):
// if the last location wraps to ~0 it is assigned to whichever parent owns slot N-2, i.e.
// it behaves like loc(N-2) for counting.
MT_HD double mt_loc_mono(long long k, long long N, double dN, double rN, double off) {
  double l = mt_div_rn((double)k, dN, rN) + off;
  if (l >= 1.0) l = (k > 0) ? (mt_div_rn((double)(k - 1), dN, rN) + off) : (l - 1.0);  // only k = N-1 can wrap
  (void)N;
  return l;
}

// cnt(C) = #{k in [0,N) : loc_k < C}; locs are non-decreasing so this is the first k
// whose loc is >= C.  A closed-form guess is corrected with exact comparisons.
MT_HD long long mt_count_below(double C, long long N, double dN, double off) {
  if (!(C == C)) return 0;  // NaN CDF value owns nothing
  const double rN = 1.0 / dN;
  double g = ceil((C - off) * dN);
  long long k = (g <= 0.0) ? 0 : (g >= dN ? N : (long long)g);
  while (k > 0 && !(mt_loc_mono(k - 1, N, dN, rN, off) < C)) --k;
  while (k < N && (mt_loc_mono(k, N, dN, rN, off) < C)) ++k;
  return k;
}

// ---------------------------------------------------------------- Philox4x32-10
struct mt_u4 {
  uint32_t x, y, z, w;
};

MT_HD void mt_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

MT_HD mt_u4 mt_philox(mt_u4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mt_mulhilo(0xD2511F53u, c.x, &h0, &l0);
    mt_mulhilo(0xCD9E8D57u, c.z, &h1, &l1);
    mt_u4 n = {h1 ^ c.y ^ k0, l1, h0 ^ c.w ^ k1, l0};
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// 53-bit uniform in [0,1) from two 32-bit words
MT_HD double mt_u01_53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)a >> 5) << 26) | ((uint64_t)b >> 6)) * (1.0 / 9007199254740992.0);
}
// first index i in [0,n) with C[i] > t (n when there is none); C non-decreasing
MT_HD long long mt_upper_bound(const double* C, long long n, double t) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (C[mid] > t) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
// categorical draw from an un-normalised inclusive CDF (C[n-1] <= S): u in [0,1) -> first i with C[i] > u*S.
// Items of zero weight (C[i] == C[i-1]) are never returned.
MT_HD long long mt_cdf_draw(const double* C, long long n, double S, double u) {
  double t = u * S;
  long long r = mt_upper_bound(C, n, t);
  if (r >= n) {  // u*S rounded up to (or past) the last CDF value: the last item of positive weight
    const double top = C[n - 1];
    long long lo = 0, hi = n - 1;  // first index that attains the maximum
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (C[mid] >= top) hi = mid;
      else lo = mid + 1;
    }
    r = lo;
  }
  return r;
}

// 21-bit uniform in (0,1): six of them come out of ONE Philox4x32-10 call (128 bits)
MT_HD float mt_u01_21(uint32_t x) { return ((float)(x & 0x1FFFFFu) + 0.5f) * (1.0f / 2097152.0f); }

// six N(0,1) draws for particle `gid` of filter step `step`: (tn[3], rot[3]).  One Philox call keyed by
// (seed; gid, step); the 128 output bits are cut into six 21-bit uniforms (Box-Muller tails reach
// 5.4 sigma, far beyond what a 1e6-particle cloud resolves).
MT_HD void mt_motion_normals(uint64_t seed, uint64_t step, uint64_t gid, float tn[3], float rot[3]) {
  mt_u4 c0 = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
  mt_u4 a = mt_philox(c0, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t lo = ((uint64_t)a.y << 32) | a.x, hi = ((uint64_t)a.w << 32) | a.z;
  const uint32_t u0 = (uint32_t)lo, u1 = (uint32_t)(lo >> 21), u2 = (uint32_t)(lo >> 42);
  const uint32_t u3 = (uint32_t)hi, u4 = (uint32_t)(hi >> 21), u5 = (uint32_t)(hi >> 42);
  float n[6];
  const uint32_t ua[3] = {u0, u2, u4}, ub[3] = {u1, u3, u5};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#if defined(__CUDA_ARCH__)
    const float r = mt_sqrt_fast(-2.0f * __logf(mt_u01_21(ua[k])));  // argument in [2.4e-7, 30]
    float s, c;
    __sincosf(6.283185307179586f * mt_u01_21(ub[k]), &s, &c);
#else
    const float r = sqrtf(-2.0f * logf(mt_u01_21(ua[k])));
    const float ang = 6.283185307179586f * mt_u01_21(ub[k]);
    const float s = sinf(ang), c = cosf(ang);
#endif
    n[2 * k] = r * c, n[2 * k + 1] = r * s;
  }
  tn[0] = n[0], tn[1] = n[1], tn[2] = n[2], rot[0] = n[3], rot[1] = n[4], rot[2] = n[5];
}
