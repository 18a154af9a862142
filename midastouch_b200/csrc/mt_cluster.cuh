// Cluster centres and particle-count annealing.
//
//  get_cluster_centers(method="quat_avg")  particle_filter.py:153-206, xyz_quat_averaged pose.py:112-147
//      per cluster label: weighted mean translation, Markley quaternion average (dominant
//      eigenvector of sum w q q^T, antipodal-canonicalised) and weighted translation std.
//      Two passes over the particles (min/max of the float32 weights decides the
//      "constant weights -> uniform" rule, 178-184), float64 block partials combined in a fixed
//      order, 4x4 Jacobi eigen-solver in the finalise kernel (the reference's Tensor.eig was
//      removed from torch; the matrix is symmetric).
//  annealing  particle_filter.py:405-447
//      "drop the k lowest-weight particles" / "duplicate the k highest" = k-th order statistic
//      by an 8-pass radix select on the order-preserving bit pattern of the float64 weights +
//      a stable compaction (ties at the threshold: lowest index first).
//
// Included by midas_b200.cu (same translation unit).
#pragma once

#define MT_MAX_CLUSTERS 16
#define MT_CL_VALS 18  // 10 moment entries + 3 w*t + 3 w*t^2 + w + count

// theseus SO3.to_quaternion restated (call site pose.py:26-34) -> (w, x, y, z)
MT_HD void mt_so3_to_quat(const float R[3][4], float q[4]) {
  const float sa0 = 0.5f * (R[2][1] - R[1][2]), sa1 = 0.5f * (R[0][2] - R[2][0]), sa2 = 0.5f * (R[1][0] - R[0][1]);
  const float tr = R[0][0] + R[1][1] + R[2][2];
  const float w = 0.5f * sqrtf(fminf(fmaxf(1.f + tr, 0.f), 4.f));
  if (!(w <= 1e-2f)) {
    const float s = 0.5f / w;
    q[0] = w, q[1] = sa0 * s, q[2] = sa1 * s, q[3] = sa2 * s;
    return;
  }
  const float cosine = 0.5f * (tr - 1.f);
  const float d0 = R[0][0], d1 = R[1][1], d2 = R[2][2];
  const int major = ((d1 > d0) && (d1 > d2) ? 1 : 0) + 2 * ((d2 > d0) && (d2 > d1) ? 1 : 0);
  float sel[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float a = (major == 0 ? R[0][k] : (major == 1 ? R[1][k] : R[2][k]));
    const float b = (major == 0 ? R[k][0] : (major == 1 ? R[k][1] : R[k][2]));
    sel[k] = 0.5f * (a + b);
  }
  if (major == 0) sel[0] -= cosine; else if (major == 1) sel[1] -= cosine; else sel[2] -= cosine;
  float nrm = sqrtf(sel[0] * sel[0] + sel[1] * sel[1] + sel[2] * sel[2]);
  if (nrm == 0.f) nrm = 1.f;
  const float sg = ((major == 0 ? sa0 : (major == 1 ? sa1 : sa2)) < 0.f) ? -1.f : 1.f;
  const float sh = sqrtf(fminf(fmaxf(1.f - w * w, 0.f), 1.f)) * sg;
  q[0] = w, q[1] = sel[0] / nrm * sh, q[2] = sel[1] / nrm * sh, q[3] = sel[2] / nrm * sh;
}

// theseus SE3.log_map / SE3.exp_map restated (call sites pose.py:101-109, log_map_averaged): tangent = [v, omega] with
// omega = Log_SO3(R) (mt_so3_log) and v = V^-1 t,  V^-1 = I - 1/2 [w]x + a [w]x^2,
// a = (1 - theta sin(theta) / (2 (1 - cos(theta)))) / theta^2  (-> 1/12 for theta -> 0).  float32 like the reference.
MT_HD void mt_se3_log(const float P[3][4], float out[6]) {
  float w[3];
  mt_so3_log(P, w);
  const float th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const float th = sqrtf(th2);
  float a;
  if (th < 5e-3f) a = 1.f / 12.f + th2 / 720.f;
  else a = (1.f - 0.5f * th * sinf(th) / (1.f - cosf(th))) / th2;
  const float t[3] = {P[0][3], P[1][3], P[2][3]};
  const float c1[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};      // w x t
  const float c2[3] = {w[1] * c1[2] - w[2] * c1[1], w[2] * c1[0] - w[0] * c1[2], w[0] * c1[1] - w[1] * c1[0]};  // w x (w x t)
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = t[k] - 0.5f * c1[k] + a * c2[k], out[3 + k] = w[k];
}
// exp of a tangent [v, omega] in float64 -> 3x4 [R|t]:  R = I + A [w]x + B [w]x^2,  t = (I + B [w]x + C [w]x^2) v,
// A = sin(th)/th, B = (1 - cos(th))/th^2, C = (th - sin(th))/th^3
MT_HD void mt_se3_exp(const double x[6], double T[3][4]) {
  const double v[3] = {x[0], x[1], x[2]}, w[3] = {x[3], x[4], x[5]};
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B, C;
  if (th < 1e-6) A = 1.0 - th2 / 6.0, B = 0.5 - th2 / 24.0, C = 1.0 / 6.0 - th2 / 120.0;
  else A = sin(th) / th, B = (1.0 - cos(th)) / th2, C = (th - sin(th)) / (th2 * th);
  const double K[3][3] = {{0.0, -w[2], w[1]}, {w[2], 0.0, -w[0]}, {-w[1], w[0], 0.0}};
  double K2[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K2[i][j] = K[i][0] * K[0][j] + K[i][1] * K[1][j] + K[i][2] * K[2][j];
  for (int i = 0; i < 3; ++i) {
    double t = 0.0;
    for (int j = 0; j < 3; ++j) {
      T[i][j] = (i == j ? 1.0 : 0.0) + A * K[i][j] + B * K2[i][j];
      t += ((i == j ? 1.0 : 0.0) + B * K[i][j] + C * K2[i][j]) * v[j];
    }
    T[i][3] = t;
  }
}

#if defined(__CUDACC__)
// pass 1: per-cluster min / max of the float32-cast weights (block partials)
__global__ void __launch_bounds__(256) k_cluster_minmax(const double* __restrict__ w, const int* __restrict__ label, long long n, int K,
                                                        float* __restrict__ part /* nblocks x K x 2 */) {
  __shared__ float s_mn[8], s_mx[8];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const float wi = (i < n) ? (float)w[i] : 0.f;
  const int li = (i < n) ? label[i] : -1;
  for (int k = 0; k < K; ++k) {
    float mn = (li == k) ? wi : FLT_MAX, mx = (li == k) ? wi : -FLT_MAX;
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_mn[threadIdx.x >> 5] = mn, s_mx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int j = 1; j < 8; ++j) mn = fminf(mn, s_mn[j]), mx = fmaxf(mx, s_mx[j]);
      part[((size_t)blockIdx.x * K + k) * 2] = mn;
      part[((size_t)blockIdx.x * K + k) * 2 + 1] = mx;
    }
  }
}
__global__ void __launch_bounds__(256) k_cluster_minmax_final(const float* __restrict__ part, int nblocks, int K, int* __restrict__ uniform) {
  __shared__ float s_mn[8], s_mx[8];
  const int k = blockIdx.x;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int b = threadIdx.x; b < nblocks; b += 256) mn = fminf(mn, part[((size_t)b * K + k) * 2]), mx = fmaxf(mx, part[((size_t)b * K + k) * 2 + 1]);
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) s_mn[threadIdx.x >> 5] = mn, s_mx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < 8; ++j) mn = fminf(mn, s_mn[j]), mx = fmaxf(mx, s_mx[j]);
    // torch.isclose(max - min, 0): |d| <= 1e-8 (float32 arithmetic, particle_filter.py:178-184)
    uniform[k] = fabsf(mx - mn) <= 1e-8f;
  }
}

// pass 2: weighted moments per cluster (block partials, float64)
__global__ void __launch_bounds__(256) k_cluster_moments(const float4* __restrict__ aos, const double* __restrict__ w,
                                                         const int* __restrict__ label, long long n, int K,
                                                         const int* __restrict__ uniform, double* __restrict__ part, int method) {
  __shared__ double s8[8];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  double v[MT_CL_VALS];
#pragma unroll
  for (int j = 0; j < MT_CL_VALS; ++j) v[j] = 0.0;
  int li = -1;
  float wf = 0.f;
  if (i < n) {
    li = label[i];
    wf = (float)w[i];  // weights.float() (particle_filter.py:163)
    const float4 a = aos[4 * i], b = aos[4 * i + 1], c = aos[4 * i + 2];
    const float P[3][4] = {{a.x, a.y, a.z, a.w}, {b.x, b.y, b.z, b.w}, {c.x, c.y, c.z, c.w}};
    if (method == 1) {  // "logmap": weighted mean of the SE(3) tangents (log_map_averaged, pose.py:101-109)
      float lg[6];
      mt_se3_log(P, lg);
      for (int r = 0; r < 6; ++r) v[r] = (double)lg[r];
    } else {
      float q[4];
      mt_so3_to_quat(P, q);
      float e[4] = {q[1], q[2], q[3], q[0]};  // (x, y, z, w)
      if (e[3] < 0.f) e[0] = -e[0], e[1] = -e[1], e[2] = -e[2], e[3] = -e[3];  // antipodal (pose.py:127)
      int t = 0;
      for (int r = 0; r < 4; ++r)
        for (int s = r; s < 4; ++s) v[t++] = (double)e[r] * (double)e[s];
    }
    v[10] = a.w, v[11] = b.w, v[12] = c.w;
    v[13] = (double)a.w * a.w, v[14] = (double)b.w * b.w, v[15] = (double)c.w * c.w;
    v[16] = 1.0, v[17] = 1.0;
  }
  for (int k = 0; k < K; ++k) {
    const double wk = (li == k) ? (uniform[k] ? 1.0 : (double)wf) : 0.0;
    for (int j = 0; j < MT_CL_VALS; ++j) {
      const double x = (j == 17) ? ((li == k) ? 1.0 : 0.0) : v[j] * wk;
      const double s = block_sum_256(x, s8);
      if (threadIdx.x == 0) part[((size_t)blockIdx.x * K + k) * MT_CL_VALS + j] = s;
    }
  }
}

// cyclic Jacobi on a symmetric 4x4 (float64); returns the eigenvector of the largest eigenvalue
__device__ void jacobi4_dominant(double A[4][4], double vec[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq, A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk, A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq, V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int k = 1; k < 4; ++k)
    if (A[k][k] > A[best][best]) best = k;
  for (int k = 0; k < 4; ++k) vec[k] = V[k][best];
}

__global__ void __launch_bounds__(256) k_cluster_final(const double* __restrict__ part, int nblocks, int K,
                                                       float* __restrict__ poses /* K x 16 */, float* __restrict__ stds /* K x 3 */, int method) {
  __shared__ double s8[8];
  const int k = blockIdx.x;
  double s[MT_CL_VALS];
  for (int j = 0; j < MT_CL_VALS; ++j) {  // thread-strided sequential sums, fixed-order combine
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) acc += part[((size_t)b * K + k) * MT_CL_VALS + j];
    s[j] = block_sum_256(acc, s8);
  }
  if (threadIdx.x) return;
  const double W = s[16];
  const double mx = s[10] / W, my = s[11] / W, mz = s[12] / W;
  float* P = poses + 16 * k;
  if (method == 1) {
    const double avg[6] = {s[0] / W, s[1] / W, s[2] / W, s[3] / W, s[4] / W, s[5] / W};
    double T[3][4];
    mt_se3_exp(avg, T);
    for (int r = 0; r < 3; ++r)
      for (int c2 = 0; c2 < 4; ++c2) P[4 * r + c2] = (float)T[r][c2];
    P[12] = P[13] = P[14] = 0.f, P[15] = 1.f;
    const double cx = (double)P[3], cy = (double)P[7], cz = (double)P[11];
    stds[3 * k] = (float)sqrt(fmax(s[13] / W - 2 * cx * mx + cx * cx, 0.0));
    stds[3 * k + 1] = (float)sqrt(fmax(s[14] / W - 2 * cy * my + cy * cy, 0.0));
    stds[3 * k + 2] = (float)sqrt(fmax(s[15] / W - 2 * cz * mz + cz * cz, 0.0));
    return;
  }
  double A[4][4];
  int t = 0;
  for (int r = 0; r < 4; ++r)
    for (int c = r; c < 4; ++c) A[r][c] = A[c][r] = s[t++] / W;
  double e[4];
  jacobi4_dominant(A, e);
  if (e[3] < 0.0) e[0] = -e[0], e[1] = -e[1], e[2] = -e[2], e[3] = -e[3];  // pose.py:140
  const double nq = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]);
  const double x = e[0] / nq, y = e[1] / nq, z = e[2] / nq, w = e[3] / nq;
  P[0] = (float)(1 - 2 * (y * y + z * z)), P[1] = (float)(2 * (x * y - z * w)), P[2] = (float)(2 * (x * z + y * w)), P[3] = (float)mx;
  P[4] = (float)(2 * (x * y + z * w)), P[5] = (float)(1 - 2 * (x * x + z * z)), P[6] = (float)(2 * (y * z - x * w)), P[7] = (float)my;
  P[8] = (float)(2 * (x * z - y * w)), P[9] = (float)(2 * (y * z + x * w)), P[10] = (float)(1 - 2 * (x * x + y * y)), P[11] = (float)mz;
  P[12] = P[13] = P[14] = 0.f, P[15] = 1.f;
  // sqrt(sum w (t - mu)^2 / sum w) with mu the float32 centre the reference subtracts (particle_filter.py:195-204)
  const double cx = (double)P[3], cy = (double)P[7], cz = (double)P[11];
  stds[3 * k] = (float)sqrt(fmax(s[13] / W - 2 * cx * mx + cx * cx, 0.0));
  stds[3 * k + 1] = (float)sqrt(fmax(s[14] / W - 2 * cy * my + cy * cy, 0.0));
  stds[3 * k + 2] = (float)sqrt(fmax(s[15] / W - 2 * cz * mz + cz * cz, 0.0));
}

// ------------------------------------------------------------------------- radix select + compaction
__device__ __forceinline__ unsigned long long f64_order_key(double x) {  // monotone map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// state[0] = prefix (bits decided so far), state[1] = remaining rank within the prefix bucket
__global__ void __launch_bounds__(256) k_select_hist(const double* __restrict__ w, long long n, int largest, int pass,
                                                     const unsigned long long* __restrict__ state, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = state[0];
  const unsigned long long mask = pass ? (~0ull << (shift + 8)) : 0ull;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    unsigned long long k = f64_order_key(w[i]);
    if (largest) k = ~k;
    if ((k & mask) == (prefix & mask)) atomicAdd(&sh[(k >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}
__global__ void k_select_pick(unsigned int* __restrict__ hist, int pass, unsigned long long* __restrict__ state) {
  if (threadIdx.x) return;
  unsigned long long r = state[1];
  const int shift = 56 - 8 * pass;
  for (int b = 0; b < 256; ++b) {
    const unsigned int c = hist[b];
    if (r < c) {
      state[0] |= (unsigned long long)b << shift;
      state[1] = r;
      break;
    }
    r -= c;
  }
  for (int b = 0; b < 256; ++b) hist[b] = 0;
}
// after 8 passes: state[0] = key of the k-th element (0-based rank k-1), state[1] = how many elements
// equal to it are selected besides... (rank inside the bucket of equals)
// selected(i) = key < T || (key == T && (number of equals before i) <= state[1])
__global__ void __launch_bounds__(256) k_select_count(const double* __restrict__ w, long long n, int largest,
                                                      const unsigned long long* __restrict__ state, int* __restrict__ blk /* nblocks x 2 */) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const unsigned long long T = state[0];
  int lt = 0, eq = 0;
  if (i < n) {
    unsigned long long k = f64_order_key(w[i]);
    if (largest) k = ~k;
    lt = k < T, eq = k == T;
  }
  const int nlt = __syncthreads_count(lt), neq = __syncthreads_count(eq);
  if (threadIdx.x == 0) blk[2 * blockIdx.x] = nlt, blk[2 * blockIdx.x + 1] = neq;
}
__global__ void __launch_bounds__(1024) k_select_scan(int* __restrict__ blk, int nblocks) {  // exclusive scans, one block
  __shared__ int s_w[32][2];
  __shared__ int s_base[2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    int v[2] = {b < nblocks ? blk[2 * b] : 0, b < nblocks ? blk[2 * b + 1] : 0};
    const int own[2] = {v[0], v[1]};
    for (int c = 0; c < 2; ++c) {
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v[c], o);
        if (lane >= o) v[c] += u;
      }
      if (lane == 31) s_w[w][c] = v[c];
    }
    __syncthreads();
    int tot[2] = {0, 0};
    for (int c = 0; c < 2; ++c) {
      int carry = s_base[c];
      for (int k = 0; k < 32; ++k) {
        if (k < w) carry += s_w[k][c];
        tot[c] += s_w[k][c];
      }
      if (b < nblocks) blk[2 * b + c] = carry + v[c] - own[c];
    }
    __syncthreads();
    if (threadIdx.x < 2) s_base[threadIdx.x] += tot[threadIdx.x];
    __syncthreads();
  }
}
// writes the selected indices (ascending) to sel[0..k) and the others to keep[0..n-k)
__global__ void __launch_bounds__(256) k_select_scatter(const double* __restrict__ w, long long n, int largest,
                                                        const unsigned long long* __restrict__ state, const int* __restrict__ blk,
                                                        int* __restrict__ sel, int* __restrict__ keep) {
  __shared__ int s_lt[8], s_eq[8];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const unsigned long long T = state[0];
  const long long eq_take = (long long)state[1] + 1;  // equals to select, lowest index first
  int lt = 0, eq = 0;
  if (i < n) {
    unsigned long long k = f64_order_key(w[i]);
    if (largest) k = ~k;
    lt = k < T, eq = k == T;
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const unsigned mlt = __ballot_sync(0xffffffffu, lt), meq = __ballot_sync(0xffffffffu, eq);
  if (lane == 0) s_lt[wp] = __popc(mlt), s_eq[wp] = __popc(meq);
  __syncthreads();
  int blt = blk[2 * blockIdx.x], beq = blk[2 * blockIdx.x + 1];
  for (int k = 0; k < wp; ++k) blt += s_lt[k], beq += s_eq[k];
  const int my_lt = blt + __popc(mlt & ((1u << lane) - 1)), my_eq = beq + __popc(meq & ((1u << lane) - 1));
  if (i >= n) return;
  // selected so far before i: (#lt before i) + min(#eq before i, eq_take)
  const long long sel_before = my_lt + (my_eq < eq_take ? my_eq : eq_take);
  const bool selected = lt || (eq && my_eq < eq_take);
  if (selected) {
    if (sel) sel[sel_before] = (int)i;
  } else if (keep) {
    keep[i - sel_before] = (int)i;
  }
}
#endif
