// cluster_particles(method="euclidean") -- particle_filter.py:208-228: sklearn DBSCAN(eps, min_samples = N/5) on the
// particle translations.  Restated for the GPU with sklearn's semantics (sklearn/cluster/_dbscan.py, _dbscan_inner.pyx):
//   * neighbourhood of i = { j : sum_k (x_ik - x_jk)^2 <= eps^2 } in float64 on the float32 coordinates (the k-d tree's
//     reduced distance), i itself included; i is a core point when it has >= min_samples neighbours;
//   * clusters = connected components of the core points, numbered in the order of their lowest-index core point
//     (the scan over i = 0..n-1 starts a new cluster at every core point that is still unlabelled);
//   * a non-core point within eps of core points takes the label of the first cluster that reaches it, i.e. the
//     lowest-numbered cluster among them (clusters are expanded one after the other); all other points: -1.
// With min_samples = N/5 every neighbourhood that matters holds a fifth of all particles, so the work is inherently
// ~N^2 pair tests; they run as shared-memory tiles (8 instructions per pair, float32 filter with an exact float64
// test inside the rounding band).  Passes: neighbour count -> min-label propagation over core points with pointer
// jumping until nothing changes -> numbering -> border assignment.
// Included by midas_b200.cu (same translation unit).
#pragma once

#define MT_DB_TILE 256

struct DbTest {
  float lo2, hi2;  // float32 filter: below lo2 certainly inside, above hi2 certainly outside
  double eps2;
};
__device__ __forceinline__ bool db_within(const DbTest& t, float xi, float yi, float zi, float xj, float yj, float zj) {
  const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const float d2 = dx * dx + dy * dy + dz * dz;
  if (d2 < t.lo2) return true;
  if (d2 > t.hi2) return false;
  const double ex = (double)xi - (double)xj, ey = (double)yi - (double)yj, ez = (double)zi - (double)zj;
  double e2 = __dmul_rn(ex, ex);
  e2 = __dadd_rn(e2, __dmul_rn(ey, ey));
  e2 = __dadd_rn(e2, __dmul_rn(ez, ez));
  return e2 <= t.eps2;
}

// translations of (n,4,4) float32 poses -> float4 (x, y, z, 0)
__global__ void k_db_points(const float* __restrict__ aos, long long n, float4* __restrict__ pts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pts[i] = make_float4(aos[16 * i + 3], aos[16 * i + 7], aos[16 * i + 11], 0.f);
}

// pass 1: neighbour counts -> core flag; label[i] = i for core points, -1 otherwise
__global__ void __launch_bounds__(MT_DB_TILE) k_db_count(const float4* __restrict__ pts, int n, DbTest t, int min_samples,
                                                         int* __restrict__ label) {
  __shared__ float4 s[MT_DB_TILE];
  const int i = blockIdx.x * MT_DB_TILE + threadIdx.x;
  const float4 p = i < n ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  int cnt = 0;
  for (int j0 = 0; j0 < n; j0 += MT_DB_TILE) {
    __syncthreads();
    s[threadIdx.x] = (j0 + threadIdx.x < n) ? pts[j0 + threadIdx.x] : make_float4(3e30f, 3e30f, 3e30f, 0.f);
    __syncthreads();
    const int m = min(MT_DB_TILE, n - j0);
#pragma unroll 8
    for (int k = 0; k < m; ++k) cnt += db_within(t, p.x, p.y, p.z, s[k].x, s[k].y, s[k].z) ? 1 : 0;
  }
  if (i < n) label[i] = (cnt >= min_samples) ? i : -1;
}

// pass 2 (repeated): core i takes the smallest root among the core points within eps; *changed counts updates
__global__ void __launch_bounds__(MT_DB_TILE) k_db_propagate(const float4* __restrict__ pts, int n, DbTest t,
                                                             const int* __restrict__ label_in, int* __restrict__ label_out,
                                                             int* __restrict__ changed) {
  __shared__ float4 s[MT_DB_TILE];
  const int i = blockIdx.x * MT_DB_TILE + threadIdx.x;
  const float4 p = i < n ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int mine = i < n ? label_in[i] : -1;
  int best = mine;
  // whole blocks of non-core points have nothing to do
  if (__syncthreads_or(mine >= 0)) {
    for (int j0 = 0; j0 < n; j0 += MT_DB_TILE) {
      __syncthreads();
      {
        const int j = j0 + threadIdx.x;
        float4 q = make_float4(3e30f, 3e30f, 3e30f, __int_as_float(-1));
        if (j < n) {
          q = pts[j];
          q.w = __int_as_float(label_in[j]);
        }
        s[threadIdx.x] = q;
      }
      __syncthreads();
      if (mine >= 0) {
        const int m = min(MT_DB_TILE, n - j0);
#pragma unroll 4
        for (int k = 0; k < m; ++k) {
          const int lj = __float_as_int(s[k].w);
          if (lj >= 0 && lj < best && db_within(t, p.x, p.y, p.z, s[k].x, s[k].y, s[k].z)) best = lj;
        }
      }
    }
  }
  if (i < n) {
    label_out[i] = best;
    if (best != mine) atomicAdd(changed, 1);
  }
}
// pointer jumping: label[i] <- label[label[i]] until it is a root (roots satisfy label[r] == r)
__global__ void k_db_jump(int* __restrict__ label, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int l = label[i];
  if (l < 0) return;
  while (true) {
    const int up = label[l];
    if (up == l) break;
    l = up;
  }
  label[i] = l;
}
// numbering: cid[r] = number of roots below r (one block, running offset)
__global__ void __launch_bounds__(1024) k_db_number(const int* __restrict__ label, int n, int* __restrict__ cid, int* __restrict__ n_clusters) {
  __shared__ int s_w[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool root = i < n && label[i] == i;
    const unsigned m = __ballot_sync(0xffffffffu, root);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < w; ++k) off += s_w[k];
    if (root) cid[i] = off + __popc(m & ((1u << lane) - 1));
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < 32; ++k) tot += s_w[k];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_clusters = s_base;
}
// pass 3: final labels (int64 like sklearn's labels_): core -> number of its root; border -> lowest cluster number
// among the core points within eps; noise -> -1
__global__ void __launch_bounds__(MT_DB_TILE) k_db_final(const float4* __restrict__ pts, int n, DbTest t, const int* __restrict__ label,
                                                         const int* __restrict__ cid, long long* __restrict__ out) {
  __shared__ float4 s[MT_DB_TILE];
  const int i = blockIdx.x * MT_DB_TILE + threadIdx.x;
  const float4 p = i < n ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int mine = i < n ? label[i] : 0;
  int best = INT_MAX;
  if (__syncthreads_or(i < n && mine < 0)) {  // only blocks that hold non-core points search
    for (int j0 = 0; j0 < n; j0 += MT_DB_TILE) {
      __syncthreads();
      {
        const int j = j0 + threadIdx.x;
        float4 q = make_float4(3e30f, 3e30f, 3e30f, __int_as_float(-1));
        if (j < n) {
          q = pts[j];
          const int lj = label[j];
          q.w = __int_as_float(lj >= 0 ? cid[lj] : -1);
        }
        s[threadIdx.x] = q;
      }
      __syncthreads();
      if (i < n && mine < 0) {
        const int m = min(MT_DB_TILE, n - j0);
#pragma unroll 4
        for (int k = 0; k < m; ++k) {
          const int cj = __float_as_int(s[k].w);
          if (cj >= 0 && cj < best && db_within(t, p.x, p.y, p.z, s[k].x, s[k].y, s[k].z)) best = cj;
        }
      }
    }
  }
  if (i < n) out[i] = mine >= 0 ? (long long)cid[mine] : (best == INT_MAX ? -1LL : (long long)best);
}
