// Exact 6-D nearest-neighbour search over the codebook keys (SE3_NN, tactile_tree.py:43-58;
// the reference uses a nanoflann k-d tree on the CPU, which is exact -- so is this).
//
// Two cooperating searches, both returning argmin_m mt_key_dist(q, key_m) with ties broken
// towards the lowest codebook index (== np.argmin over the float32 distances):
//
//  1. hint-graph search (one thread per query).  Every codebook key h carries the list of
//     its MT_NBR_K nearest other keys sorted by delta_j = |key_j - key_h|.  For a query q
//     with a hint h (the particle's match one step ago) and d_h = |q - key_h|, the true
//     nearest key lies inside ball(q, best) with best <= d_h, hence within d_h + best of
//     key_h (triangle inequality): scanning the list until delta_j > d_h + best proves the
//     answer.  ~8 candidates per query in steady state.  If the list is exhausted first
//     the query is handed to (2) with the best candidate so far.
//     Rotation vectors flip sign at angle pi, so a pose that was matched to key h can jump
//     2*pi*w away from it in key space; keys within 0.35 rad of pi therefore carry a
//     "partner" (the key nearest to their antipodal image) which is tried as the centre too.
//  2. box-hierarchy search (one warp per query): the keys in k-d tree leaf order, leaves of 32
//     keys and two 32-ary levels of bounding boxes (all six coordinates) above them; the
//     lanes evaluate the bounds of 32 siblings at once and the warp descends best-first until
//     no remaining box can hold a key as close as the best one found.
//
// Float32 rounding: computed distances carry a relative error < 1e-6; every pruning
// test is inflated by 1e-5 (relative) so that no candidate that could win or tie under the
// computed distances is ever skipped.
#pragma once
#include <float.h>
#include <limits.h>
#include <string.h>

#include "mt_math.cuh"

#ifndef MT_NN_BLOCK
#define MT_NN_BLOCK 256
#endif
#define MT_NBR_K 64  // neighbours per key and build pass (the build kernel keeps two per lane); lists are 1, 2 or 4 passes long
#define MT_NBR_K_MAX 256

// Search index of the fallback path: the keys in the leaf order of a balanced k-d tree, 32 consecutive keys per leaf, and two
// levels of 32-ary inner nodes above the leaves; every node stores its axis-aligned bounding box in all six
// key coordinates (12 floats = 3 float4: lo0..lo3 | lo4,lo5,hi0,hi1 | hi2..hi5).
struct BvhParams {
  int n_leaf, n_l1, n_l2;
  float cell;  // largest key extent / 1023 (diagnostic)
};

#if defined(__CUDACC__)
struct NNTables {
  const float4* keys_orig;    // M x 2 float4 (k0..k3 | k4,k5,partner bits,delta_0), original order
  const float4* keys_sorted;  // M x 2 float4 (k0..k3 | k4,k5,original index bits,0), k-d tree leaf order
  const float4* bvh_leaf;     // n_leaf x 3 float4
  const float4* bvh_l1;       // n_l1 x 3 float4 (node j covers leaves 32j .. 32j+31)
  const float4* bvh_l2;       // n_l2 x 3 float4 (node j covers level-1 nodes 32j .. 32j+31)
  const float4* nbr;          // M x K x 2 float4: (k0..k3 | k4,k5,delta,idx bits)
  BvhParams b;
  int M;
  int K;                      // neighbours per key: MT_NBR_K, or a multiple of it for dense codebooks (mt_codebook_upload)
};
#endif

// cell coordinate of a key component: identical float32 formula on host and device so that
// monotonicity arguments about search boxes hold bit-for-bit.
MT_HD int mt_cell_coord(float x, float org, float inv_h, int dim) {
  float f = floorf(MT_FMUL(MT_FSUB(x, org), inv_h));
  int c = (f < 0.f) ? 0 : (f >= (float)dim ? dim - 1 : (int)f);
  return c;
}

// image of a key under v -> v (|v| - 2 pi)/|v| (the same rotation, other sign of the axis);
// false when the rotation angle is further than 0.35 rad from pi.
MT_HD bool mt_key_antipode(const float k[6], float out[6]) {
  const float w = 0.01f, two_pi_w = 6.283185307179586f * w;
  const float n = sqrtf(k[3] * k[3] + k[4] * k[4] + k[5] * k[5]);
  if (!(n > (3.14159265f - 0.35f) * w)) return false;
  const float s = (n - two_pi_w) / n;
  out[0] = k[0], out[1] = k[1], out[2] = k[2];
  out[3] = k[3] * s, out[4] = k[4] * s, out[5] = k[5] * s;
  return true;
}

MT_HD bool mt_better(float d, int i, float best_d, int best_i) { return d < best_d || (d == best_d && i < best_i); }

// stop bound of the hint-graph scan: delta beyond this cannot beat or tie `best`
MT_HD float mt_hint_limit(float dh, float best_d) { return (dh + sqrtf(best_d)) * 1.00001f + 1e-30f; }

// Hint-graph scan over a neighbour list given as plain floats (8 per entry); shared by the
// device kernels and the host harness.  Returns true when (best_d, best_i) is proven exact.
MT_HD bool mt_hint_scan(const float q[6], const float kh[6], int hint, const float* list8, int K, float& best_d,
                        int& best_i) {
  best_d = mt_key_dist(q, kh);
  best_i = hint;
  if (!(best_d == best_d)) {  // NaN query: np.argmin semantics (first NaN) -> index 0
    best_i = 0;
    return true;
  }
  const float dh = sqrtf(best_d);
  float lim = mt_hint_limit(dh, best_d);
  for (int j = 0; j < K; ++j) {
    const float* e = list8 + 8 * j;
    if (e[6] > lim) return true;
    float d = mt_key_dist(q, e);
    int idx;
#if defined(__CUDA_ARCH__)
    idx = __float_as_int(e[7]);
#else
    memcpy(&idx, e + 7, 4);
#endif
    if (mt_better(d, idx, best_d, best_i)) {
      best_d = d;
      best_i = idx;
      lim = mt_hint_limit(dh, best_d);
    }
  }
  return false;
}

// squared distance from q to an axis-aligned box (lo[6] | hi[6]): a lower bound of the distance from q to
// every key inside it.  Shared by the device search and its host restatement below.
MT_HD float mt_box_lower_bound(const float box[12], const float q[6]) {
  float acc = 0.f;
  for (int k = 0; k < 6; ++k) {
    const float d = fmaxf(fmaxf(box[k] - q[k], q[k] - box[6 + k]), 0.f);
    acc = fmaf(d, d, acc);
  }
  return acc;
}
// a node is visited when its bound, shrunk against float32 rounding, does not exceed the best distance so far
// (<=, not <: a key at exactly the best distance with a lower index must still be seen)
MT_HD bool mt_box_may_hold(float bound, float best_d) { return bound * 0.9999f <= best_d; }

#include <algorithm>
#include <numeric>
#include <vector>
// Host side of the search index (mt_codebook_upload): k-d tree leaf order, leaves of 32 keys, two 32-ary levels.
struct MtBvhHost {
  BvhParams bp;
  std::vector<int> order;             // order[j] = original index of the j-th key in leaf (k-d tree) order
  std::vector<float> keys_sorted;     // M x 8 floats: key, original index bits, 0
  std::vector<float> leaf, l1, l2;    // 12 floats per node
};
inline bool mt_bvh_build(const float* h_keys, int M, MtBvhHost& out) {
  float lo[6], hi[6];
  for (int k = 0; k < 6; ++k) lo[k] = FLT_MAX, hi[k] = -FLT_MAX;
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 6; ++k) {
      const float v = h_keys[6 * m + k];
      if (!(fabsf(v) <= FLT_MAX)) return false;  // NaN / Inf key (the scans rely on finite keys: 0 * x == 0)
      lo[k] = std::min(lo[k], v);
      hi[k] = std::max(hi[k], v);
    }
  float maxext = 0.f;
  for (int k = 0; k < 6; ++k) maxext = std::max(maxext, hi[k] - lo[k]);
  if (!(maxext > 0.f) || !(maxext <= FLT_MAX)) maxext = 1e-3f;
  const float cell = maxext / 1023.f;  // (diagnostic only)
  // Order = leaves of a balanced k-d tree: a range of keys is split at the median of its widest coordinate until 32
  // keys are left; the left part always holds a multiple of the node size of its level (32 keys below 1024, 1024 below
  // 32768, ...), so leaves, level-1 and level-2 nodes are whole subtrees with disjoint, tight boxes.  (Round 1 sorted
  // by a 6-D Morton code: a run of 32 consecutive codes straddles cell boundaries of every level, and on the curved
  // key manifold of a real codebook its bounding box is several times larger -- measured on the stand-ins, leaves a
  // search has to open with a perfect seed: cotter pin 13.7 -> 4.7, drill 5.9 -> 2.7, worst case 102 -> 47 / 33 -> 15.)
  // Ties in a coordinate are broken by the index, so the partition is unique; leaves are stored in index order.
  out.order.resize(M);
  std::iota(out.order.begin(), out.order.end(), 0);
  {
    std::vector<std::pair<int, int>> todo;  // [begin, end)
    todo.emplace_back(0, M);
    while (!todo.empty()) {
      const int a = todo.back().first, b = todo.back().second, n = b - a;
      todo.pop_back();
      if (n <= 32) {
        std::sort(out.order.begin() + a, out.order.begin() + b);
        continue;
      }
      float klo[6], khi[6];
      for (int k = 0; k < 6; ++k) klo[k] = FLT_MAX, khi[k] = -FLT_MAX;
      for (int j = a; j < b; ++j)
        for (int k = 0; k < 6; ++k) {
          const float v = h_keys[6 * (size_t)out.order[j] + k];
          klo[k] = std::min(klo[k], v), khi[k] = std::max(khi[k], v);
        }
      int dim = 0;
      for (int k = 1; k < 6; ++k)
        if (khi[k] - klo[k] > khi[dim] - klo[dim]) dim = k;
      const int unit = n <= 1024 ? 32 : (n <= 32768 ? 1024 : (n <= 1048576 ? 32768 : 1048576));
      const int nl = ((n + unit - 1) / unit / 2) * unit;  // >= unit: n > unit here
      std::nth_element(out.order.begin() + a, out.order.begin() + a + nl, out.order.begin() + b, [&](int x, int y) {
        const float vx = h_keys[6 * (size_t)x + dim], vy = h_keys[6 * (size_t)y + dim];
        return vx < vy || (vx == vy && x < y);
      });
      todo.emplace_back(a + nl, b);
      todo.emplace_back(a, a + nl);
    }
  }
  out.keys_sorted.assign(8 * (size_t)M, 0.f);
  for (int m = 0; m < M; ++m) {
    for (int k = 0; k < 6; ++k) out.keys_sorted[8 * (size_t)m + k] = h_keys[6 * (size_t)out.order[m] + k];
    memcpy(&out.keys_sorted[8 * (size_t)m + 6], &out.order[m], sizeof(int));  // original index rides in the padding
  }
  BvhParams& bp = out.bp;
  bp.n_leaf = (M + 31) / 32, bp.n_l1 = (bp.n_leaf + 31) / 32, bp.n_l2 = (bp.n_l1 + 31) / 32, bp.cell = cell;
  auto make_level = [](const std::vector<float>& child, int n_child, int n_node) {
    std::vector<float> o(12 * (size_t)n_node);
    for (int j = 0; j < n_node; ++j) {
      float* b = &o[12 * (size_t)j];
      for (int k = 0; k < 6; ++k) b[k] = FLT_MAX, b[6 + k] = -FLT_MAX;
      for (int ch = 32 * j; ch < std::min(32 * j + 32, n_child); ++ch)
        for (int k = 0; k < 6; ++k) {
          b[k] = std::min(b[k], child[12 * (size_t)ch + k]);
          b[6 + k] = std::max(b[6 + k], child[12 * (size_t)ch + 6 + k]);
        }
    }
    return o;
  };
  std::vector<float> pts(12 * (size_t)M);  // a key is a degenerate box
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 6; ++k) pts[12 * (size_t)m + k] = pts[12 * (size_t)m + 6 + k] = out.keys_sorted[8 * (size_t)m + k];
  out.leaf = make_level(pts, M, bp.n_leaf);
  out.l1 = make_level(out.leaf, bp.n_leaf, bp.n_l1);
  out.l2 = make_level(out.l1, bp.n_l1, bp.n_l2);
  return true;
}
#if !defined(__CUDACC__)
// (host test harness only) Host restatement of nn_bvh_search (same pruning rule, plain nested loops): returns the index, *visited = leaves
// whose keys were evaluated.  seed_i < 0: no candidate.
inline int mt_bvh_search_host(const MtBvhHost& B, int M, const float q[6], float seed_d, int seed_i, int* visited) {
  for (int k = 0; k < 6; ++k)
    if (!(q[k] == q[k])) return 0;  // NaN query: np.argmin semantics
  float best_d = seed_i < 0 ? FLT_MAX : seed_d;
  int best_i = seed_i < 0 ? INT_MAX : seed_i, leaves = 0;
  for (int n2 = 0; n2 < B.bp.n_l2; ++n2) {
    if (!mt_box_may_hold(mt_box_lower_bound(&B.l2[12 * (size_t)n2], q), best_d)) continue;
    for (int n1 = 32 * n2; n1 < std::min(32 * n2 + 32, B.bp.n_l1); ++n1) {
      if (!mt_box_may_hold(mt_box_lower_bound(&B.l1[12 * (size_t)n1], q), best_d)) continue;
      for (int nl = 32 * n1; nl < std::min(32 * n1 + 32, B.bp.n_leaf); ++nl) {
        if (!mt_box_may_hold(mt_box_lower_bound(&B.leaf[12 * (size_t)nl], q), best_d)) continue;
        ++leaves;
        for (int p = 32 * nl; p < std::min(32 * nl + 32, M); ++p) {
          const float* k = &B.keys_sorted[8 * (size_t)p];
          float d = mt_key_dist(q, k);
          if (!(d == d)) continue;  // Inf query coordinates
          int o;
          memcpy(&o, k + 6, sizeof(int));
          if (mt_better(d, o, best_d, best_i)) best_d = d, best_i = o;
        }
      }
    }
  }
  if (visited) *visited = leaves;
  return best_i == INT_MAX ? 0 : best_i;
}
#endif

#if defined(__CUDACC__)
// ------------------------------------------------------------------------- device side
__device__ __forceinline__ int load_key(const float4* __restrict__ t, int i, float k[6], float* extra = nullptr) {
  float4 a = mt_ldk(t + 2 * (size_t)i), b = mt_ldk(t + 2 * (size_t)i + 1);
  k[0] = a.x, k[1] = a.y, k[2] = a.z, k[3] = a.w, k[4] = b.x, k[5] = b.y;
  if (extra) *extra = b.w;     // keys_orig: delta_0 = distance to the nearest other key
  return __float_as_int(b.z);  // partner (keys_orig) / original index (keys_sorted)
}


// 16-byte read-only load that the compiler keeps where it is written (see the pipelined scan below)
__device__ __forceinline__ float4 mt_ldnc(const float4* p) {
  // (default L2 priority: the lists are 2 KB per key, too many to pin; the streamed particle arrays are evict_first,
  // so the lists in use still outlive them)
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Issue as soon as the hint is known (kernel A does it before the motion arithmetic): the
// search below is a chain of dependent reads -- key of the hint, then its neighbour list -- and
// these prefetches turn it into cache hits.
__device__ __forceinline__ void nn_prefetch(const NNTables& T, int hint) {
  if (hint < 0 || hint >= T.M) return;
  const float4* L = T.nbr + (size_t)hint * (2 * T.K);
  asm volatile("prefetch.global.L1 [%0];" ::"l"(T.keys_orig + 2 * (size_t)hint));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(L));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(L + 8));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(L + 16));
}

// Device form of mt_hint_limit: one MUFU.SQRT instead of the IEEE square root (8 instructions + a slow-path call).
// Its error (2^-22 relative; a subnormal best_d flushes to 0, i.e. a term < 1.1e-19 is dropped) is covered by the
// same 1e-5 inflation and by the absolute 1e-18 (far below the spacing of distinct float32 keys).
__device__ __forceinline__ float mt_hint_limit_dev(float dh, float best_d) {
  return fmaf(dh + mt_sqrt_fast(best_d), 1.00001f, 1e-18f);
}
// (distance, index) as one 64-bit word: distances are >= +0, so their bit patterns order like the floats, and the
// unsigned comparison of two words is mt_better() -- smaller distance, ties to the lower index -- in two instructions.
// A NaN distance (0x7fc00000) orders above everything finite and +Inf, like `d < best_d` being false.
__device__ __forceinline__ unsigned long long mt_dist_word(float d, int i) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
}

#ifndef MT_FAST_SCAN
#define MT_FAST_SCAN 1
#endif
// (1) one thread per query, in two parts so that a kernel can cut the scan short and hand the rest to a compacted
// second pass (the scan lengths are long-tailed; in lockstep a warp pays for its longest lane):
//   nn_hint_begin  centre = hint or its near-pi partner, best = that key.  1: proven exact already, -1: no usable
//                  hint (best_i = INT_MAX), 0: scan the centre's list
//   nn_hint_scan   list entries [j0, j1) (multiples of 4) of the centre.  1: proven exact, 0: reached j1
// nn_hint_search = begin + scan over the whole list; false -> needs the box-hierarchy search (best_* = best so far).
__device__ __forceinline__ int nn_hint_begin(const NNTables& T, const float q[6], int hint, float& best_d, int& best_i, int& centre,
                                             float& dh) {
  centre = hint;
  dh = 0.f;
  if (hint < 0 || hint >= T.M) {
    best_d = FLT_MAX;
    best_i = INT_MAX;
    return -1;
  }
  float kh[6], delta0;
  const int partner = load_key(T.keys_orig, hint, kh, &delta0);
  best_d = mt_key_dist(q, kh);
  best_i = hint;
  if (!(best_d == best_d)) {  // NaN query: np.argmin semantics (first NaN) -> index 0
    best_i = 0;
    return 1;
  }
  if (partner >= 0) {  // near angle pi: the pose may have jumped to the other sign of the axis
    float kp[6], dp0;
    load_key(T.keys_orig, partner, kp, &dp0);
    const float dp = mt_key_dist(q, kp);
    if (mt_better(dp, partner, best_d, best_i)) best_d = dp, best_i = partner, centre = partner, delta0 = dp0;
  }
#if MT_FAST_SCAN
  dh = mt_sqrt_fast(best_d);  // enters only the inflated stop bound
  return delta0 > mt_hint_limit_dev(dh, best_d) ? 1 : 0;  // nearest other key of the centre already out of reach?
#else
  dh = sqrtf(best_d);
  return delta0 > mt_hint_limit(dh, best_d) ? 1 : 0;
#endif
}

template <bool BAIL>  // BAIL: long-list codebooks (T.K > MT_NBR_K), see MT_SCAN_BAIL
__device__ __forceinline__ int nn_hint_scan_t(const NNTables& T, const float q[6], int centre, float dh, int j0, int j1, float& best_d,
                                              int& best_i) {
#if MT_FAST_SCAN
  float lim = mt_hint_limit_dev(dh, best_d);
  unsigned long long bw = mt_dist_word(best_d, best_i);
#else
  float lim = mt_hint_limit(dh, best_d);
#endif
  const int K = BAIL ? T.K : MT_NBR_K;  // (!BAIL <=> T.K == MT_NBR_K: the list stride and the loop bounds are constants then)
  const float4* __restrict__ L = T.nbr + (size_t)centre * (2 * K);
  // Software pipeline: the two entries of the next half trip are requested before the current ones are evaluated, so
  // a trip waits for loads issued ~40 instructions (x the other resident warps) earlier.  The loads are volatile asm
  // (plain ld.global.nc underneath, i.e. cached in L1 like __ldg) so that neither NVVM nor ptxas sinks them below the
  // exit tests.  Two register sets (x*, y*) alternate between "being evaluated" and "in flight".
  float4 xa0 = mt_ldnc(L + 2 * j0), xb0 = mt_ldnc(L + 2 * j0 + 1), xa1 = mt_ldnc(L + 2 * j0 + 2), xb1 = mt_ldnc(L + 2 * j0 + 3);
  float4 ya0, yb0, ya1, yb1;
  // (64-entry lists -- sparse codebooks, where an inconclusive scan is a 1-in-2000 event -- run the plain loop, BAIL =
  // false: there the test only added instructions to every trip, +1.6 us on the drill)
  [[maybe_unused]] float dlast = FLT_MAX;
  if (BAIL) dlast = mt_ldnc(L + 2 * (K - 1) + 1).z;  // delta of the list's last entry (requested with the first trip)
#if MT_FAST_SCAN
#define MT_SCAN_DONE(r)                                                    \
  {                                                                        \
    best_d = __uint_as_float((unsigned)(bw >> 32)), best_i = (int)(unsigned)bw; \
    return r;                                                              \
  }
#define MT_SCAN_ENTRY(A, B)                                                                             \
  {                                                                                                     \
    if (B.z > lim) MT_SCAN_DONE(1)                                                                      \
    const float k[6] = {A.x, A.y, A.z, A.w, B.x, B.y};                                                  \
    const float d = mt_key_dist(q, k);                                                                  \
    const unsigned long long w = mt_dist_word(d, __float_as_int(B.w));                                  \
    if (w < bw) bw = w;                                                                                 \
  }
// the stop bound only ever shrinks, so it may lag: it is refreshed once per pair of entries, unconditionally
// (three instructions per pair instead of four predicated ones per entry)
#define MT_SCAN_LIMIT() lim = mt_hint_limit_dev(dh, __uint_as_float((unsigned)(bw >> 32)));
// Hopeless scans leave early: the list certifies the answer only if its last delta exceeds d_h + sqrt(best); once 12
// entries have been seen (a decent seed for the box search) a scan whose bound would still be beyond the end of the
// list even if `best` halved is handed to the queue right away instead of walking the whole list first.  (A wrong
// guess costs a box search, never the result.)  On dense codebooks (256-entry lists, 30 % of the scans inconclusive)
// every warp used to walk all 64 trips for its hopeless lanes.
#define MT_SCAN_BAIL()                                                                                          \
  if (BAIL && j >= 8 && fmaf(0.7071f, mt_sqrt_fast(__uint_as_float((unsigned)(bw >> 32))), dh) > dlast) MT_SCAN_DONE(0)
#else
#define MT_SCAN_LIMIT()
#define MT_SCAN_BAIL()
#define MT_SCAN_DONE(r) return r;
#define MT_SCAN_ENTRY(A, B)                                                                             \
  {                                                                                                     \
    if (B.z > lim) return 1;                                                                            \
    const float k[6] = {A.x, A.y, A.z, A.w, B.x, B.y};                                                  \
    const float d = mt_key_dist(q, k);                                                                  \
    const int idx = __float_as_int(B.w);                                                                \
    if (mt_better(d, idx, best_d, best_i)) best_d = d, best_i = idx, lim = mt_hint_limit(dh, d);        \
  }
#endif
#pragma unroll 1
  for (int j = j0; j < j1; j += 4) {
    ya0 = mt_ldnc(L + 2 * j + 4), yb0 = mt_ldnc(L + 2 * j + 5), ya1 = mt_ldnc(L + 2 * j + 6), yb1 = mt_ldnc(L + 2 * j + 7);
    MT_SCAN_ENTRY(xa0, xb0)
    MT_SCAN_ENTRY(xa1, xb1)
    MT_SCAN_LIMIT()
    const int jn = min(j + 4, K - 2);  // past the end of the list: a harmless re-read, never used
    xa0 = mt_ldnc(L + 2 * jn), xb0 = mt_ldnc(L + 2 * jn + 1), xa1 = mt_ldnc(L + 2 * jn + 2), xb1 = mt_ldnc(L + 2 * jn + 3);
    MT_SCAN_ENTRY(ya0, yb0)
    MT_SCAN_ENTRY(ya1, yb1)
    MT_SCAN_LIMIT()
    MT_SCAN_BAIL()
  }
  MT_SCAN_DONE(0)
#undef MT_SCAN_BAIL
#undef MT_SCAN_ENTRY
#undef MT_SCAN_DONE
#undef MT_SCAN_LIMIT
}

__device__ __forceinline__ int nn_hint_scan(const NNTables& T, const float q[6], int centre, float dh, int j0, int j1, float& best_d,
                                            int& best_i) {
  return T.K > MT_NBR_K ? nn_hint_scan_t<true>(T, q, centre, dh, j0, j1, best_d, best_i)
                        : nn_hint_scan_t<false>(T, q, centre, dh, j0, j1, best_d, best_i);
}
// the whole list
__device__ __forceinline__ int nn_hint_scan_all(const NNTables& T, const float q[6], int centre, float dh, float& best_d, int& best_i) {
  return T.K > MT_NBR_K ? nn_hint_scan_t<true>(T, q, centre, dh, 0, T.K, best_d, best_i)
                        : nn_hint_scan_t<false>(T, q, centre, dh, 0, MT_NBR_K, best_d, best_i);
}

__device__ __forceinline__ bool nn_hint_search(const NNTables& T, const float q[6], int hint, float& best_d, int& best_i) {
  int centre;
  float dh;
  const int st = nn_hint_begin(T, q, hint, best_d, best_i, centre, dh);
  if (st) return st > 0;
  return nn_hint_scan_all(T, q, centre, dh, best_d, best_i) != 0;
}

// bound of one node (3 float4 = lo[6] | hi[6])
__device__ __forceinline__ float bvh_lower_bound(const float4* __restrict__ node, const float q[6]) {
  const float4 a = mt_ldk(node), b = mt_ldk(node + 1), c = mt_ldk(node + 2);
  const float box[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  return mt_box_lower_bound(box, q);
}

// warp-uniform pick of the smallest remaining bound: returns its lane (bounds are >= 0, so their bit
// patterns order like the floats) and the bound itself; the picked lane's entry is set to +Inf.
__device__ __forceinline__ int bvh_pick(float& lb, float& picked) {
  const unsigned bits = __float_as_uint(lb);
  const unsigned m = __reduce_min_sync(0xffffffffu, bits);
  const int who = __ffs(__ballot_sync(0xffffffffu, bits == m)) - 1;
  picked = __uint_as_float(m);
  if ((int)(threadIdx.x & 31) == who) lb = __int_as_float(0x7f800000);
  return who;
}

// (2) all 32 lanes of a warp, one query (q / best_* warp-uniform): exact nearest key given an optional
// candidate (best_i == INT_MAX: none).  Best-first descent through the three levels: the lanes evaluate the
// bounds of up to 32 siblings at once, the warp then visits them in order of increasing bound and stops a
// level as soon as the smallest remaining bound exceeds the best distance found so far (bounds are scaled
// by 0.9999 before the comparison: float32 rounding of the bound must never hide a key, and a key at exactly
// the best distance with a lower index still has to be seen).  A leaf is one coalesced 1 KB read: 32 keys,
// one per lane.  The work depends on how many boxes the ball (q, sqrt(best)) touches in all six coordinates --
// also for queries far off the key manifold, which a translation-only grid cannot prune.
// stats (nullable): [0] += leaves visited, [3] = max leaves visited by one query.
#ifndef MT_BVH_BATCH
#define MT_BVH_BATCH 4
#endif
// warp-wide (best_d, best_i) = lexicographic min over lanes
__device__ __forceinline__ void warp_best(float& d, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, d, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (mt_better(od, oi, d, i)) d = od, i = oi;
  }
}

__device__ __noinline__ int nn_bvh_search(const NNTables& T, const float q[6], float best_d, int best_i, int* stats) {
  if (!(q[0] == q[0]) || !(q[1] == q[1]) || !(q[2] == q[2]) || !(q[3] == q[3]) || !(q[4] == q[4]) || !(q[5] == q[5]))
    return 0;  // NaN query: np.argmin semantics
  const int lane = threadIdx.x & 31;
  const float INF = __int_as_float(0x7f800000);
  if (best_i == INT_MAX) best_d = FLT_MAX;
  int leaves = 0;
  for (int c2 = 0; c2 < T.b.n_l2; c2 += 32) {
    float lb2 = (c2 + lane < T.b.n_l2) ? bvh_lower_bound(T.bvh_l2 + 3 * (size_t)(c2 + lane), q) : INF;
    for (;;) {
      float v2;
      const int w2 = bvh_pick(lb2, v2);
      if (!mt_box_may_hold(v2, best_d)) break;
      const int i1 = 32 * (c2 + w2) + lane;
      float lb1 = (i1 < T.b.n_l1) ? bvh_lower_bound(T.bvh_l1 + 3 * (size_t)i1, q) : INF;
      for (;;) {
        float v1;
        const int w1 = bvh_pick(lb1, v1);
        if (!mt_box_may_hold(v1, best_d)) break;
        const int il = 32 * (32 * (c2 + w2) + w1) + lane;
        float lbl = (il < T.b.n_leaf) ? bvh_lower_bound(T.bvh_leaf + 3 * (size_t)il, q) : INF;
        for (;;) {
          // up to MT_BVH_BATCH qualifying leaves per round: their keys are fetched together (one memory
          // latency instead of one per leaf) and reduced once
          int base[MT_BVH_BATCH];
          int nb = 0;
#pragma unroll
          for (int bq = 0; bq < MT_BVH_BATCH; ++bq) {
            base[bq] = -1;
            if (nb == bq) {  // the previous pick qualified
              float vl;
              const int wl = bvh_pick(lbl, vl);
              if (mt_box_may_hold(vl, best_d)) base[bq] = 32 * (32 * (32 * (c2 + w2) + w1) + wl), ++nb;
            }
          }
          if (nb == 0) break;
          float d = INF;
          int o = INT_MAX;
#pragma unroll
          for (int bq = 0; bq < MT_BVH_BATCH; ++bq) {
            const int pidx = base[bq] + lane;
            if (base[bq] >= 0 && pidx < T.M) {
              float k[6];
              const int ob = load_key(T.keys_sorted, pidx, k);
              float db = mt_key_dist(q, k);
              if (!(db == db)) db = INF;  // Inf query coordinates: Inf - Inf
              if (mt_better(db, ob, d, o)) d = db, o = ob;
            }
          }
          warp_best(d, o);
          if (o != INT_MAX && mt_better(d, o, best_d, best_i)) best_d = d, best_i = o;
          leaves += nb;
          if (nb < MT_BVH_BATCH) break;  // the last pick did not qualify: nothing left at this level
        }
      }
    }
  }
  if (stats && lane == 0) {
    atomicAdd(stats, leaves);
    atomicMax(stats + 3, leaves);
  }
  return best_i == INT_MAX ? 0 : best_i;  // Inf query: every distance is Inf/NaN
}

// The same search by a TEAM of warps (one query, `nparts` warps of one block, this warp = `part`): every warp walks the
// same best-first sequence of level-1 nodes but only opens those whose rank in that sequence is congruent to `part`;
// the team's best (distance bits << 32 | index, mt_dist_word) lives in shared memory, is improved with atomicMin and
// re-read before every pruning test.  A pruning test against a stale (larger) best only opens a box too many, never
// one too few, so the result is the same exact argmin.  The step waits for its slowest search (a single warp needs
// ~0.65 us per leaf, 25 us for the 38-leaf worst case); a team cuts that tail.
__device__ __noinline__ void nn_bvh_search_team(const NNTables& T, const float q[6], int part, int nparts,
                                                unsigned long long* s_best, int* s_leaves) {
  const int lane = threadIdx.x & 31;
  const float INF = __int_as_float(0x7f800000);
  float best_d;
  int best_i;
#define MT_TEAM_REFRESH()                                                   \
  {                                                                         \
    const unsigned long long w_ = *(volatile unsigned long long*)s_best;    \
    best_d = __uint_as_float((unsigned)(w_ >> 32)), best_i = (int)(unsigned)w_; \
  }
  MT_TEAM_REFRESH()
  int leaves = 0;
  for (int c2 = 0; c2 < T.b.n_l2; c2 += 32) {
    float lb2 = (c2 + lane < T.b.n_l2) ? bvh_lower_bound(T.bvh_l2 + 3 * (size_t)(c2 + lane), q) : INF;
    for (;;) {
      float v2;
      const int w2 = bvh_pick(lb2, v2);
      MT_TEAM_REFRESH()
      if (!mt_box_may_hold(v2, best_d)) break;
      const int i1 = 32 * (c2 + w2) + lane;
      float lb1 = (i1 < T.b.n_l1) ? bvh_lower_bound(T.bvh_l1 + 3 * (size_t)i1, q) : INF;
      for (int rank = 0;; ++rank) {
        float v1;
        const int w1 = bvh_pick(lb1, v1);
        MT_TEAM_REFRESH()
        if (!mt_box_may_hold(v1, best_d)) break;
        if (rank % nparts != part) continue;  // a team mate's node
        const int il = 32 * (32 * (c2 + w2) + w1) + lane;
        float lbl = (il < T.b.n_leaf) ? bvh_lower_bound(T.bvh_leaf + 3 * (size_t)il, q) : INF;
        for (;;) {
          int base[MT_BVH_BATCH];
          int nb = 0;
#pragma unroll
          for (int bq = 0; bq < MT_BVH_BATCH; ++bq) {
            base[bq] = -1;
            if (nb == bq) {
              float vl;
              const int wl = bvh_pick(lbl, vl);
              if (mt_box_may_hold(vl, best_d)) base[bq] = 32 * (32 * (32 * (c2 + w2) + w1) + wl), ++nb;
            }
          }
          if (nb == 0) break;
          float d = INF;
          int o = INT_MAX;
#pragma unroll
          for (int bq = 0; bq < MT_BVH_BATCH; ++bq) {
            const int pidx = base[bq] + lane;
            if (base[bq] >= 0 && pidx < T.M) {
              float k[6];
              const int ob = load_key(T.keys_sorted, pidx, k);
              float db = mt_key_dist(q, k);
              if (!(db == db)) db = INF;
              if (mt_better(db, ob, d, o)) d = db, o = ob;
            }
          }
          const unsigned dm = __reduce_min_sync(0xffffffffu, __float_as_uint(d));  // distances are >= +0
          const unsigned om = __reduce_min_sync(0xffffffffu, __float_as_uint(d) == dm ? (unsigned)o : 0x7fffffffu);
          if (om != 0x7fffffffu && mt_better(__uint_as_float(dm), (int)om, best_d, best_i) && lane == 0)
            atomicMin(s_best, ((unsigned long long)dm << 32) | om);
          __syncwarp();
          MT_TEAM_REFRESH()
          leaves += nb;
          if (nb < MT_BVH_BATCH) break;
        }
      }
    }
  }
#undef MT_TEAM_REFRESH
  if (lane == 0) atomicAdd(s_leaves, leaves);
}

// (Measured and dropped: fetching the leaf boxes of four level-1 nodes per round trip, eight leaves per round trip, and
// taking children in lane order from a ballot instead of best-first -- the first two were 10 % slower (a search is one
// warp's dependent instruction stream, ~1 us per leaf; wider rounds add instructions, not overlap), the last one cut the
// average by 5 % but tripled the leaves of the worst search (104 against 37), and the slowest search is what the step waits for.)
__device__ __forceinline__ int nn_search_warp(const NNTables& T, const float q[6], float best_d, int best_i) {
  return nn_bvh_search(T, q, best_d, best_i, nullptr);
}

// Driver used by every kernel that assigns neighbours: each thread first tries the hint
// graph; the leftovers of a warp are then served one at a time by the whole warp (no block
// barrier: warps that had no leftovers carry on).  All 32 lanes must call.
__device__ __forceinline__ int nn_assign(const NNTables& T, bool active, const float q[6], int hint, int* fallback_counter) {
  float bd = FLT_MAX;
  int bi = INT_MAX;
  const bool todo = active && !nn_hint_search(T, q, hint, bd, bi);
  unsigned m = __ballot_sync(0xffffffffu, todo);
  if (m && fallback_counter && (threadIdx.x & 31) == 0) atomicAdd(fallback_counter, __popc(m));
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    float qq[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) qq[k] = __shfl_sync(0xffffffffu, q[k], src);
    const float sd = __shfl_sync(0xffffffffu, bd, src);
    const int si = __shfl_sync(0xffffffffu, bi, src);
    const int res = nn_search_warp(T, qq, sd, si);
    if ((threadIdx.x & 31) == src) bi = res;
  }
  return bi;
}

// Build kernel (codebook upload).  One warp per key; the running sorted list of 64 entries lives
// in registers, two per lane (slot l in `lo`, slot 32 + l in `hi`); the codebook streams
// through shared memory.
//   MODE 0: the MT_NBR_K nearest other keys of every key, ascending (distance, index)
//   MODE 1: the key nearest to the antipodal image of every near-pi key -> partner[h]
//   MODE 2: the k <= 64 nearest keys of nq external query keys (SE3_NN with nn > 1, tactile_tree.py:43-52):
//           out_idx[h * k + j] = index of the j-th nearest key of query h, ascending (distance, index) -- exhaustive,
//           every query streams the whole codebook (meant for the handful of queries such calls make)
//   MODE 0 builds lists longer than 64 in passes of 64: pass p writes entries [64 p, 64 p + 64) of the K-entry list and
//   only admits keys that come strictly after the previous pass's last entry in (squared distance, index) order
//   (`bound`: 2 floats per key = that entry's squared distance and index bits, read and then overwritten).
template <int MODE>
__global__ void __launch_bounds__(256) k_build_nbr(const float4* __restrict__ keys, int M, float4* __restrict__ nbr,
                                                   int* __restrict__ partner, const float* __restrict__ qkeys = nullptr,
                                                   int nq = 0, int k_out = 0, int* __restrict__ out_idx = nullptr,
                                                   int K = MT_NBR_K, int pass = 0, float2* __restrict__ bound = nullptr) {
  static_assert(MT_NBR_K == 64, "two list slots per lane");
  constexpr bool PARTNER = MODE == 1;
  constexpr bool QUERY = MODE == 2;
  __shared__ float4 sk[2 * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.x * 8 + warp;
  const int H = QUERY ? nq : M;
  float kh[6] = {0, 0, 0, 0, 0, 0};
  bool act = h < H;
  if (act) {
    if (QUERY) {
#pragma unroll
      for (int k = 0; k < 6; ++k) kh[k] = qkeys[6 * (size_t)h + k];
    } else {
      load_key(keys, h, kh);
    }
    if (PARTNER) {
      float ka[6];
      act = mt_key_antipode(kh, ka);
#pragma unroll
      for (int k = 0; k < 6; ++k) kh[k] = act ? ka[k] : kh[k];
    }
  }
  float lo_v = FLT_MAX, hi_v = FLT_MAX;  // squared distances, ascending over (lo[0..31], hi[0..31])
  int lo_i = -1, hi_i = -1;
  float lb_d = -1.f;  // admit only (d, m) > (lb_d, lb_i) lexicographically
  int lb_i = -1;
  if (MODE == 0 && pass > 0 && act) {
    const float2 b = bound[h];
    lb_d = b.x, lb_i = __float_as_int(b.y);
  }
  for (int m0 = 0; m0 < M; m0 += 256) {
    const int cnt = min(256, M - m0);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * cnt; t += 256) sk[t] = keys[2 * (size_t)m0 + t];
    __syncthreads();
    if (!act) continue;
    for (int s0 = 0; s0 < cnt; s0 += 32) {
      const int m = m0 + s0 + lane;
      float d = FLT_MAX;
      if (s0 + lane < cnt && (PARTNER || QUERY || m != h)) {
        const float4 a = sk[2 * (s0 + lane)], b = sk[2 * (s0 + lane) + 1];
        const float k[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
        d = mt_key_dist(kh, k);
        if (!(d == d)) d = FLT_MAX;
        if (MODE == 0 && (d < lb_d || (d == lb_d && m <= lb_i))) d = FLT_MAX;  // taken by an earlier pass
      }
      unsigned cand = __ballot_sync(0xffffffffu, d < __shfl_sync(0xffffffffu, hi_v, 31));
      while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        const float cd = __shfl_sync(0xffffffffu, d, src);
        if (!(cd < __shfl_sync(0xffffffffu, hi_v, 31))) continue;
        // stable position: equal distances keep index order
        const int pos = __popc(__ballot_sync(0xffffffffu, lo_v <= cd)) + __popc(__ballot_sync(0xffffffffu, hi_v <= cd));
        const float lo31v = __shfl_sync(0xffffffffu, lo_v, 31);
        const int lo31i = __shfl_sync(0xffffffffu, lo_i, 31);
        const float lv = __shfl_up_sync(0xffffffffu, lo_v, 1), hv = __shfl_up_sync(0xffffffffu, hi_v, 1);
        const int li = __shfl_up_sync(0xffffffffu, lo_i, 1), hi2 = __shfl_up_sync(0xffffffffu, hi_i, 1);
        if (pos < 32) {
          hi_v = lane ? hv : lo31v, hi_i = lane ? hi2 : lo31i;
          if (lane > pos) lo_v = lv, lo_i = li;
          if (lane == pos) lo_v = cd, lo_i = m0 + s0 + src;
        } else {
          if (lane > pos - 32) hi_v = hv, hi_i = hi2;
          if (lane == pos - 32) hi_v = cd, hi_i = m0 + s0 + src;
        }
      }
    }
  }
  if (h >= H) return;
  if constexpr (PARTNER) {
    if (lane == 0) partner[h] = act ? lo_i : -1;
  } else if constexpr (QUERY) {
    if (lane < k_out) out_idx[(size_t)h * k_out + lane] = lo_i;
    if (32 + lane < k_out) out_idx[(size_t)h * k_out + 32 + lane] = hi_i;
  } else {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int idx = half ? hi_i : lo_i;
    const float val = half ? hi_v : lo_v;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, __int_as_float(0x7f800000), __int_as_float(-1));
    if (idx >= 0) {
      a = keys[2 * (size_t)idx];
      const float4 kb = keys[2 * (size_t)idx + 1];
      b = make_float4(kb.x, kb.y, sqrtf(val), __int_as_float(idx));
    }
    const size_t slot = (size_t)h * K + (size_t)MT_NBR_K * pass + 32 * half + lane;
    nbr[slot * 2] = a;
    nbr[slot * 2 + 1] = b;
    if (bound && half == 1 && lane == 31) bound[h] = make_float2(idx >= 0 ? val : FLT_MAX, __int_as_float(idx >= 0 ? idx : INT_MAX));
  }
  }
}

// keys_orig padding: (partner index bits, delta_0 = distance to the nearest other key)
__global__ void k_set_partner(float4* __restrict__ keys, int M, const int* __restrict__ partner,
                              const float4* __restrict__ nbr, int K) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  keys[2 * (size_t)m + 1].z = __int_as_float(partner[m]);
  keys[2 * (size_t)m + 1].w = nbr[(size_t)m * K * 2 + 1].z;
}
// the 64th-neighbour distance of every key (decides the list length at upload)
__global__ void k_nbr_last_delta(const float4* __restrict__ nbr, int M, int K, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) out[m] = nbr[((size_t)m * K + MT_NBR_K - 1) * 2 + 1].z;
}
// re-strides 64-entry lists into the first 64 entries of K-entry lists
__global__ void k_nbr_restride(const float4* __restrict__ in, int M, int K, float4* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)M * MT_NBR_K * 2) return;
  const size_t m = t / (MT_NBR_K * 2), r = t % (MT_NBR_K * 2);
  out[m * K * 2 + r] = in[t];
}

// rank[m] = position of key m in the cell-sorted order (engine: spatial sort of particles)
__global__ void k_key_rank(const float4* __restrict__ keys_sorted, int M, int* __restrict__ rank) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < M) rank[__float_as_int(keys_sorted[2 * (size_t)p + 1].z)] = p;
}
#endif  // __CUDACC__
