// Tactile code network forward pass (TCN = MinkLoc3D: sparse 3-D FPN + GeM) on sm_100a.
//
// Replaces the MinkowskiEngine calls of contrib/tcn_minkloc/{tcn,minkloc,minkfpn}.py for the
// shipped configuration (config/tcn/default.yaml: planes 32,64,64, layers 1,1,1, one top-down
// block, conv0 kernel 5, 256-d output).  Sparse tensors are (sorted unique 64-bit coordinate
// keys, row-major float32 features); a coordinate map is an open-addressing hash table
// key -> row.  One generic kernel evaluates every convolution: one warp per output point,
// lanes over output channels, kernel offsets resolved by hash look-ups (gather form: no
// atomics, fixed summation order), BatchNorm(eval) / residual / ReLU fused in the epilogue.
// Semantics restated in oracle/tcn_oracle.py (parity unpinned against MinkowskiEngine itself).
//
// Included at the end of midas_b200.cu (same translation unit: shares set_err / CK).
#pragma once

#define TCN_EMPTY 0xFFFFFFFFFFFFFFFFull
#define TCN_NCONV 16
#define TCN_NBN 13
// conv ids
#define TCN_CONV0 0
#define TCN_DOWN(s) (1 + (s))
#define TCN_BLK_C1(s) (4 + 2 * (s))
#define TCN_BLK_C2(s) (5 + 2 * (s))
#define TCN_BLK_DS(s) (10 + (s))
#define TCN_LAT0 13
#define TCN_TCONV 14
#define TCN_LAT1 15
// bn ids
#define TCN_BN0 0
#define TCN_BN_DOWN(s) (1 + (s))
#define TCN_BN_N1(s) (4 + 2 * (s))
#define TCN_BN_N2(s) (5 + 2 * (s))
#define TCN_BN_DS(s) (10 + (s))

struct TcnConv {
  float* w;  // (kvol, cin, cout)
  int kvol, cin, cout;
};
struct TcnBn {
  float* scale;  // gamma / sqrt(var + eps)
  float* shift;  // beta - mean * scale
  int c;
};
struct TcnTable {
  unsigned long long* keys;
  int* vals;
  unsigned mask;  // capacity - 1 (power of two)
};

struct mt_tcn {
  int device, max_points, max_batch;
  TcnConv conv[TCN_NCONV];
  TcnBn bn[TCN_NBN];
  float gem_p, gem_eps;
  unsigned cap;            // hash capacity per level
  TcnTable tab[4];         // level 0..3 (tensor stride 1,2,4,8)
  unsigned long long* keys[4];
  int* d_n;                // [4] active points per level (device)
  float* pool;             // feature scratch
  double* gem_part;        // max_batch x TCN_GEM_SLICES x 256 GeM partial sums
  size_t pool_floats;
};

// ---- coordinate keys: 9 bits batch | 3 x 18 bits (coordinate + 2^17)
__host__ __device__ __forceinline__ unsigned long long tcn_pack(int b, int x, int y, int z) {
  const unsigned long long o = 1ull << 17;
  return ((unsigned long long)b << 54) | ((unsigned long long)(x + o) << 36) | ((unsigned long long)(y + o) << 18) |
         (unsigned long long)(z + o);
}
__device__ __forceinline__ void tcn_unpack(unsigned long long k, int& b, int& x, int& y, int& z) {
  const int o = 1 << 17;
  b = (int)(k >> 54);
  x = (int)((k >> 36) & 0x3FFFF) - o;
  y = (int)((k >> 18) & 0x3FFFF) - o;
  z = (int)(k & 0x3FFFF) - o;
}
__device__ __forceinline__ int tcn_floor_to(int c, int m) {  // floor(c / m) * m for m = 2^j
  return c & ~(m - 1);                                       // two's complement: exact for negatives too
}
__device__ __forceinline__ unsigned tcn_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int tcn_lookup(const TcnTable& t, unsigned long long key) {
  unsigned s = tcn_hash(key) & t.mask;
  for (;;) {
    const unsigned long long k = t.keys[s];
    if (k == key) return t.vals[s];
    if (k == TCN_EMPTY) return -1;
    s = (s + 1) & t.mask;
  }
}
// insert (or find) `key`; returns its slot
__device__ __forceinline__ unsigned tcn_insert(const TcnTable& t, unsigned long long key) {
  unsigned s = tcn_hash(key) & t.mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(t.keys + s, TCN_EMPTY, key);
    if (prev == TCN_EMPTY || prev == key) return s;
    s = (s + 1) & t.mask;
  }
}

__global__ void k_tcn_clear(TcnTable t) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= t.mask) t.keys[i] = TCN_EMPTY, t.vals[i] = INT_MAX;
}

// level 0: table[key_i] = i
__global__ void k_tcn_index(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n, TcnTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *d_n) return;
  t.vals[tcn_insert(t, keys[i])] = i;
}

// coarser level, step 1: every child registers with its parent (floor to 2*stride); the parent
// remembers its first child (smallest row) -- a deterministic representative
__global__ void k_tcn_parent_first(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n, int stride2, TcnTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *d_n) return;
  int b, x, y, z;
  tcn_unpack(keys[i], b, x, y, z);
  const unsigned long long pk = tcn_pack(b, tcn_floor_to(x, stride2), tcn_floor_to(y, stride2), tcn_floor_to(z, stride2));
  atomicMin(t.vals + tcn_insert(t, pk), i);
}

// step 2 (one block): parents ordered by their first child -> rows; writes the parent keys, the row
// into the table and the level's point count
__global__ void __launch_bounds__(1024) k_tcn_parent_rows(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n,
                                                          int stride2, TcnTable t, unsigned long long* __restrict__ pkeys,
                                                          int* __restrict__ d_n_next) {
  __shared__ int s_w[32];
  __shared__ int s_base;
  const int n = *d_n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    unsigned long long pk = 0;
    unsigned slot = 0;
    int first = 0;
    if (i < n) {
      int b, x, y, z;
      tcn_unpack(keys[i], b, x, y, z);
      pk = tcn_pack(b, tcn_floor_to(x, stride2), tcn_floor_to(y, stride2), tcn_floor_to(z, stride2));
      slot = tcn_insert(t, pk);  // exists: returns its slot
      first = (t.vals[slot] == i);
    }
    int v = first;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) s_w[w] = v;
    __syncthreads();
    int carry = s_base;
    for (int k = 0; k < w; ++k) carry += s_w[k];
    int total = 0;
    for (int k = 0; k < 32; ++k) total += s_w[k];
    if (first) pkeys[carry + v - 1] = pk;
    __syncthreads();
    // the row replaces the first-child marker only after every thread of this pass has read it
    if (first) t.vals[slot] = -(carry + v - 1) - 2;  // tagged (negative) until the final pass below
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *d_n_next = s_base;
}
__global__ void k_tcn_untag(TcnTable t) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= t.mask && t.keys[i] != TCN_EMPTY && t.vals[i] < 0) t.vals[i] = -(t.vals[i] + 2);
}

// Generic sparse convolution, one warp per output point.
//   MODE 0: out[p] = sum_i in[p + off_i] W[i]; off_i = (o - centre) * dil per axis, x fastest,
//           centre = k/2 for odd k, 0 for even k (regular, strided-down and 1x1 convolutions)
//   MODE 1: transposed k=2,s=2: out[p] = in[parent(p)] W[i(p - parent(p))], dil = stride of p
// epilogue: * scale + shift (BatchNorm eval) -> + residual -> ReLU
template <int MODE, int WPP>  // WPP warps per output point: the input channels are split between them
__global__ void __launch_bounds__(256) k_tcn_conv(const unsigned long long* __restrict__ out_keys, const int* __restrict__ d_nout,
                                                  TcnTable in_tab, const float* __restrict__ in_feat, int cin,
                                                  const float* __restrict__ W, int cout, int k, int dil,
                                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                                  const float* __restrict__ residual, int relu, int accumulate,
                                                  float* __restrict__ out_feat) {
  constexpr int PPB = 8 / WPP;  // points per 256-thread block
  __shared__ float s_acc[WPP > 1 ? 8 : 1][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = warp % WPP;  // which slice of the input channels this warp multiplies
  const int p = blockIdx.x * PPB + warp / WPP;
  const bool live = p < *d_nout;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (live) {
    int b, x, y, z;
    tcn_unpack(out_keys[p], b, x, y, z);
    const int kvol = (MODE == 1) ? 1 : k * k * k;
    const int centre = (k & 1) ? (k >> 1) : 0;
    const int cslice = (cin + WPP - 1) / WPP, cbeg = part * cslice, cend = min(cin, cbeg + cslice);
    for (int i0 = 0; i0 < kvol; i0 += 32) {
      // the 32 lanes resolve 32 kernel offsets at once (the look-ups are independent memory chains)
      const int i = i0 + lane;
      int row = -1, wi = i;
      if (i < kvol) {
        unsigned long long nk;
        if (MODE == 0) {
          const int ox = (i % k - centre) * dil, oy = ((i / k) % k - centre) * dil, oz = (i / (k * k) - centre) * dil;
          nk = tcn_pack(b, x + ox, y + oy, z + oz);
        } else {
          const int px = tcn_floor_to(x, 2 * dil), py = tcn_floor_to(y, 2 * dil), pz = tcn_floor_to(z, 2 * dil);
          wi = ((z - pz) / dil * 2 + (y - py) / dil) * 2 + (x - px) / dil;
          nk = tcn_pack(b, px, py, pz);
        }
        row = tcn_lookup(in_tab, nk);
      }
      unsigned found = __ballot_sync(0xffffffffu, row >= 0);
      while (found) {  // ascending offset order: the summation order is fixed
        const int src = __ffs(found) - 1;
        found &= found - 1;
        const int r = __shfl_sync(0xffffffffu, row, src);
        const float* __restrict__ Wi = W + (size_t)__shfl_sync(0xffffffffu, wi, src) * cin * cout;
        if (!in_feat) {  // conv0: the reference assigns a dummy feature 1 to every point (tcn.py:131-134)
          if (part == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (lane + 32 * j < cout) acc[j] += __ldg(Wi + lane + 32 * j);
          }
          continue;
        }
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          const float xv = (c0 + lane < cend) ? __ldg(in_feat + (size_t)r * cin + c0 + lane) : 0.f;
          const int cn = min(32, cend - c0);
          if (cn == 32) {
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {  // unrolled: eight rows of W in flight
              const float xs = __shfl_sync(0xffffffffu, xv, t);
              const float* __restrict__ Wr = Wi + (size_t)(c0 + t) * cout;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (lane + 32 * j < cout) acc[j] = fmaf(xs, __ldg(Wr + lane + 32 * j), acc[j]);
            }
          } else {
#pragma unroll 4
            for (int t = 0; t < cn; ++t) {
              const float xs = __shfl_sync(0xffffffffu, xv, t);
              const float* __restrict__ Wr = Wi + (size_t)(c0 + t) * cout;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (lane + 32 * j < cout) acc[j] = fmaf(xs, __ldg(Wr + lane + 32 * j), acc[j]);
            }
          }
        }
      }
    }
  }
  if (WPP > 1) {  // combine the channel slices in a fixed order
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[warp][lane + 32 * j] = acc[j];
    __syncthreads();
    if (part != 0) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      for (int q = 0; q < WPP; ++q) v += s_acc[warp + q][lane + 32 * j];
      acc[j] = v;
    }
  }
  if (!live) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    if (c < cout) {
      float v = acc[j];
      if (scale) v = v * scale[c] + shift[c];
      if (residual) v += residual[(size_t)p * cout + c];
      if (accumulate) v += out_feat[(size_t)p * cout + c];
      if (relu) v = fmaxf(v, 0.f);
      out_feat[(size_t)p * cout + c] = v;
    }
  }
}

// GeM pooling (minkloc.py:84-95) over the points of every batch element + optional L2
// normalisation (tcn.py:140-143) -> float64 (tcn.py:148).  Points of a batch element are contiguous
// (keys are batch-major).  Pass 1: grid (batch, TCN_GEM_SLICES), thread = channel, every block sums
// clamp(x)^p over its slice of the points; pass 2: one block per batch element combines the slices in
// a fixed order.
#define TCN_GEM_SLICES 64
__device__ __forceinline__ void tcn_batch_range(const unsigned long long* __restrict__ keys, int n, int b, int& lo_out, int& hi_out) {
  int r[2];
  for (int s = 0; s < 2; ++s) {  // first row with batch >= b + s
    const unsigned long long target = (unsigned long long)(b + s);
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((keys[mid] >> 54) < target) lo = mid + 1; else hi = mid;
    }
    r[s] = lo;
  }
  lo_out = r[0], hi_out = r[1];
}
__global__ void __launch_bounds__(256) k_tcn_gem_partial(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n,
                                                         const float* __restrict__ feat, int c, float p, float eps,
                                                         double* __restrict__ part /* batch x slices x c */) {
  int lo, hi;
  tcn_batch_range(keys, *d_n, blockIdx.x, lo, hi);
  const int per = (hi - lo + TCN_GEM_SLICES - 1) / TCN_GEM_SLICES;
  const int s0 = lo + blockIdx.y * per, s1 = min(s0 + per, hi);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double acc = 0.0;
    for (int r = s0; r < s1; ++r) acc += (double)powf(fmaxf(feat[(size_t)r * c + ch], eps), p);
    part[((size_t)blockIdx.x * TCN_GEM_SLICES + blockIdx.y) * c + ch] = acc;
  }
}
__global__ void __launch_bounds__(256) k_tcn_gem(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n,
                                                 const double* __restrict__ part, int c, float p, int normalize,
                                                 double* __restrict__ out) {
  __shared__ double s_sq[8];
  int lo, hi;
  tcn_batch_range(keys, *d_n, blockIdx.x, lo, hi);
  const int b = blockIdx.x;
  double acc_total = 0.0;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double acc = 0.0;
    for (int sl = 0; sl < TCN_GEM_SLICES; ++sl) acc += part[((size_t)b * TCN_GEM_SLICES + sl) * c + ch];
    const double g = (hi > lo) ? pow(acc / (double)(hi - lo), 1.0 / (double)p) : 0.0;
    out[(size_t)b * c + ch] = g;
    acc_total += g * g;
  }
  if (!normalize) return;
  acc_total = warp_sum(acc_total);
  if ((threadIdx.x & 31) == 0) s_sq[threadIdx.x >> 5] = acc_total;
  __syncthreads();
  double tot = 0.0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_sq[k];
  const double inv = 1.0 / fmax(sqrt(tot), 1e-12);  // F.normalize eps
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) out[(size_t)b * c + ch] *= inv;
}

// ------------------------------------------------------------------------- C ABI
extern "C" int mt_tcn_create(int device, int max_points, int max_batch, mt_tcn** out) {
  if (!out || max_points <= 0 || max_batch <= 0 || max_batch > 511) return set_err(MT_ERR_ARG, "mt_tcn_create: bad argument");
  CK(cudaSetDevice(device));
  mt_tcn* t = new mt_tcn();
  memset(t, 0, sizeof(*t));
  t->device = device, t->max_points = max_points, t->max_batch = max_batch;
  t->gem_p = 3.f, t->gem_eps = 1e-6f;
  unsigned cap = 1024;
  while (cap < 2u * (unsigned)max_points) cap <<= 1;
  t->cap = cap;
  for (int l = 0; l < 4; ++l) {
    CK(cudaMalloc(&t->tab[l].keys, sizeof(unsigned long long) * cap));
    CK(cudaMalloc(&t->tab[l].vals, sizeof(int) * cap));
    t->tab[l].mask = cap - 1;
    CK(cudaMalloc(&t->keys[l], sizeof(unsigned long long) * max_points));
  }
  CK(cudaMalloc(&t->d_n, sizeof(int) * 4));
  t->pool_floats = (size_t)max_points * 1184;
  CK(cudaMalloc(&t->pool, sizeof(float) * t->pool_floats));
  CK(cudaMalloc(&t->gem_part, sizeof(double) * (size_t)max_batch * 64 * 256));
  *out = t;
  return MT_OK;
}

extern "C" int mt_tcn_destroy(mt_tcn* t) {
  if (!t) return MT_OK;
  cudaSetDevice(t->device);
  for (int l = 0; l < 4; ++l) cudaFree(t->tab[l].keys), cudaFree(t->tab[l].vals), cudaFree(t->keys[l]);
  for (int i = 0; i < TCN_NCONV; ++i) cudaFree(t->conv[i].w);
  for (int i = 0; i < TCN_NBN; ++i) cudaFree(t->bn[i].scale), cudaFree(t->bn[i].shift);
  cudaFree(t->d_n), cudaFree(t->pool), cudaFree(t->gem_part);
  delete t;
  return MT_OK;
}

extern "C" int mt_tcn_set_conv(mt_tcn* t, int id, const float* h_kernel, int kvol, int cin, int cout) {
  if (!t || id < 0 || id >= TCN_NCONV || !h_kernel || kvol <= 0 || cin <= 0 || cout <= 0 || cout > 256)
    return set_err(MT_ERR_ARG, "mt_tcn_set_conv: bad argument (cout <= 256)");
  CK(cudaSetDevice(t->device));
  cudaFree(t->conv[id].w);
  t->conv[id].w = nullptr;
  const size_t n = (size_t)kvol * cin * cout;
  CK(cudaMalloc(&t->conv[id].w, sizeof(float) * n));
  CK(cudaMemcpy(t->conv[id].w, h_kernel, sizeof(float) * n, cudaMemcpyHostToDevice));
  t->conv[id].kvol = kvol, t->conv[id].cin = cin, t->conv[id].cout = cout;
  return MT_OK;
}

extern "C" int mt_tcn_set_bn(mt_tcn* t, int id, const float* w, const float* b, const float* mean, const float* var, int c,
                             float eps) {
  if (!t || id < 0 || id >= TCN_NBN || !w || !b || !mean || !var || c <= 0) return set_err(MT_ERR_ARG, "mt_tcn_set_bn: bad argument");
  CK(cudaSetDevice(t->device));
  std::vector<float> sc(c), sh(c);
  for (int i = 0; i < c; ++i) {  // float64 fold, then float32 like the stored parameters
    const double s = (double)w[i] / sqrt((double)var[i] + (double)eps);
    sc[i] = (float)s;
    sh[i] = (float)((double)b[i] - (double)mean[i] * s);
  }
  cudaFree(t->bn[id].scale), cudaFree(t->bn[id].shift);
  t->bn[id].scale = t->bn[id].shift = nullptr;
  CK(cudaMalloc(&t->bn[id].scale, sizeof(float) * c));
  CK(cudaMalloc(&t->bn[id].shift, sizeof(float) * c));
  CK(cudaMemcpy(t->bn[id].scale, sc.data(), sizeof(float) * c, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(t->bn[id].shift, sh.data(), sizeof(float) * c, cudaMemcpyHostToDevice));
  t->bn[id].c = c;
  return MT_OK;
}

extern "C" int mt_tcn_set_gem(mt_tcn* t, float p, float eps) {
  if (!t || !(p > 0.f)) return set_err(MT_ERR_ARG, "mt_tcn_set_gem: bad argument");
  t->gem_p = p, t->gem_eps = eps;
  return MT_OK;
}

static int tcn_conv_launch(mt_tcn* t, int mode, int conv_id, int bn_id, const unsigned long long* out_keys, const int* d_nout,
                           int nmax, const TcnTable& in_tab, const float* in_feat, int k, int dil, const float* residual, int relu,
                           int accumulate, float* out_feat, cudaStream_t st) {
  const TcnConv& c = t->conv[conv_id];
  if (!c.w) return set_err(MT_ERR_STATE, "mt_tcn_forward: a convolution has no weights (mt_tcn_set_conv)");
  if (mode == 0 && c.kvol != k * k * k) return set_err(MT_ERR_STATE, "mt_tcn_forward: kernel volume does not match the layer");
  const float* sc = nullptr;
  const float* sh = nullptr;
  if (bn_id >= 0) {
    if (!t->bn[bn_id].scale || t->bn[bn_id].c != c.cout) return set_err(MT_ERR_STATE, "mt_tcn_forward: BatchNorm missing / wrong width");
    sc = t->bn[bn_id].scale, sh = t->bn[bn_id].shift;
  }
  // few points and many channels: four warps per point split the input channels (shorter FMA chains);
  // conv0 (one input feature) keeps a warp per point
  const bool split = in_feat != nullptr && c.cin >= 32;
  const unsigned grid = (unsigned)(split ? (nmax + 1) / 2 : (nmax + 7) / 8);
  if (mode == 0 && split)
    k_tcn_conv<0, 4><<<grid, 256, 0, st>>>(out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else if (mode == 0)
    k_tcn_conv<0, 1><<<grid, 256, 0, st>>>(out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else if (split)
    k_tcn_conv<1, 4><<<grid, 256, 0, st>>>(out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else
    k_tcn_conv<1, 1><<<grid, 256, 0, st>>>(out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  CK_LAUNCH();
  return MT_OK;
}

// d_keys: n sorted unique packed coordinates (tcn_pack: batch | x | y | z) of the quantised clouds
// d_out: (batch, feature) float64.  d_counts (nullable): active points per level (4 ints, device).
extern "C" int mt_tcn_forward(mt_tcn* t, const unsigned long long* d_keys, int n, int batch, int normalize, double* d_out,
                              int* d_counts, void* stream) {
  if (!t || !d_keys || !d_out || n <= 0 || batch <= 0) return set_err(MT_ERR_ARG, "mt_tcn_forward: bad argument");
  if (n > t->max_points || batch > t->max_batch) return set_err(MT_ERR_CAPACITY, "mt_tcn_forward: n / batch exceed the context capacity");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned tg = (t->cap + 255) / 256, ng = (unsigned)((n + 255) / 256);
  for (int l = 0; l < 4; ++l) {
    k_tcn_clear<<<tg, 256, 0, st>>>(t->tab[l]);
    CK_LAUNCH();
  }
  CK(cudaMemcpyAsync(t->keys[0], d_keys, sizeof(unsigned long long) * n, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(t->d_n, &n, sizeof(int), cudaMemcpyHostToDevice, st));  // n is copied at enqueue time (pageable source)
  k_tcn_index<<<ng, 256, 0, st>>>(t->keys[0], t->d_n, t->tab[0]);
  CK_LAUNCH();
  for (int l = 1; l < 4; ++l) {  // coordinate maps at tensor stride 2, 4, 8
    const int stride2 = 1 << l;
    k_tcn_parent_first<<<ng, 256, 0, st>>>(t->keys[l - 1], t->d_n + l - 1, stride2, t->tab[l]);
    CK_LAUNCH();
    k_tcn_parent_rows<<<1, 1024, 0, st>>>(t->keys[l - 1], t->d_n + l - 1, stride2, t->tab[l], t->keys[l], t->d_n + l);
    CK_LAUNCH();
    k_tcn_untag<<<tg, 256, 0, st>>>(t->tab[l]);
    CK_LAUNCH();
  }
  // feature buffers carved from the pool (upper bound n rows each)
  const int c0 = t->conv[TCN_CONV0].cout, p0 = t->conv[TCN_BLK_C2(0)].cout, p1 = t->conv[TCN_BLK_C2(1)].cout,
            p2 = t->conv[TCN_BLK_C2(2)].cout, f = t->conv[TCN_LAT0].cout;
  if (!c0 || !p0 || !p1 || !p2 || !f) return set_err(MT_ERR_STATE, "mt_tcn_forward: weights not loaded");
  const int widths[] = {c0, c0, p0, p0, p0, p0, p1, p1, p1, p1, p2, p2, p2, f, f};
  size_t need = 0;
  for (int w : widths) need += (size_t)n * w;
  if (need > t->pool_floats) return set_err(MT_ERR_CAPACITY, "mt_tcn_forward: feature pool too small for these channel widths");
  float* ptr = t->pool;
  auto take = [&](int w) { float* r = ptr; ptr += (size_t)n * w; return r; };
  float* x0 = take(c0);
  int r;
#define TCN_RUN(call) if ((r = (call)) != MT_OK) return r
  // conv0 + bn0 + relu (minkfpn.py:113-115)
  TCN_RUN(tcn_conv_launch(t, 0, TCN_CONV0, TCN_BN0, t->keys[0], t->d_n, n, t->tab[0], nullptr, 5, 1, nullptr, 1, 0, x0, st));
  float* x = x0;
  float* fmap = nullptr;
  for (int s = 0; s < 3; ++s) {  // bottom-up: strided conv + bn + relu + BasicBlock (minkfpn.py:120-126)
    const int stride = 1 << s, l = s + 1, pl = t->conv[TCN_BLK_C2(s)].cout;
    float* xa = take(t->conv[TCN_DOWN(s)].cout);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_DOWN(s), TCN_BN_DOWN(s), t->keys[l], t->d_n + l, n, t->tab[l - 1], x, 2, stride, nullptr, 1, 0, xa, st));
    float* y = take(pl);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_C1(s), TCN_BN_N1(s), t->keys[l], t->d_n + l, n, t->tab[l], xa, 3, 2 * stride, nullptr, 1, 0, y, st));
    const float* res = xa;
    if (t->conv[TCN_BLK_DS(s)].w) {  // channel change: residual = bn(conv1x1(x)) (resnet.py:89-101)
      float* rs = take(pl);
      TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_DS(s), TCN_BN_DS(s), t->keys[l], t->d_n + l, n, t->tab[l], xa, 1, 2 * stride, nullptr, 0, 0, rs, st));
      res = rs;
    } else if (t->conv[TCN_DOWN(s)].cout != pl) {
      return set_err(MT_ERR_STATE, "mt_tcn_forward: block changes width but has no downsample branch");
    }
    float* xo = take(pl);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_C2(s), TCN_BN_N2(s), t->keys[l], t->d_n + l, n, t->tab[l], y, 3, 2 * stride, res, 1, 0, xo, st));
    x = xo;
    if (s == 1) fmap = xo;
  }
  // lateral 1x1 at the top, transposed conv down to stride 4, + lateral 1x1 of the stage-1 map (minkfpn.py:130-136)
  float* z = take(f);
  TCN_RUN(tcn_conv_launch(t, 0, TCN_LAT0, -1, t->keys[3], t->d_n + 3, n, t->tab[3], x, 1, 8, nullptr, 0, 0, z, st));
  float* fp = take(f);
  TCN_RUN(tcn_conv_launch(t, 0, TCN_LAT1, -1, t->keys[2], t->d_n + 2, n, t->tab[2], fmap, 1, 4, nullptr, 0, 0, fp, st));
  TCN_RUN(tcn_conv_launch(t, 1, TCN_TCONV, -1, t->keys[2], t->d_n + 2, n, t->tab[3], z, 2, 4, nullptr, 0, 1, fp, st));
#undef TCN_RUN
  k_tcn_gem_partial<<<dim3(batch, TCN_GEM_SLICES), 256, 0, st>>>(t->keys[2], t->d_n + 2, fp, f, t->gem_p, t->gem_eps, t->gem_part);
  CK_LAUNCH();
  k_tcn_gem<<<batch, 256, 0, st>>>(t->keys[2], t->d_n + 2, t->gem_part, f, t->gem_p, normalize, d_out);
  CK_LAUNCH();
  if (d_counts) CK(cudaMemcpyAsync(d_counts, t->d_n, sizeof(int) * 4, cudaMemcpyDeviceToDevice, st));
  return MT_OK;
}
