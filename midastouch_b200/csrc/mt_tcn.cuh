// Tactile code network forward pass (TCN = MinkLoc3D: sparse 3-D FPN + GeM) on sm_100a.
//
// Replaces the MinkowskiEngine calls of contrib/tcn_minkloc/{tcn,minkloc,minkfpn}.py for the
// shipped configuration (config/tcn/default.yaml: planes 32,64,64, layers 1,1,1, one top-down
// block, conv0 kernel 5, 256-d output).  Sparse tensors are (unique 64-bit coordinate keys in order of
// first occurrence, row-major float32 features); a coordinate map is an open-addressing hash table
// key -> row.  Per forward pass:
//   1. k_tcn_insert_all / k_tcn_count_all / k_tcn_scatter_all: the raw points (float clouds, quantised
//      in the kernel, or packed keys) register with the coordinate maps of all four levels at once;
//      the rows of a level are its distinct keys in order of their first raw point (deterministic,
//      batch-major), found by an ordered multi-block compaction -- this replaces torch.unique (a
//      64-bit radix sort) and three single-block passes.
//   2. k_tcn_kmaps: the kernel maps (row of every kernel offset, -1 = absent, + a bit mask of the
//      offsets present) of the 3x3x3, the strided 2x2x2 and the transposed 2x2x2 convolutions: the
//      hash tables are probed once per (point, offset), not once per layer.  The same kernel groups the
//      (output point, kernel offset) pairs that exist by offset (pair lists).
//   3. k_tcn_pair_mma + k_tcn_pair_reduce: every gathering convolution as a gather-GEMM on the tensor
//      cores with work proportional to the pairs present: a work item is up to 32 pairs of one offset x 16
//      output channels, operands straight from L1/L2 into mma.sync.m16n8k8 TF32 fragments (16-byte loads,
//      channels permuted onto the k slots), every product as three MMAs on the (big, small) TF32 split
//      of both operands (3xTF32: float32-grade accuracy); the product rows of a point are added in
//      ascending offset order (fixed summation order, no float atomics) with BatchNorm(eval) / residual /
//      accumulate / ReLU in that pass.  1x1 layers use the direct form k_tcn_conv_mma (one warp = 16
//      output points x 8 NT channels).  The weights of an offset are read once per 32 pairs instead of
//      once per point (the CUDA-core kernel was bound by exactly that L1 traffic: 1 GB for the 256 -> 256
//      transposed convolution).
//   4. every kernel is launched with programmatic dependent launch (tcn_pdl / tcn_launch): the ~30 short,
//      strictly dependent kernels overlap their launch latencies.
//   k_tcn_conv0 serves conv0 (one dummy input feature: a sum of kernel rows); k_tcn_conv (one warp per
//   output point, hash look-ups, CUDA cores) remains for channel widths the MMA tiling does not divide.
// Semantics restated in oracle/tcn_oracle.py (parity unpinned against MinkowskiEngine itself).
//
// Included at the end of midas_b200.cu (same translation unit: shares set_err / CK).
#pragma once

#define TCN_EMPTY 0xFFFFFFFFFFFFFFFFull
#define TCN_NCONV 16
#define TCN_CTL_PAIRS 8
#define TCN_CTL_BLO 256
#define TCN_CTL_BHI 768
#define TCN_CTL_INTS 1280
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
  float* w;   // (kvol, cin, cout)
  float* wt;  // (kvol, cout, cin): the B fragments of the MMA kernels are 16-byte loads along cin
  int kvol, cin, cout;
};
struct TcnBn {
  float* scale;  // gamma / sqrt(var + eps)
  float* shift;  // beta - mean * scale
  int c;
};
struct TcnTable {
  unsigned long long* keys;
  int* vals;      // smallest raw point index that maps to the key (its deterministic representative)
  int* rows;      // row of the key in the level's feature matrix
  unsigned mask;  // capacity - 1 (power of two)
};
struct TcnTabs {
  TcnTable t[4];
};

struct mt_tcn {
  int device, max_points, max_batch;
  TcnConv conv[TCN_NCONV];
  TcnBn bn[TCN_NBN];
  float gem_p, gem_eps;
  unsigned cap;            // hash capacity per level
  TcnTable tab[4];         // level 0..3 (tensor stride 1,2,4,8)
  unsigned long long* keys[4];
  int* d_n;                // control block (device): [0..3] active points per level | [4] status flag (coordinate out of
                           // range) | [8 + 32 m + i] pairs of kernel offset i in map m, [8 + 32 m + 31] pairs of map m
                           // | [TCN_CTL_BLO + b], [TCN_CTL_BHI + b] level-2 row range of batch element b
  int2* plist[7];          // map m (0..2: 3x3x3 at level 1..3, 3..5: children of level 1..3, 6: parent of level 2): kvol x
                           // max_points (input row, partial-sum slot) pairs, grouped by kernel offset
  int* pbase[7];           // first partial-sum slot of every output point (its pairs follow in ascending offset order)
  float* partial;          // partial sums of the pair GEMM: one row of cout floats per pair
  size_t partial_floats;
  int* slot_of;            // 4 x max_points: hash slot of raw point i at level l
  int* blk_cnt;            // 4 x ceil(max_points / TCN_CBLOCK): first-point counts per block (ordered compaction)
  int* kmap3[4];           // level 1..3: n x 27 rows of the 3x3x3 neighbourhood (dilation = level stride)
  int* kmap2[4];           // level 1..3: n x 8 rows of the children at level l - 1
  int* kmapt;              // level 2: n x 8, the parent's row at level 3 in column "position inside the parent"
  unsigned* kmask3[4];     // bit i set: offset i present
  unsigned* kmask2[4];
  unsigned* kmaskt;
  int gem_slices;
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
    if (k == key) return t.rows[s];
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


// Programmatic dependent launch: the 31 kernels of a forward pass are short and strictly dependent, so the gap between
// them (~1.5 us each) is a fifth of the pass.  Every kernel opens with tcn_pdl(): it lets the NEXT kernel's blocks be
// scheduled right away (they park in their own griddepcontrol.wait) and then waits until the PREVIOUS kernel has
// completed and its writes are visible -- the data dependence is unchanged, only the launch latency is overlapped.
__device__ __forceinline__ void tcn_pdl() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static void tcn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
  static const bool no_pdl = getenv("MIDAS_B200_TCN_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);  // (errors surface in the CK_LAUNCH() that follows every launch)
}

__global__ void k_tcn_clear_all(TcnTabs T, int* d_n) {
  tcn_pdl();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < TCN_CTL_BLO) d_n[i] = 0;
  if (i < 512) d_n[TCN_CTL_BLO + i] = INT_MAX, d_n[TCN_CTL_BHI + i] = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l)
    if (i <= T.t[l].mask) T.t[l].keys[i] = TCN_EMPTY, T.t[l].vals[i] = INT_MAX;
}

// Every raw point registers with the coordinate maps of all four levels (tensor stride 1, 2, 4, 8: the key of level l
// is the level-0 key floored to 2^l, which equals flooring level by level); a key remembers its first raw point.
// pts != NULL: (n_raw, 3) float32 clouds of P points each, voxel = floor(x * inv_q) -- torch's `cloud / q` on CUDA is a
// multiplication by a float32 reciprocal (tcn.py:124-130), reproduced bit for bit; keys_in otherwise (packed).
__global__ void __launch_bounds__(256) k_tcn_insert_all(const float* __restrict__ pts, const unsigned long long* __restrict__ keys_in,
                                                        int n_raw, int P, float inv_q, TcnTabs T, int* __restrict__ slot_of,
                                                        int stride, int* __restrict__ d_flag) {
  tcn_pdl();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid >> 2, l = tid & 3;  // four threads per raw point, one per level: four independent atomic chains
  if (i >= n_raw) return;
  int b, x, y, z;
  if (keys_in) {
    tcn_unpack(keys_in[i], b, x, y, z);
  } else {
    const float fx = floorf(__fmul_rn(pts[3 * (size_t)i], inv_q)), fy = floorf(__fmul_rn(pts[3 * (size_t)i + 1], inv_q)),
                fz = floorf(__fmul_rn(pts[3 * (size_t)i + 2], inv_q));
    const float lim = 131071.f;  // 18-bit coordinates
    const bool ok = fabsf(fx) <= lim && fabsf(fy) <= lim && fabsf(fz) <= lim;  // (false for NaN too)
    if (!ok) *d_flag = 1;  // the descriptors of this call are poisoned with NaN (k_tcn_gem)
    b = i / P;
    x = ok ? (int)fx : 0, y = ok ? (int)fy : 0, z = ok ? (int)fz : 0;
  }
  const int m = 1 << l;
  const TcnTable tl = l == 0 ? T.t[0] : (l == 1 ? T.t[1] : (l == 2 ? T.t[2] : T.t[3]));
  const unsigned s = tcn_insert(tl, tcn_pack(b, tcn_floor_to(x, m), tcn_floor_to(y, m), tcn_floor_to(z, m)));
  atomicMin(tl.vals + s, i);
  slot_of[(size_t)l * stride + i] = (int)s;
}

// ordered compaction, pass 1: first points (the representative of their key) per block of TCN_CBLOCK raw points and level.
// (Blocks of 256, not 1024: one frame is 4096 raw points, and four blocks of 1024 threads put 4096 dependent look-ups
// per level on each of only four SMs -- more than their load queues hold: 10.5 us for the second pass.)
#define TCN_CBLOCK 256
__global__ void __launch_bounds__(TCN_CBLOCK) k_tcn_count_all(int n_raw, TcnTabs T, const int* __restrict__ slot_of, int stride,
                                                        int* __restrict__ blk_cnt, int nblk) {
  tcn_pdl();
  __shared__ int s_w[4][TCN_CBLOCK / 32];
  const int i = blockIdx.x * TCN_CBLOCK + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const bool f = i < n_raw && T.t[l].vals[slot_of[(size_t)l * stride + i]] == i;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[l][w] = __popc(bal);
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int c = 0;
    for (int k = 0; k < TCN_CBLOCK / 32; ++k) c += s_w[threadIdx.x][k];
    blk_cnt[threadIdx.x * nblk + blockIdx.x] = c;
  }
}
// pass 2: row = number of first points before this one; writes the level's keys, the key's row and the level counts
__global__ void __launch_bounds__(TCN_CBLOCK) k_tcn_scatter_all(int n_raw, TcnTabs T, const int* __restrict__ slot_of, int stride,
                                                          const int* __restrict__ blk_cnt, int nblk,
                                                          unsigned long long* __restrict__ k0, unsigned long long* __restrict__ k1,
                                                          unsigned long long* __restrict__ k2, unsigned long long* __restrict__ k3,
                                                          int* __restrict__ d_n) {
  tcn_pdl();
  __shared__ int s_w[4][TCN_CBLOCK / 32];
  __shared__ int s_carry[4];
  const int i = blockIdx.x * TCN_CBLOCK + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  static_assert(TCN_CBLOCK >= 128, "four warps compute the carries");
  if (w < 4) {  // warp l: first points of level l in the blocks before this one
    int c = 0;
    for (int bq = lane; bq < (int)blockIdx.x; bq += 32) c += blk_cnt[w * nblk + bq];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_carry[w] = c;
  }
  bool f[4];
  unsigned below[4];
  int slot[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    slot[l] = i < n_raw ? slot_of[(size_t)l * stride + i] : 0;
    f[l] = i < n_raw && T.t[l].vals[slot[l]] == i;
    const unsigned bal = __ballot_sync(0xffffffffu, f[l]);
    below[l] = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_w[l][w] = __popc(bal);
  }
  __syncthreads();
  unsigned long long* const kk[4] = {k0, k1, k2, k3};
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    int base = s_carry[l];
    for (int k = 0; k < w; ++k) base += s_w[l][k];
    if (f[l]) {
      const int row = base + (int)below[l];
      const unsigned long long key = T.t[l].keys[slot[l]];
      T.t[l].rows[slot[l]] = row;
      kk[l][row] = key;
      if (l == 2) {  // rows are batch-major: the GeM pooling reads its row range from here
        const int bq = (int)(key >> 54);
        atomicMin(d_n + TCN_CTL_BLO + bq, row), atomicMax(d_n + TCN_CTL_BHI + bq, row + 1);
      }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == TCN_CBLOCK - 1) d_n[l] = base + s_w[l][TCN_CBLOCK / 32 - 1];
  }
}

// Kernel maps, one warp per (level 1..3, point): rows of the 27 neighbours at the level's own dilation (BasicBlock
// convolutions), of the 8 children one level below (strided 2x2x2 convolution, offset index x fastest like
// MinkowskiEngine's kernel layout) and -- level 2 only -- of the parent at level 3 in the column of the point's
// position inside it (transposed 2x2x2 convolution).  -1 = absent; the masks carry one bit per offset present.
struct TcnMaps {  // index: level 1..3 (0 unused)
  int* m3[4];
  int* m2[4];
  int* mt;
  unsigned* q3[4];
  unsigned* q2[4];
  unsigned* qt;
  int2* pl[7];
  int* pb[7];
  int ncap;
};
// one warp, one output point (8 per block): append its (input row, partial-sum slot) pairs to the lists of their kernel
// offsets.  The block aggregates: one atomic per (block, offset) and one for the block's partial-sum slots -- the centre
// offset is present at every point, and one atomic per point on its counter was most of this kernel's time.
__device__ __forceinline__ void tcn_pairs(int* __restrict__ ctr, int2* __restrict__ pl, int* __restrict__ pb, int ncap, int p, int row,
                                          int lane, int w, bool valid, int (*s_hit)[32], int* s_np, int* s_gbase, int* s_sbase) {
  const unsigned bal = __ballot_sync(0xffffffffu, row >= 0);
  s_hit[w][lane] = row >= 0;
  if (lane == 0) s_np[w] = __popc(bal);
  __syncthreads();
  int rank = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int hk = s_hit[k][lane];
    rank += (k < w) ? hk : 0;
    tot += hk;
  }
  if (w == 0) {
    if (tot) s_gbase[lane] = atomicAdd(ctr + lane, tot);
    if (lane == 0) {
      int np = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) np += s_np[k];
      *s_sbase = np ? atomicAdd(ctr + 31, np) : 0;
    }
  }
  __syncthreads();
  int base = *s_sbase;
  for (int k = 0; k < w; ++k) base += s_np[k];
  if (valid && lane == 0) pb[p] = base;
  if (row >= 0) pl[(size_t)lane * ncap + s_gbase[lane] + rank] = make_int2(row, base + __popc(bal & ((1u << lane) - 1)));
  __syncthreads();  // the shared arrays are reused by the next map
}
__global__ void __launch_bounds__(256) k_tcn_kmaps(TcnTabs T, const unsigned long long* __restrict__ k1, const unsigned long long* __restrict__ k2,
                                                   const unsigned long long* __restrict__ k3, int* __restrict__ d_n, TcnMaps M) {
  tcn_pdl();
  __shared__ int s_hit[8][32];
  __shared__ int s_np[8], s_gbase[32], s_sbase;
  const int l = blockIdx.y + 1, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = blockIdx.x * 8 + w;
  if (blockIdx.x * 8 >= d_n[l]) return;  // (block-uniform)
  const bool valid = p < d_n[l];
  const unsigned long long* keys = l == 1 ? k1 : (l == 2 ? k2 : k3);
  int* m3 = l == 1 ? M.m3[1] : (l == 2 ? M.m3[2] : M.m3[3]);
  int* m2 = l == 1 ? M.m2[1] : (l == 2 ? M.m2[2] : M.m2[3]);
  unsigned* q3 = l == 1 ? M.q3[1] : (l == 2 ? M.q3[2] : M.q3[3]);
  unsigned* q2 = l == 1 ? M.q2[1] : (l == 2 ? M.q2[2] : M.q2[3]);
  int2* pl3 = l == 1 ? M.pl[0] : (l == 2 ? M.pl[1] : M.pl[2]);
  int2* pl2 = l == 1 ? M.pl[3] : (l == 2 ? M.pl[4] : M.pl[5]);
  int* pb3 = l == 1 ? M.pb[0] : (l == 2 ? M.pb[1] : M.pb[2]);
  int* pb2 = l == 1 ? M.pb[3] : (l == 2 ? M.pb[4] : M.pb[5]);
  const TcnTable tl = l == 1 ? T.t[1] : (l == 2 ? T.t[2] : T.t[3]), tc = l == 1 ? T.t[0] : (l == 2 ? T.t[1] : T.t[2]);
  int b = 0, x = 0, y = 0, z = 0;
  if (valid) tcn_unpack(keys[p], b, x, y, z);
  const int s = 1 << l, h = s >> 1;
  // the three groups of look-ups are independent: all of them are requested before any list is written
  int row3 = -1, row2 = -1, rowt = -1, wi = 0;
  if (valid && lane < 27) row3 = tcn_lookup(tl, tcn_pack(b, x + (lane % 3 - 1) * s, y + ((lane / 3) % 3 - 1) * s, z + (lane / 9 - 1) * s));
  if (valid && lane < 8) row2 = tcn_lookup(tc, tcn_pack(b, x + (lane & 1) * h, y + ((lane >> 1) & 1) * h, z + (lane >> 2) * h));
  if (l == 2 && valid) {
    const int px = tcn_floor_to(x, 8), py = tcn_floor_to(y, 8), pz = tcn_floor_to(z, 8);
    wi = ((z - pz) / 4 * 2 + (y - py) / 4) * 2 + (x - px) / 4;
    if (lane == wi) rowt = tcn_lookup(T.t[3], tcn_pack(b, px, py, pz));
  }
  if (valid) {
    if (lane < 27) m3[(size_t)p * 27 + lane] = row3;
    if (lane < 8) m2[(size_t)p * 8 + lane] = row2;
  }
  unsigned bal = __ballot_sync(0xffffffffu, row3 >= 0);
  if (valid && lane == 0) q3[p] = bal;
  tcn_pairs(d_n + TCN_CTL_PAIRS + 32 * (l - 1), pl3, pb3, M.ncap, p, row3, lane, w, valid, s_hit, s_np, s_gbase, &s_sbase);
  bal = __ballot_sync(0xffffffffu, row2 >= 0);
  if (valid && lane == 0) q2[p] = bal;
  tcn_pairs(d_n + TCN_CTL_PAIRS + 32 * (l + 2), pl2, pb2, M.ncap, p, row2, lane, w, valid, s_hit, s_np, s_gbase, &s_sbase);
  if (l == 2) {
    bal = __ballot_sync(0xffffffffu, rowt >= 0);
    if (valid) {
      if (lane < 8) M.mt[(size_t)p * 8 + lane] = rowt;
      if (lane == 0) M.qt[p] = bal;
    }
    tcn_pairs(d_n + TCN_CTL_PAIRS + 32 * 6, M.pl[6], M.pb[6], M.ncap, p, rowt, lane, w, valid, s_hit, s_np, s_gbase, &s_sbase);
  }
}

// Gather-GEMM convolution on the tensor cores (see the head of this file).  map: (n_out, kvol) input rows or NULL (1x1: the
// point itself); hitmask: (n_out) offsets present or NULL.  One warp: output points [16 mt, 16 mt + 16) x channels
// [8 NT slab, 8 NT (slab + 1)); fragments of mma.sync.m16n8k8 (row.col, TF32 in, float32 accumulate):
//   A (16 x 8)  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)       g = lane / 4, t = lane % 4
//   B (8 x 8)   b0 (t, g)  b1 (t + 4, g)
//   C (16 x 8)  c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1)
__device__ __forceinline__ void tcn_split_tf32(float x, unsigned& big, unsigned& small) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(big) : "f"(x));
  const float r = x - __uint_as_float(big);  // exact
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(small) : "f"(r));
}
__device__ __forceinline__ void tcn_mma_tf32(float c[4], const unsigned a[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// K is consumed 16 channels at a time with one 16-byte load per operand row: lane (g, t) holds channels 4t .. 4t+3 of
// its rows, and the two MMAs of the chunk take channels (4t, 4t+1) and (4t+2, 4t+3) as their k slots (t, t+4) -- the
// assignment of channels to k slots is free as long as A and B agree.  Wt is the (kvol, cout, cin) copy of the weights.
__device__ __forceinline__ float tcn_f4(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
template <int NT>
__global__ void __launch_bounds__(128) k_tcn_conv_mma(const int* __restrict__ map, const unsigned* __restrict__ hitmask, int kvol,
                                                      const int* __restrict__ d_nout, const float* __restrict__ in_feat, int cin,
                                                      const float* __restrict__ Wt, int cout, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, const float* __restrict__ residual, int relu,
                                                      int accumulate, float* __restrict__ out_feat) {
  tcn_pdl();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nslab = cout / (8 * NT);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int p0 = (wid / nslab) * 16, n0 = (wid % nslab) * 8 * NT;
  const int n = *d_nout;
  if (p0 >= n) return;
  const int plo = p0 + g, phi = p0 + g + 8;
  const bool vlo = plo < n, vhi = phi < n;
  unsigned m = 1u;
  if (hitmask) m = __reduce_or_sync(0xffffffffu, (lane < 16 && p0 + lane < n) ? hitmask[p0 + lane] : 0u);
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  while (m) {  // ascending offsets: the summation order of every output is fixed
    const int i = __ffs(m) - 1;
    m &= m - 1;
    const int rlo = vlo ? (map ? map[(size_t)plo * kvol + i] : plo) : -1;
    const int rhi = vhi ? (map ? map[(size_t)phi * kvol + i] : phi) : -1;
    const float* __restrict__ xlo = in_feat + (size_t)(rlo < 0 ? 0 : rlo) * cin + 4 * t;
    const float* __restrict__ xhi = in_feat + (size_t)(rhi < 0 ? 0 : rhi) * cin + 4 * t;
    const float* __restrict__ Wi = Wt + ((size_t)i * cout + n0 + g) * cin + 4 * t;
#pragma unroll 2
    for (int k0 = 0; k0 < cin; k0 += 16) {
      const float4 alo = rlo >= 0 ? __ldg((const float4*)(xlo + k0)) : zero4;
      const float4 ahi = rhi >= 0 ? __ldg((const float4*)(xhi + k0)) : zero4;
      float4 bv[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bv[nt] = __ldg((const float4*)(Wi + (size_t)8 * nt * cin + k0));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned ab[4], as[4];
        tcn_split_tf32(tcn_f4(alo, 2 * h), ab[0], as[0]);
        tcn_split_tf32(tcn_f4(ahi, 2 * h), ab[1], as[1]);
        tcn_split_tf32(tcn_f4(alo, 2 * h + 1), ab[2], as[2]);
        tcn_split_tf32(tcn_f4(ahi, 2 * h + 1), ab[3], as[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          unsigned bb0, bs0, bb1, bs1;
          tcn_split_tf32(tcn_f4(bv[nt], 2 * h), bb0, bs0);
          tcn_split_tf32(tcn_f4(bv[nt], 2 * h + 1), bb1, bs1);
          tcn_mma_tf32(acc[nt], as, bb0, bb1);  // small terms first
          tcn_mma_tf32(acc[nt], ab, bs0, bs1);
          tcn_mma_tf32(acc[nt], ab, bb0, bb1);
        }
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = n0 + 8 * nt + 2 * t;
    float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
    if (scale) sc0 = scale[c], sc1 = scale[c + 1], sh0 = shift[c], sh1 = shift[c + 1];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int p = hh ? phi : plo;
      if (!(hh ? vhi : vlo)) continue;
      float v0 = acc[nt][2 * hh], v1 = acc[nt][2 * hh + 1];
      if (scale) v0 = v0 * sc0 + sh0, v1 = v1 * sc1 + sh1;
      float2* o = (float2*)(out_feat + (size_t)p * cout + c);
      if (residual) {
        const float2 r = *(const float2*)(residual + (size_t)p * cout + c);
        v0 += r.x, v1 += r.y;
      }
      if (accumulate) {
        const float2 r = *o;
        v0 += r.x, v1 += r.y;
      }
      if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
      *o = make_float2(v0, v1);
    }
  }
}

// conv0 (minkfpn.py:113-115): the reference gives every point the dummy feature 1 (tcn.py:131-134), so the layer is the
// sum of the kernel rows of the offsets present + BatchNorm + ReLU.  One warp per point; the lanes probe the hash map
// for all k^3 <= 128 offsets first (independent chains, one round trip), then add the rows in ascending offset order.
__global__ void __launch_bounds__(256) k_tcn_conv0(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n, TcnTable tab,
                                                   const float* __restrict__ W, int cout, int k, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, float* __restrict__ out_feat) {
  tcn_pdl();
  const int lane = threadIdx.x & 31, p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= *d_n) return;
  int b, x, y, z;
  tcn_unpack(keys[p], b, x, y, z);
  const int kvol = k * k * k, centre = k >> 1;
  bool hit[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = 32 * r + lane;
    hit[r] = i < kvol && tcn_lookup(tab, tcn_pack(b, x + i % k - centre, y + (i / k) % k - centre, z + i / (k * k) - centre)) >= 0;
  }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    unsigned m = __ballot_sync(0xffffffffu, hit[r]);
    while (m) {
      const int i = 32 * r + __ffs(m) - 1;
      m &= m - 1;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < cout) acc[j] += __ldg(W + (size_t)i * cout + lane + 32 * j);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    if (c < cout) out_feat[(size_t)p * cout + c] = fmaxf(scale ? acc[j] * scale[c] + shift[c] : acc[j], 0.f);
  }
}

// Pair form of the gather-GEMM: work proportional to the (output point, kernel offset) pairs that exist.  k_tcn_kmaps
// grouped the pairs by kernel offset; a work item is up to 32 pairs of ONE offset (two m16 tiles sharing their B
// fragments) x 8 NT output channels, so no MMA row is spent on an absent neighbour and the weights of an offset are
// read once per 32 pairs.  The product rows go to `partial` (one row per pair; a point's rows are consecutive, in
// ascending offset order) and k_tcn_pair_reduce adds them up per point in that order -- the result does not depend
// on how the atomics of k_tcn_kmaps ordered the lists -- and applies BatchNorm / residual / accumulate / ReLU.
// Persistent grid: every warp derives the tile table from the kvol pair counters and strides over the items.
template <int NT>
__global__ void __launch_bounds__(128) k_tcn_pair_mma(const int2* __restrict__ pl, const int* __restrict__ ctr, int kvol, int ncap,
                                                      const float* __restrict__ in_feat, int cin, const float* __restrict__ Wt, int cout,
                                                      float* __restrict__ partial) {
  tcn_pdl();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nslab = cout / (8 * NT);
  const int cnt_i = lane < kvol ? ctr[lane] : 0;
  const int tiles_i = (cnt_i + 31) >> 5;
  int incl = tiles_i;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31) * nslab;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
    const int tile = item / nslab, n0 = (item % nslab) * 8 * NT;
    const int i = __ffs(__ballot_sync(0xffffffffu, incl > tile)) - 1;  // the offset this tile belongs to
    const int first = __shfl_sync(0xffffffffu, incl - tiles_i, i), c_i = __shfl_sync(0xffffffffu, cnt_i, i);
    const int pair0 = (tile - first) * 32;
    const bool two = c_i - pair0 > 16;  // second m16 tile in use
    int2 e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = pair0 + g + 8 * q;
      e[q] = idx < c_i ? pl[(size_t)i * ncap + idx] : make_int2(-1, -1);
    }
    const float* __restrict__ xr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) xr[q] = in_feat + (size_t)(e[q].x < 0 ? 0 : e[q].x) * cin + 4 * t;
    const float* __restrict__ Wi = Wt + ((size_t)i * cout + n0 + g) * cin + 4 * t;
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int k0 = 0; k0 < cin; k0 += 16) {  // (channel -> k slot assignment: see k_tcn_conv_mma)
      float4 av[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) av[q] = (e[q].x >= 0) ? __ldg((const float4*)(xr[q] + k0)) : zero4;  // (rows 2, 3 absent unless `two`)
      float4 bv[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bv[nt] = __ldg((const float4*)(Wi + (size_t)8 * nt * cin + k0));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned ab[2][4], as[2][4];
        tcn_split_tf32(tcn_f4(av[0], 2 * h), ab[0][0], as[0][0]);
        tcn_split_tf32(tcn_f4(av[1], 2 * h), ab[0][1], as[0][1]);
        tcn_split_tf32(tcn_f4(av[0], 2 * h + 1), ab[0][2], as[0][2]);
        tcn_split_tf32(tcn_f4(av[1], 2 * h + 1), ab[0][3], as[0][3]);
        if (two) {
          tcn_split_tf32(tcn_f4(av[2], 2 * h), ab[1][0], as[1][0]);
          tcn_split_tf32(tcn_f4(av[3], 2 * h), ab[1][1], as[1][1]);
          tcn_split_tf32(tcn_f4(av[2], 2 * h + 1), ab[1][2], as[1][2]);
          tcn_split_tf32(tcn_f4(av[3], 2 * h + 1), ab[1][3], as[1][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          unsigned bb0, bs0, bb1, bs1;
          tcn_split_tf32(tcn_f4(bv[nt], 2 * h), bb0, bs0);
          tcn_split_tf32(tcn_f4(bv[nt], 2 * h + 1), bb1, bs1);
          tcn_mma_tf32(acc[0][nt], as[0], bb0, bb1);  // small terms first
          tcn_mma_tf32(acc[0][nt], ab[0], bs0, bs1);
          tcn_mma_tf32(acc[0][nt], ab[0], bb0, bb1);
          if (two) {
            tcn_mma_tf32(acc[1][nt], as[1], bb0, bb1);
            tcn_mma_tf32(acc[1][nt], ab[1], bs0, bs1);
            tcn_mma_tf32(acc[1][nt], ab[1], bb0, bb1);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // q = 2 mt + half: rows g (+8) of m16 tile mt
      if (e[q].y < 0) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        *(float2*)(partial + (size_t)e[q].y * cout + n0 + 8 * nt + 2 * t) =
            make_float2(acc[q >> 1][nt][2 * (q & 1)], acc[q >> 1][nt][2 * (q & 1) + 1]);
    }
  }
}
// out[p] = epilogue(sum of the point's partial rows, ascending kernel offset); one thread = four channels
__global__ void __launch_bounds__(256) k_tcn_pair_reduce(const int* __restrict__ pbase, const unsigned* __restrict__ hitmask,
                                                         const int* __restrict__ d_nout, const float* __restrict__ partial, int cout,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ residual, int relu, int accumulate,
                                                         float* __restrict__ out_feat) {
  tcn_pdl();
  const int q4 = cout >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = (int)(idx / q4), c = (int)(idx % q4) * 4;
  if (p >= *d_nout) return;
  const int base = pbase[p], cnt = __popc(hitmask[p]);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < cnt; ++j) {
    const float4 a = *(const float4*)(partial + (size_t)(base + j) * cout + c);
    v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
  }
  if (scale) {
    const float4 sc = *(const float4*)(scale + c), sh = *(const float4*)(shift + c);
    v.x = v.x * sc.x + sh.x, v.y = v.y * sc.y + sh.y, v.z = v.z * sc.z + sh.z, v.w = v.w * sc.w + sh.w;
  }
  float4* o = (float4*)(out_feat + (size_t)p * cout + c);
  if (residual) {
    const float4 r = *(const float4*)(residual + (size_t)p * cout + c);
    v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
  }
  if (accumulate) {
    const float4 r = *o;
    v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
  }
  if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
  *o = v;
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
  tcn_pdl();
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
// (rows are batch-major).  Pass 1: grid (batch, slices), thread = channel, every block sums
// clamp(x)^p over its slice of the points; pass 2: one block per batch element combines the slices in
// a fixed order.  slices = 256 for small batches (16 points per thread at 4096 points), 64 otherwise.
#define TCN_GEM_SLICES 64
__global__ void __launch_bounds__(256) k_tcn_gem_partial(const int* __restrict__ ctl, const float* __restrict__ feat, int c, float p,
                                                         float eps, double* __restrict__ part /* batch x slices x c */) {
  tcn_pdl();
  int lo = ctl[TCN_CTL_BLO + blockIdx.x];
  const int hi = ctl[TCN_CTL_BHI + blockIdx.x];
  if (hi == 0) lo = 0;  // no point in this batch element
  const int slices = gridDim.y;
  const int per = (hi - lo + slices - 1) / slices;
  const int s0 = lo + blockIdx.y * per, s1 = min(s0 + per, hi);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double acc = 0.0;
#pragma unroll 4
    for (int r = s0; r < s1; ++r)  // x^p = 2^(p log2 x), x >= eps > 0: two MUFU operations (~1e-6 relative on the terms that matter)
      acc += (double)exp2f(p * __log2f(fmaxf(__ldg(feat + (size_t)r * c + ch), eps)));
    part[((size_t)blockIdx.x * slices + blockIdx.y) * c + ch] = acc;
  }
}
__global__ void __launch_bounds__(1024) k_tcn_gem(const int* __restrict__ ctl, const double* __restrict__ part, int slices, int c, float p,
                                                  int normalize, double* __restrict__ out) {
  tcn_pdl();
  __shared__ double s_grp[4][256];
  __shared__ double s_sq[8];
  int lo = ctl[TCN_CTL_BLO + blockIdx.x];
  const int hi = ctl[TCN_CTL_BHI + blockIdx.x];
  if (hi == 0) lo = 0;
  const int b = blockIdx.x;
  const bool poisoned = ctl[4] != 0;  // a voxel coordinate did not fit the 18-bit key fields (or was NaN)
  // thread = (channel, quarter of the slices): the loads of a quarter are independent, the quarters are added in order
  const int ch = threadIdx.x & 255, grp = threadIdx.x >> 8;
  const int per = (slices + 3) >> 2, s0 = grp * per, s1 = min(s0 + per, slices);
  double acc = 0.0;
  if (ch < c) {
#pragma unroll 8
    for (int sl = s0; sl < s1; ++sl) acc += part[((size_t)b * slices + sl) * c + ch];
  }
  s_grp[grp][ch] = acc;
  __syncthreads();
  double g = 0.0;
  if (threadIdx.x < 256 && ch < c) {
    const double tot = ((s_grp[0][ch] + s_grp[1][ch]) + s_grp[2][ch]) + s_grp[3][ch];
    g = (hi > lo) ? pow(tot / (double)(hi - lo), 1.0 / (double)p) : 0.0;
    if (poisoned) g = __longlong_as_double(0x7ff8000000000000ll);
    if (!normalize) out[(size_t)b * c + ch] = g;
  }
  if (!normalize) return;
  double sq = (threadIdx.x < 256) ? warp_sum(g * g) : 0.0;
  if (threadIdx.x < 256 && (threadIdx.x & 31) == 0) s_sq[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 256 && ch < c) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += s_sq[k];
    out[(size_t)b * c + ch] = g * (1.0 / fmax(sqrt(tot), 1e-12));  // F.normalize eps
  }
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
    CK(cudaMalloc(&t->tab[l].rows, sizeof(int) * cap));
    t->tab[l].mask = cap - 1;
    CK(cudaMalloc(&t->keys[l], sizeof(unsigned long long) * max_points));
    if (l >= 1) {
      CK(cudaMalloc(&t->kmap3[l], sizeof(int) * 27 * (size_t)max_points));
      CK(cudaMalloc(&t->kmap2[l], sizeof(int) * 8 * (size_t)max_points));
      CK(cudaMalloc(&t->kmask3[l], sizeof(unsigned) * (size_t)max_points));
      CK(cudaMalloc(&t->kmask2[l], sizeof(unsigned) * (size_t)max_points));
    }
  }
  CK(cudaMalloc(&t->kmapt, sizeof(int) * 8 * (size_t)max_points));
  CK(cudaMalloc(&t->kmaskt, sizeof(unsigned) * (size_t)max_points));
  CK(cudaMalloc(&t->d_n, sizeof(int) * TCN_CTL_INTS));
  for (int m = 0; m < 7; ++m) {
    CK(cudaMalloc(&t->plist[m], sizeof(int2) * (m < 3 ? 27 : 8) * (size_t)max_points));
    CK(cudaMalloc(&t->pbase[m], sizeof(int) * (size_t)max_points));
  }
  t->partial_floats = (size_t)max_points * 2048;  // 27 offsets x 64 channels / 8 offsets x 256 channels per point
  CK(cudaMalloc(&t->partial, sizeof(float) * t->partial_floats));
  CK(cudaMalloc(&t->slot_of, sizeof(int) * 4 * (size_t)max_points));
  CK(cudaMalloc(&t->blk_cnt, sizeof(int) * 4 * (size_t)((max_points + 255) / 256)));
  t->pool_floats = (size_t)max_points * 1184;
  CK(cudaMalloc(&t->pool, sizeof(float) * t->pool_floats));
  t->gem_slices = max_batch * TCN_GEM_SLICES > 256 ? max_batch * TCN_GEM_SLICES : 256;  // rows of the partial-sum buffer
  CK(cudaMalloc(&t->gem_part, sizeof(double) * (size_t)t->gem_slices * 256));
  *out = t;
  return MT_OK;
}

extern "C" int mt_tcn_destroy(mt_tcn* t) {
  if (!t) return MT_OK;
  cudaSetDevice(t->device);
  for (int l = 0; l < 4; ++l) {
    cudaFree(t->tab[l].keys), cudaFree(t->tab[l].vals), cudaFree(t->tab[l].rows), cudaFree(t->keys[l]);
    cudaFree(t->kmap3[l]), cudaFree(t->kmap2[l]), cudaFree(t->kmask3[l]), cudaFree(t->kmask2[l]);
  }
  cudaFree(t->kmapt), cudaFree(t->kmaskt), cudaFree(t->slot_of), cudaFree(t->blk_cnt), cudaFree(t->partial);
  for (int m = 0; m < 7; ++m) cudaFree(t->plist[m]), cudaFree(t->pbase[m]);
  for (int i = 0; i < TCN_NCONV; ++i) cudaFree(t->conv[i].w), cudaFree(t->conv[i].wt);
  for (int i = 0; i < TCN_NBN; ++i) cudaFree(t->bn[i].scale), cudaFree(t->bn[i].shift);
  cudaFree(t->d_n), cudaFree(t->pool), cudaFree(t->gem_part);
  delete t;
  return MT_OK;
}

extern "C" int mt_tcn_set_conv(mt_tcn* t, int id, const float* h_kernel, int kvol, int cin, int cout) {
  if (!t || id < 0 || id >= TCN_NCONV || !h_kernel || kvol <= 0 || cin <= 0 || cout <= 0 || cout > 256)
    return set_err(MT_ERR_ARG, "mt_tcn_set_conv: bad argument (cout <= 256)");
  CK(cudaSetDevice(t->device));
  cudaFree(t->conv[id].w), cudaFree(t->conv[id].wt);
  t->conv[id].w = t->conv[id].wt = nullptr;
  const size_t n = (size_t)kvol * cin * cout;
  CK(cudaMalloc(&t->conv[id].w, sizeof(float) * n));
  CK(cudaMemcpy(t->conv[id].w, h_kernel, sizeof(float) * n, cudaMemcpyHostToDevice));
  std::vector<float> tr(n);
  for (int i = 0; i < kvol; ++i)
    for (int k = 0; k < cin; ++k)
      for (int c = 0; c < cout; ++c) tr[((size_t)i * cout + c) * cin + k] = h_kernel[((size_t)i * cin + k) * cout + c];
  CK(cudaMalloc(&t->conv[id].wt, sizeof(float) * n));
  CK(cudaMemcpy(t->conv[id].wt, tr.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
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

// One convolution.  map / hitmask / kvol describe the gather (NULL: 1x1, the point itself); layers whose widths the MMA
// tiling divides (cin % 16 == 0, cout % 16 == 0: every layer of the shipped network but conv0) run on the tensor cores,
// the others through the hash-probing CUDA-core kernel (mode / k / dil / in_tab as before).
static int tcn_conv_launch(mt_tcn* t, int mode, int conv_id, int bn_id, const unsigned long long* out_keys, const int* d_nout,
                           int nmax, const TcnTable& in_tab, const float* in_feat, int k, int dil, const int* map,
                           const unsigned* hitmask, int map_id, const float* residual, int relu, int accumulate, float* out_feat,
                           cudaStream_t st) {
  const TcnConv& c = t->conv[conv_id];
  if (!c.w) return set_err(MT_ERR_STATE, "mt_tcn_forward: a convolution has no weights (mt_tcn_set_conv)");
  if (mode == 0 && c.kvol != k * k * k) return set_err(MT_ERR_STATE, "mt_tcn_forward: kernel volume does not match the layer");
  if (mode == 1 && c.kvol != 8) return set_err(MT_ERR_STATE, "mt_tcn_forward: the transposed convolution has 8 kernel offsets");
  const float* sc = nullptr;
  const float* sh = nullptr;
  if (bn_id >= 0) {
    if (!t->bn[bn_id].scale || t->bn[bn_id].c != c.cout) return set_err(MT_ERR_STATE, "mt_tcn_forward: BatchNorm missing / wrong width");
    sc = t->bn[bn_id].scale, sh = t->bn[bn_id].shift;
  }
  const bool have_map = map != nullptr || k == 1;
  static const bool no_mma = getenv("MIDAS_B200_TCN_NO_MMA") != nullptr;  // diagnostics: every layer through the CUDA-core kernel
  static const bool no_pairs = getenv("MIDAS_B200_TCN_NO_PAIRS") != nullptr;  // diagnostics: dense tiles instead of pair lists
  if (!no_mma && !no_pairs && in_feat && map && map_id >= 0 && c.cin % 16 == 0 && c.cout % 16 == 0 &&
      (size_t)c.kvol * c.cout * t->max_points <= t->partial_floats) {
    // work proportional to the pairs present: pair GEMM into `partial`, then the per-point sum + epilogue
    const int* ctr = t->d_n + TCN_CTL_PAIRS + 32 * map_id;
    const long long items = ((long long)(nmax + 31) / 32 + c.kvol) * (c.cout / 16);  // at least the centre offset's tiles
    const unsigned grid = (unsigned)std::min<long long>((items + 3) / 4, 148 * 12);
    tcn_launch(k_tcn_pair_mma<2>, grid, 128, st, t->plist[map_id], ctr, c.kvol, t->max_points, in_feat, c.cin, c.wt, c.cout, t->partial);
    CK_LAUNCH();
    const long long thr = (long long)nmax * (c.cout / 4);
    tcn_launch(k_tcn_pair_reduce, (unsigned)((thr + 255) / 256), 256, st, t->pbase[map_id], hitmask, d_nout, t->partial, c.cout, sc, sh, residual, relu,
                                                                   accumulate, out_feat);
    CK_LAUNCH();
    return MT_OK;
  }
  if (!no_mma && in_feat && have_map && c.cin % 16 == 0 && c.cout % 16 == 0) {
    const int mtiles = (nmax + 15) / 16;
    const int kvol = map ? c.kvol : 1;
    if (c.cout % 32 == 0 && (long long)mtiles * (c.cout / 16) > 4096) {  // many tiles: wider slabs, fewer re-reads of the gathered rows
      const long long warps = (long long)mtiles * (c.cout / 32);
      tcn_launch(k_tcn_conv_mma<4>, (unsigned)((warps + 3) / 4), 128, st, map, hitmask, kvol, d_nout, in_feat, c.cin, c.wt, c.cout, sc, sh, residual,
                                                                   relu, accumulate, out_feat);
    } else {
      const long long warps = (long long)mtiles * (c.cout / 16);
      tcn_launch(k_tcn_conv_mma<2>, (unsigned)((warps + 3) / 4), 128, st, map, hitmask, kvol, d_nout, in_feat, c.cin, c.wt, c.cout, sc, sh, residual,
                                                                   relu, accumulate, out_feat);
    }
    CK_LAUNCH();
    return MT_OK;
  }
  // few points and many channels: four warps per point split the input channels (shorter FMA chains);
  // conv0 (one input feature) keeps a warp per point
  const bool split = in_feat != nullptr && c.cin >= 32;
  const unsigned grid = (unsigned)(split ? (nmax + 1) / 2 : (nmax + 7) / 8);
  if (mode == 0 && split)
    tcn_launch(k_tcn_conv<0, 4>, grid, 256, st, out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else if (mode == 0)
    tcn_launch(k_tcn_conv<0, 1>, grid, 256, st, out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else if (split)
    tcn_launch(k_tcn_conv<1, 4>, grid, 256, st, out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  else
    tcn_launch(k_tcn_conv<1, 1>, grid, 256, st, out_keys, d_nout, in_tab, in_feat, c.cin, c.w, c.cout, k, dil, sc, sh, residual, relu, accumulate, out_feat);
  CK_LAUNCH();
  return MT_OK;
}

// the whole forward pass from n raw points (float clouds or packed keys)
static int tcn_run(mt_tcn* t, const float* d_pts, const unsigned long long* d_keys, int n, int P, float inv_q, int batch,
                   int normalize, double* d_out, int* d_counts, cudaStream_t st) {
  const int c0 = t->conv[TCN_CONV0].cout, p0 = t->conv[TCN_BLK_C2(0)].cout, p1 = t->conv[TCN_BLK_C2(1)].cout,
            p2 = t->conv[TCN_BLK_C2(2)].cout, f = t->conv[TCN_LAT0].cout;
  if (!c0 || !p0 || !p1 || !p2 || !f) return set_err(MT_ERR_STATE, "mt_tcn_forward: weights not loaded");
  if (f > 256) return set_err(MT_ERR_STATE, "mt_tcn_forward: feature size above 256");
  const int widths[] = {c0, c0, p0, p0, p0, p0, p1, p1, p1, p1, p2, p2, p2, f, f};
  size_t need = 0;
  for (int w : widths) need += (size_t)n * w;
  if (need > t->pool_floats) return set_err(MT_ERR_CAPACITY, "mt_tcn_forward: feature pool too small for these channel widths");
  TcnTabs T;
  for (int l = 0; l < 4; ++l) T.t[l] = t->tab[l];
  int* d_flag = t->d_n + 4;
  const int nblk = (n + TCN_CBLOCK - 1) / TCN_CBLOCK, stride = t->max_points;
  tcn_launch(k_tcn_clear_all, (t->cap + 255) / 256, 256, st, T, t->d_n);
  CK_LAUNCH();
  tcn_launch(k_tcn_insert_all, (4 * n + 255) / 256, 256, st, d_pts, d_keys, n, P, inv_q, T, t->slot_of, stride, d_flag);
  CK_LAUNCH();
  tcn_launch(k_tcn_count_all, nblk, TCN_CBLOCK, st, n, T, t->slot_of, stride, t->blk_cnt, nblk);
  CK_LAUNCH();
  tcn_launch(k_tcn_scatter_all, nblk, TCN_CBLOCK, st, n, T, t->slot_of, stride, t->blk_cnt, nblk, t->keys[0], t->keys[1], t->keys[2], t->keys[3], t->d_n);
  CK_LAUNCH();
  TcnMaps M;
  memset(&M, 0, sizeof(M));
  for (int l = 1; l < 4; ++l) M.m3[l] = t->kmap3[l], M.m2[l] = t->kmap2[l], M.q3[l] = t->kmask3[l], M.q2[l] = t->kmask2[l];
  M.mt = t->kmapt, M.qt = t->kmaskt, M.ncap = t->max_points;
  for (int m = 0; m < 7; ++m) M.pl[m] = t->plist[m], M.pb[m] = t->pbase[m];
  tcn_launch(k_tcn_kmaps, dim3((n + 7) / 8, 3), 256, st, T, t->keys[1], t->keys[2], t->keys[3], t->d_n, M);
  CK_LAUNCH();
  // feature buffers carved from the pool (upper bound n rows each)
  float* ptr = t->pool;
  auto take = [&](int w) { float* r = ptr; ptr += (size_t)n * w; return r; };
  float* x0 = take(c0);
  int r;
#define TCN_RUN(call) if ((r = (call)) != MT_OK) return r
  // conv0 + bn0 + relu (minkfpn.py:113-115)
  {
    const TcnConv& cv = t->conv[TCN_CONV0];
    int k0 = 1;
    while (k0 * k0 * k0 < cv.kvol) ++k0;
    if (cv.cin == 1 && (k0 & 1) && k0 * k0 * k0 == cv.kvol && cv.kvol <= 128 && t->bn[TCN_BN0].scale && t->bn[TCN_BN0].c == cv.cout) {
      tcn_launch(k_tcn_conv0, (n + 7) / 8, 256, st, t->keys[0], t->d_n, t->tab[0], cv.w, cv.cout, k0, t->bn[TCN_BN0].scale, t->bn[TCN_BN0].shift, x0);
      CK_LAUNCH();
    } else {
      TCN_RUN(tcn_conv_launch(t, 0, TCN_CONV0, TCN_BN0, t->keys[0], t->d_n, n, t->tab[0], nullptr, k0, 1, nullptr, nullptr, -1, nullptr, 1, 0, x0, st));
    }
  }
  float* x = x0;
  float* fmap = nullptr;
  for (int s = 0; s < 3; ++s) {  // bottom-up: strided conv + bn + relu + BasicBlock (minkfpn.py:120-126)
    const int stride_in = 1 << s, l = s + 1, pl = t->conv[TCN_BLK_C2(s)].cout;
    float* xa = take(t->conv[TCN_DOWN(s)].cout);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_DOWN(s), TCN_BN_DOWN(s), t->keys[l], t->d_n + l, n, t->tab[l - 1], x, 2, stride_in, t->kmap2[l],
                            t->kmask2[l], 3 + s, nullptr, 1, 0, xa, st));
    float* y = take(pl);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_C1(s), TCN_BN_N1(s), t->keys[l], t->d_n + l, n, t->tab[l], xa, 3, 2 * stride_in, t->kmap3[l],
                            t->kmask3[l], s, nullptr, 1, 0, y, st));
    const float* res = xa;
    if (t->conv[TCN_BLK_DS(s)].w) {  // channel change: residual = bn(conv1x1(x)) (resnet.py:89-101)
      float* rs = take(pl);
      TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_DS(s), TCN_BN_DS(s), t->keys[l], t->d_n + l, n, t->tab[l], xa, 1, 2 * stride_in, nullptr, nullptr,
                              -1, nullptr, 0, 0, rs, st));
      res = rs;
    } else if (t->conv[TCN_DOWN(s)].cout != pl) {
      return set_err(MT_ERR_STATE, "mt_tcn_forward: block changes width but has no downsample branch");
    }
    float* xo = take(pl);
    TCN_RUN(tcn_conv_launch(t, 0, TCN_BLK_C2(s), TCN_BN_N2(s), t->keys[l], t->d_n + l, n, t->tab[l], y, 3, 2 * stride_in, t->kmap3[l],
                            t->kmask3[l], s, res, 1, 0, xo, st));
    x = xo;
    if (s == 1) fmap = xo;
  }
  // lateral 1x1 at the top, transposed conv down to stride 4, + lateral 1x1 of the stage-1 map (minkfpn.py:130-136)
  float* z = take(f);
  TCN_RUN(tcn_conv_launch(t, 0, TCN_LAT0, -1, t->keys[3], t->d_n + 3, n, t->tab[3], x, 1, 8, nullptr, nullptr, -1, nullptr, 0, 0, z, st));
  float* fp = take(f);
  TCN_RUN(tcn_conv_launch(t, 0, TCN_LAT1, -1, t->keys[2], t->d_n + 2, n, t->tab[2], fmap, 1, 4, nullptr, nullptr, -1, nullptr, 0, 0, fp, st));
  TCN_RUN(tcn_conv_launch(t, 1, TCN_TCONV, -1, t->keys[2], t->d_n + 2, n, t->tab[3], z, 2, 4, t->kmapt, t->kmaskt, 6, nullptr, 0, 1, fp, st));
#undef TCN_RUN
  const int slices = ((long long)batch * 256 <= t->gem_slices) ? 256 : TCN_GEM_SLICES;
  tcn_launch(k_tcn_gem_partial, dim3(batch, slices), 256, st, t->d_n, fp, f, t->gem_p, t->gem_eps, t->gem_part);
  CK_LAUNCH();
  tcn_launch(k_tcn_gem, batch, 1024, st, t->d_n, t->gem_part, slices, f, t->gem_p, normalize, d_out);
  CK_LAUNCH();
  if (d_counts) CK(cudaMemcpyAsync(d_counts, t->d_n, sizeof(int) * 4, cudaMemcpyDeviceToDevice, st));
  return MT_OK;
}

// d_keys: n packed coordinates (tcn_pack: batch | x | y | z) of the quantised clouds, batch-major (sorted unique keys as
// ME.utils.sparse_quantize + batched_coordinates produce them; duplicates are merged)
// d_out: (batch, feature) float64.  d_counts (nullable): active points per level (4 ints, device).
extern "C" int mt_tcn_forward(mt_tcn* t, const unsigned long long* d_keys, int n, int batch, int normalize, double* d_out,
                              int* d_counts, void* stream) {
  if (!t || !d_keys || !d_out || n <= 0 || batch <= 0) return set_err(MT_ERR_ARG, "mt_tcn_forward: bad argument");
  if (n > t->max_points || batch > t->max_batch) return set_err(MT_ERR_CAPACITY, "mt_tcn_forward: n / batch exceed the context capacity");
  CK(cudaSetDevice(t->device));
  return tcn_run(t, nullptr, d_keys, n, 1, 0.f, batch, normalize, d_out, d_counts, (cudaStream_t)stream);
}

// d_clouds: (batch, points, 3) float32, the sampled and scaled clouds of tcn.py:96-123; quantisation (floor(x * inv_q),
// inv_q = the float32 reciprocal torch's `cloud / q` multiplies by on CUDA: float32(1 / q) in torch 2.x), duplicate removal
// and the network run in one enqueue -- no sort, no host synchronisation.  A coordinate beyond +-131071 voxels (or NaN)
// poisons the descriptors of the call with NaN.
extern "C" int mt_tcn_embed(mt_tcn* t, const float* d_clouds, int batch, int points, float inv_q, int normalize, double* d_out,
                            int* d_counts, void* stream) {
  if (!t || !d_clouds || !d_out || batch <= 0 || points <= 0 || !(inv_q > 0.f)) return set_err(MT_ERR_ARG, "mt_tcn_embed: bad argument");
  const long long n = (long long)batch * points;
  if (n > t->max_points || batch > t->max_batch) return set_err(MT_ERR_CAPACITY, "mt_tcn_embed: batch x points / batch exceed the context capacity");
  CK(cudaSetDevice(t->device));
  return tcn_run(t, d_clouds, nullptr, (int)n, points, inv_q, batch, normalize, d_out, d_counts, (cudaStream_t)stream);
}
