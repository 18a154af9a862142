// libmidas_b200: hand-written sm_100a kernels + C ABI for the MidasTouch particle-filter
// hot path (include/midas_b200.h).  No host compute path: every entry point launches
// CUDA kernels on the caller's stream.
//
// Kernel inventory (DESIGN.md has the roofline for each):
//   k_cosine_rows      1 x D against rows x D cosine (codebook query / get_similarity)
//   k_cosine_batched   Q x D against rows x D, fp32 SIMT tiles
//   k_step_a           motion + SE(3) key + exact grid 1-NN + weight lookup + chunk sums
//   k_step_b           normalise + float64 prefix + systematic draw + child scatter
//   k_nn_grid/brute    standalone SE3_NN index search
//   k_resample_*       systematic resampling of explicit float64 weights
//   small: converters, gathers, rmse, softmax
#include <cuda.h>
#include <cuda_runtime.h>
#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/midas_b200.h"
#include "mt_math.cuh"
#define MT_NN_BLOCK 256
#include "mt_nn.cuh"
#include "mt_mesh.cuh"

#define MT_MAX_D 6144  // query staged in 48 KB of shared memory as float64
#define MT_MAX_WORLD 16
// exchange buffer of the sharded fused step, double-buffered by step parity.  A weight sum travels as two
// self-validating 64-bit words, (low half of the double << 32 | seq) and (high half << 32 | seq), seq = the low 32
// bits of the exchange counter: every 8-byte store is atomic, so a reader that sees the expected seq in both
// words has the value -- no fence on either side.
struct Xchg {
  unsigned long long w[2][MT_MAX_WORLD][2];
};
#define MT_CHUNK 256  // particles per chunk == threads per block of the sweep kernels

// nn_cur / nn_next hold the codebook match of every particle; a particle whose weight was
// zeroed by the drift test or check_quats stores -(idx + 2) (so -1 stays "no match yet").
__host__ __device__ __forceinline__ int nn_masked(int idx) { return -(idx + 2); }
__host__ __device__ __forceinline__ int nn_index(int stored) { return stored < -1 ? -(stored + 2) : stored; }
__host__ __device__ __forceinline__ bool nn_is_masked(int stored) { return stored < -1; }

// ------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return set_err(MT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)
#define CK_LAUNCH() CK(cudaGetLastError())

extern "C" const char* mt_last_error(void) { return g_err; }
extern "C" int mt_version(void) { return 100; }

// ------------------------------------------------------------------------- context
struct mt_ctx {
  int device;
  int sm_count;
  size_t cap;
  int M, D;
  // codebook
  float4* d_keys_orig;    // M x 2 float4 (k0..k3 | k4,k5,0,0), original order
  float4* d_keys_sorted;  // same, sorted by grid cell
  int* d_sorted_orig;     // scratch: partner index per key during upload
  float4* d_bvh;          // boxes of the search index: leaves | level 1 | level 2 (3 float4 each)
  float4* d_nbr;          // M x nbr_k x 2 float4 neighbour lists (mt_nn.cuh)
  int nbr_k;              // list length: 64, 128 or 256 (chosen at upload)
  BvhParams bvh;
  const void* d_emb;
  int emb_dtype;
  double* d_sim;   // cos(q, E_m)
  double* d_esim;  // exp(cos)
  double* d_rnorm; // max(|E_m|, 1e-8), computed by the first query after an upload
  bool rnorm_ready;
  int query_blocks_per_sm;
  bool cb_ready;
  void* d_scratch;       // grow-on-demand scratch of the cluster / selection entry points
  size_t scratch_bytes;
  cudaEvent_t timing[4];  // optional: recorded around the kernels of mt_step_a (bench instrumentation)
  // down-sampled mesh vertices for the drift test (mt_mesh.cuh)
  double* d_mesh_verts;
  float4* d_mesh_verts32;
  int* d_mesh_vox;
  unsigned* d_mesh_vox2;
  MeshVoxels vox;
  int* d_mesh_cells;
  MeshGrid mesh;
  int mesh_V;
  bool mesh_ready;
  // scratch
  int chunk_cap;
  double* d_part;     // per-chunk weight sums
  double* d_prefix;   // exclusive prefix of d_part, [nchunks] = total
  double* d_rm_part;  // 2 x chunk_cap rmse partials
  int warp_cap;
  double* d_wpart;    // per-warp weight sums of kernel A (32 particles each)
  double* d_wrm;      // 2 x warp_cap rmse partials of kernel A
  int* d_wcnt;        // per-warp count of particles that passed the drift test
  float4* d_rec;      // search records of the queued particles (32 B each: key | best distance, best index), k_step_a -> k_step_nnq
  int* d_queue2;      // drift tests left for the grid search
  int* d_queue;       // particles whose hint-graph search was not conclusive (kernel A -> A2)
  unsigned int* d_qctl;  // [0] searches queued, [1] queue head, [3] deferred drift tests queued
  unsigned long long* d_bar;  // grid barrier of k_step_bw (monotone counter)
  unsigned long long bar_target;
  double* d_bw;       // 3 x MT_BW_MAX_GRID block totals (weights, rmse_t, rmse_r)
  int* d_bwcnt;       // MT_BW_MAX_GRID on-surface counts
  int bw_blocks_per_sm;
  // sharded fused step: every GPU writes its weight sum straight into its peers' exchange buffers over
  // NVLink (CUDA IPC mappings) -- no collective call between the two phases of k_step_bw
  struct Xchg* d_xchg;        // own exchange buffer (peer-mapped by the other ranks)
  struct Xchg** d_peers;      // device array: peers[r] = rank r's buffer as seen from this GPU
  struct Xchg* h_peers[MT_MAX_WORLD];
  int peers_world;            // 0: not imported
  unsigned long long xchg_count;  // fused sharded steps launched so far (same on every rank)
  unsigned long long* d_xdbg;     // last exchange: globaltimer at barrier exit / sums sent / all sums received (block 0)
  double* d_q64;      // staged query, float64, MT_MAX_D entries
  double* d_scal;     // [0] local weight sum, [1] max, [2] min, [3] softmax denom, [4..7] spare
  unsigned int* d_ticket;
  int* d_flags;  // MT_STAT_* slots (include/midas_b200.h)
  // batched query: float32 split planes of the codebook (big = TF32-exact part, small = remainder) + their TMA maps
  float *d_plane_big, *d_plane_small;
  CUtensorMap tm_a_big, tm_a_small;
  bool planes_ready;
  // mt_step: the codebook query runs on a side stream (or as a parallel branch of the step's CUDA graph)
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  struct StepGraph* graphs;  // MT_STEP_GRAPHS cached instantiations (one per buffer parity / configuration)
  int graph_next;
  unsigned long long graph_replays;
};

static void step_graphs_free(mt_ctx* c);
static size_t nchunks_of(long long n) { return (size_t)((n + MT_CHUNK - 1) / MT_CHUNK); }

extern "C" int mt_ctx_create(int device, size_t capacity, int M, int D, mt_ctx** out) {
  if (!out || capacity == 0 || M <= 0 || D <= 0) return set_err(MT_ERR_ARG, "mt_ctx_create: bad argument");
  if (capacity > 0x07ffffffull) return set_err(MT_ERR_CAPACITY, "mt_ctx_create: at most 2^27 - 1 particles per context (queue entries carry the index in 27 bits)");
  CK(cudaSetDevice(device));
  mt_ctx* c = new mt_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  c->cap = capacity;
  c->M = M;
  c->D = D;
  c->chunk_cap = (int)nchunks_of((long long)capacity) + 2;
  CK(cudaMalloc(&c->d_keys_orig, sizeof(float4) * 2 * M));
  CK(cudaMalloc(&c->d_keys_sorted, sizeof(float4) * 2 * M));
  CK(cudaMalloc(&c->d_sorted_orig, sizeof(int) * M));
  CK(cudaMalloc(&c->d_nbr, sizeof(float4) * 2 * MT_NBR_K * (size_t)M));
  c->nbr_k = MT_NBR_K;
  CK(cudaMalloc(&c->d_sim, sizeof(double) * M));
  CK(cudaMalloc(&c->d_esim, sizeof(double) * M));
  CK(cudaMalloc(&c->d_rnorm, sizeof(double) * M));
  CK(cudaMalloc(&c->d_part, sizeof(double) * c->chunk_cap));
  c->warp_cap = (int)((capacity + 31) / 32) + 8;
  CK(cudaMalloc(&c->d_wpart, sizeof(double) * c->warp_cap));
  CK(cudaMalloc(&c->d_wrm, sizeof(double) * 2 * c->warp_cap));
  CK(cudaMalloc(&c->d_wcnt, sizeof(int) * c->warp_cap));
  CK(cudaMalloc(&c->d_queue, sizeof(int) * (capacity + 32)));
  CK(cudaMalloc(&c->d_queue2, sizeof(int) * (capacity + 32)));
  CK(cudaMalloc(&c->d_rec, sizeof(float4) * 2 * (capacity + 32)));
  CK(cudaMalloc(&c->d_qctl, sizeof(unsigned int) * 8));
  CK(cudaMemset(c->d_qctl, 0, sizeof(unsigned int) * 8));
  CK(cudaMalloc(&c->d_bar, sizeof(unsigned long long)));
  CK(cudaMemset(c->d_bar, 0, sizeof(unsigned long long)));
  CK(cudaMalloc(&c->d_xchg, sizeof(Xchg)));
  CK(cudaMemset(c->d_xchg, 0, sizeof(Xchg)));
  CK(cudaMalloc(&c->d_peers, sizeof(Xchg*) * MT_MAX_WORLD));
  CK(cudaMalloc(&c->d_xdbg, sizeof(unsigned long long) * 64));
  CK(cudaMemset(c->d_xdbg, 0, sizeof(unsigned long long) * 64));
  CK(cudaMalloc(&c->d_bw, sizeof(double) * 3 * 1184));
  CK(cudaMalloc(&c->d_bwcnt, sizeof(int) * 1184));
  CK(cudaMalloc(&c->d_prefix, sizeof(double) * (c->chunk_cap + 1)));
  CK(cudaMalloc(&c->d_rm_part, sizeof(double) * 2 * c->chunk_cap));
  CK(cudaMalloc(&c->d_scal, sizeof(double) * 8));
  CK(cudaMalloc(&c->d_q64, sizeof(double) * MT_MAX_D));
  CK(cudaMalloc(&c->d_ticket, sizeof(unsigned int) * 4));
  CK(cudaMalloc(&c->d_flags, sizeof(int) * MT_STAT_COUNT));
  CK(cudaMemset(c->d_ticket, 0, sizeof(unsigned int) * 4));
  CK(cudaMemset(c->d_flags, 0, sizeof(int) * MT_STAT_COUNT));
  CK(cudaMemset(c->d_scal, 0, sizeof(double) * 8));
  CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  *out = c;
  return MT_OK;
}

extern "C" int mt_ctx_destroy(mt_ctx* c) {
  if (!c) return MT_OK;
  cudaSetDevice(c->device);
  step_graphs_free(c);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  cudaFree(c->d_keys_orig);
  cudaFree(c->d_keys_sorted);
  cudaFree(c->d_sorted_orig);
  cudaFree(c->d_bvh);
  cudaFree(c->d_nbr);
  cudaFree(c->d_plane_big);
  cudaFree(c->d_plane_small);
  cudaFree(c->d_mesh_verts);
  cudaFree(c->d_mesh_verts32);
  cudaFree(c->d_mesh_vox);
  cudaFree(c->d_mesh_vox2);
  cudaFree(c->d_mesh_cells);
  cudaFree(c->d_sim);
  cudaFree(c->d_esim);
  cudaFree(c->d_rnorm);
  cudaFree(c->d_part);
  cudaFree(c->d_prefix);
  cudaFree(c->d_rm_part);
  cudaFree(c->d_wpart);
  cudaFree(c->d_wrm);
  cudaFree(c->d_wcnt);
  cudaFree(c->d_queue);
  cudaFree(c->d_queue2);
  cudaFree(c->d_rec);
  cudaFree(c->d_scratch);
  cudaFree(c->d_qctl);
  cudaFree(c->d_bar);
  for (int r = 0; r < c->peers_world; ++r)
    if (c->h_peers[r] && c->h_peers[r] != c->d_xchg) cudaIpcCloseMemHandle(c->h_peers[r]);
  cudaFree(c->d_xchg);
  cudaFree(c->d_peers);
  cudaFree(c->d_xdbg);
  cudaFree(c->d_bw);
  cudaFree(c->d_bwcnt);
  cudaFree(c->d_scal);
  cudaFree(c->d_q64);
  cudaFree(c->d_ticket);
  cudaFree(c->d_flags);
  delete c;
  return MT_OK;
}

// ------------------------------------------------------------------------- codebook grid
extern "C" int mt_codebook_upload(mt_ctx* c, const float* h_keys, const void* d_emb, int emb_dtype) {
  if (!c || !h_keys) return set_err(MT_ERR_ARG, "mt_codebook_upload: null argument");
  if (emb_dtype != MT_DTYPE_F32 && emb_dtype != MT_DTYPE_F64) return set_err(MT_ERR_ARG, "mt_codebook_upload: dtype");
  CK(cudaSetDevice(c->device));
  const int M = c->M;
  // ---- search index: 6-D Morton order, leaves of 32 keys, two levels of 32-ary boxes (mt_nn.cuh)
  MtBvhHost bvh;
  if (!mt_bvh_build(h_keys, M, bvh)) return set_err(MT_ERR_ARG, "mt_codebook_upload: NaN or Inf key");
  const BvhParams bp = bvh.bp;
  const std::vector<float>&ks = bvh.keys_sorted, &leaf = bvh.leaf, &l1 = bvh.l1, &l2 = bvh.l2;
  std::vector<float> ko(8 * (size_t)M, 0.f);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 6; ++k) ko[8 * (size_t)m + k] = h_keys[6 * m + k];
  if (c->d_bvh) cudaFree(c->d_bvh), c->d_bvh = nullptr;
  CK(cudaMalloc(&c->d_bvh, sizeof(float) * 12 * ((size_t)bp.n_leaf + bp.n_l1 + bp.n_l2)));
  CK(cudaMemcpy(c->d_bvh, leaf.data(), sizeof(float) * leaf.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_bvh + 3 * (size_t)bp.n_leaf, l1.data(), sizeof(float) * l1.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_bvh + 3 * ((size_t)bp.n_leaf + bp.n_l1), l2.data(), sizeof(float) * l2.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_keys_orig, ko.data(), sizeof(float) * 8 * M, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_keys_sorted, ks.data(), sizeof(float) * 8 * M, cudaMemcpyHostToDevice));
  // ---- neighbour lists.  Pass 0 builds the 64 nearest other keys of every key.  On a dense codebook (thin parts:
  // thousands of keys per cm of a rod) the particles sit further off the key manifold -- in rotation -- than the 64th
  // neighbour is away, the triangle-inequality ball of the hint scan then holds more than 64 keys and most scans end
  // inconclusive (measured on the cotter-pin stand-in: 60 % of the particles, each then a 20-leaf box search).  Such
  // codebooks get longer lists: 128 or 256 entries, built in further passes of 64 (each admits only keys beyond the
  // previous pass's last entry).  Criterion (a heuristic): median 64th-neighbour distance below 1e-2 / 8e-3 key units
  // (the stand-ins: cotter pin 4.4e-3, sugar box 1.1e-2, mug 1.2e-2, drill 1.3e-2); MIDAS_B200_NBR_K = 64 | 128 | 256
  // overrides it.  Measured on the cotter-pin stand-in at 2^20 particles: 64 -> 128 -> 256 entries = 597k -> 417k ->
  // 301k box searches per step, 2.47 -> 2.14 -> 1.96 ms per step -- a help, not a cure: a particle that slides 0.7 mm
  // round a 3 mm rod is 27 degrees (4.7e-3 key units) away from every key at its new position.
  if (c->nbr_k != MT_NBR_K) {
    cudaFree(c->d_nbr);
    c->d_nbr = nullptr;
    CK(cudaMalloc(&c->d_nbr, sizeof(float4) * 2 * MT_NBR_K * (size_t)M));
    c->nbr_k = MT_NBR_K;
  }
  float2* d_bound = nullptr;
  CK(cudaMalloc(&d_bound, sizeof(float2) * (size_t)M));
  k_build_nbr<0><<<(M + 7) / 8, 256>>>(c->d_keys_orig, M, c->d_nbr, nullptr, nullptr, 0, 0, nullptr, MT_NBR_K, 0, d_bound);
  CK_LAUNCH();
  int want = MT_NBR_K;
  if (const char* e = getenv("MIDAS_B200_NBR_K")) {
    want = atoi(e);
    if (want != 64 && want != 128 && want != 256) {
      cudaFree(d_bound);
      return set_err(MT_ERR_ARG, "MIDAS_B200_NBR_K must be 64, 128 or 256");
    }
  } else if (M > MT_NBR_K_MAX) {
    float* d_last = nullptr;
    CK(cudaMalloc(&d_last, sizeof(float) * (size_t)M));
    k_nbr_last_delta<<<(M + 255) / 256, 256>>>(c->d_nbr, M, MT_NBR_K, d_last);
    CK_LAUNCH();
    std::vector<float> last(M);
    CK(cudaMemcpy(last.data(), d_last, sizeof(float) * (size_t)M, cudaMemcpyDeviceToHost));
    cudaFree(d_last);
    std::nth_element(last.begin(), last.begin() + M / 2, last.end());
    const float med = last[M / 2];
    want = med < 8e-3f ? 256 : (med < 1e-2f ? 128 : 64);
  }
  if (want > MT_NBR_K && M > want) {
    float4* big = nullptr;
    CK(cudaMalloc(&big, sizeof(float4) * 2 * (size_t)want * (size_t)M));
    const size_t tot = (size_t)M * MT_NBR_K * 2;
    k_nbr_restride<<<(unsigned)((tot + 255) / 256), 256>>>(c->d_nbr, M, want, big);
    CK_LAUNCH();
    for (int pass = 1; pass < want / MT_NBR_K; ++pass) {
      k_build_nbr<0><<<(M + 7) / 8, 256>>>(c->d_keys_orig, M, big, nullptr, nullptr, 0, 0, nullptr, want, pass, d_bound);
      CK_LAUNCH();
    }
    CK(cudaDeviceSynchronize());
    cudaFree(c->d_nbr);
    c->d_nbr = big;
    c->nbr_k = want;
  }
  cudaFree(d_bound);
  k_build_nbr<1><<<(M + 7) / 8, 256>>>(c->d_keys_orig, M, nullptr, c->d_sorted_orig);
  CK_LAUNCH();
  k_set_partner<<<(M + 255) / 256, 256>>>(c->d_keys_orig, M, c->d_sorted_orig, c->d_nbr, c->nbr_k);
  CK_LAUNCH();
  CK(cudaDeviceSynchronize());
  c->bvh = bp;
  c->d_emb = d_emb;
  c->emb_dtype = emb_dtype;
  c->rnorm_ready = false;
  c->planes_ready = false;
  c->query_blocks_per_sm = 0;
  c->cb_ready = true;
  return MT_OK;
}

extern "C" int mt_codebook_grid_info(mt_ctx* c, float* h, int dims[3], int* occupied) {
  if (!c || !c->cb_ready) return set_err(MT_ERR_STATE, "mt_codebook_grid_info: no codebook");
  if (h) *h = c->bvh.cell;
  if (dims) dims[0] = c->bvh.n_leaf, dims[1] = c->bvh.n_l1, dims[2] = c->bvh.n_l2;
  if (occupied) *occupied = c->bvh.n_leaf;
  return MT_OK;
}

extern "C" int mt_codebook_nbr_info(mt_ctx* c, const float** d_nbr, int* k) {
  if (!c || !c->cb_ready) return set_err(MT_ERR_STATE, "mt_codebook_nbr_info: no codebook");
  if (d_nbr) *d_nbr = (const float*)c->d_nbr;
  if (k) *k = c->nbr_k;
  return MT_OK;
}

extern "C" int mt_codebook_rank(mt_ctx* c, int32_t* d_rank, void* stream) {
  if (!c || !c->cb_ready) return set_err(MT_ERR_STATE, "mt_codebook_rank: no codebook");
  if (!d_rank) return set_err(MT_ERR_ARG, "mt_codebook_rank: null output");
  k_key_rank<<<(c->M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->d_keys_sorted, c->M, d_rank);
  CK_LAUNCH();
  return MT_OK;
}

extern "C" int mt_ctx_set_timing_events(mt_ctx* c, void* const* events4) {
  if (!c) return set_err(MT_ERR_ARG, "mt_ctx_set_timing_events: null context");
  for (int k = 0; k < 4; ++k) c->timing[k] = events4 ? (cudaEvent_t)events4[k] : nullptr;
  return MT_OK;
}

extern "C" int mt_ctx_stats(mt_ctx* c, long long* h_out, int reset) {
  if (!c) return set_err(MT_ERR_ARG, "mt_ctx_stats: null context");
  CK(cudaSetDevice(c->device));
  int f[MT_STAT_COUNT];
  CK(cudaMemcpy(f, c->d_flags, sizeof(f), cudaMemcpyDeviceToHost));
  if (h_out)
    for (int k = 0; k < MT_STAT_COUNT; ++k) h_out[k] = f[k];
  if (reset) {
    CK(cudaMemset(c->d_flags, 0, 5 * sizeof(int)));
    CK(cudaMemset(c->d_flags + 7, 0, (MT_STAT_COUNT - 7) * sizeof(int)));
  }
  return MT_OK;
}

// ------------------------------------------------------------------------- mesh (drift test)
static MeshTables mesh_of(mt_ctx* c);
#ifndef MT_VOX_DIV
#define MT_VOX_DIV 4.0   // voxel edge of the drift-test classes = invalid_dist / MT_VOX_DIV ...
#endif
#ifndef MT_VOX_MAX
#define MT_VOX_MAX 48.0e6  // ... unless that needs more voxels than this (4 B each + 2 bits each)
#endif
extern "C" int mt_mesh_upload(mt_ctx* c, const double* h_vertices, long long V, double cell) {
  if (!c || !h_vertices || V <= 0 || !(cell > 0.0)) return set_err(MT_ERR_ARG, "mt_mesh_upload: bad argument");
  CK(cudaSetDevice(c->device));
  const double cell0 = cell;
  double lo[3], hi[3];
  for (int k = 0; k < 3; ++k) lo[k] = DBL_MAX, hi[k] = -DBL_MAX;
  for (long long v = 0; v < V; ++v)
    for (int k = 0; k < 3; ++k) {
      double x = h_vertices[3 * v + k];
      if (!(x == x)) return set_err(MT_ERR_ARG, "mt_mesh_upload: NaN vertex");
      lo[k] = std::min(lo[k], x);
      hi[k] = std::max(hi[k], x);
    }
  MeshGrid g;
  for (;;) {  // grow the cell until the grid fits in 8M cells
    double total = 1;
    for (int k = 0; k < 3; ++k) g.dims[k] = (int)floor((hi[k] - lo[k]) / cell) + 1, total *= g.dims[k];
    if (total <= 8.0e6) break;
    cell *= 1.5;
  }
  g.cell = cell;
  g.inv_cell = 1.0 / cell;
  for (int k = 0; k < 3; ++k) g.org[k] = lo[k], g.orgf[k] = (float)lo[k];
  g.inv_cellf = (float)g.inv_cell;
  const long long ncell = (long long)g.dims[0] * g.dims[1] * g.dims[2];
  std::vector<int> cellv(V), order(V), start(ncell + 1, 0);
  for (long long v = 0; v < V; ++v) {
    int x = mt_mesh_cell(h_vertices[3 * v], g.org[0], g.inv_cell, g.dims[0]);
    int y = mt_mesh_cell(h_vertices[3 * v + 1], g.org[1], g.inv_cell, g.dims[1]);
    int z = mt_mesh_cell(h_vertices[3 * v + 2], g.org[2], g.inv_cell, g.dims[2]);
    cellv[v] = (z * g.dims[1] + y) * g.dims[0] + x;
    start[cellv[v] + 1]++;
  }
  for (long long i = 0; i < ncell; ++i) start[i + 1] += start[i];
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cellv[a] < cellv[b]; });
  std::vector<double> sorted(3 * (size_t)V);
  std::vector<float> sorted32(4 * (size_t)V, 0.f);
  g.coord_max = 0.0;
  for (long long v = 0; v < V; ++v)
    for (int k = 0; k < 3; ++k) {
      const double x = h_vertices[3 * (size_t)order[v] + k];
      sorted[3 * v + k] = x;
      sorted32[4 * v + k] = (float)x;
      g.coord_max = std::max(g.coord_max, fabs(x));
    }
  cudaFree(c->d_mesh_verts);
  cudaFree(c->d_mesh_verts32);
  cudaFree(c->d_mesh_cells);
  cudaFree(c->d_mesh_vox);
  cudaFree(c->d_mesh_vox2);
  c->d_mesh_vox2 = nullptr;
  c->d_mesh_verts = nullptr, c->d_mesh_verts32 = nullptr, c->d_mesh_cells = nullptr, c->d_mesh_vox = nullptr, c->mesh_ready = false;
  memset(&c->vox, 0, sizeof(c->vox));
  CK(cudaMalloc(&c->d_mesh_verts, sizeof(double) * 3 * V));
  CK(cudaMalloc(&c->d_mesh_verts32, sizeof(float4) * V));
  CK(cudaMemcpy(c->d_mesh_verts32, sorted32.data(), sizeof(float4) * V, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_mesh_cells, sizeof(int) * (ncell + 1)));
  CK(cudaMemcpy(c->d_mesh_verts, sorted.data(), sizeof(double) * 3 * V, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_mesh_cells, start.data(), sizeof(int) * (ncell + 1), cudaMemcpyHostToDevice));
  c->mesh = g;
  c->mesh_V = (int)V;
  c->mesh_ready = true;
  // voxel classes for invalid_dist == cell0 (the default, tdn.render.pen.max): voxel edge dist/4,
  // at most 64 M voxels
  {
    const double dist = cell0;
    double v = dist / MT_VOX_DIV;
    MeshVoxels vx;
    memset(&vx, 0, sizeof(vx));
    for (;;) {
      double total = 1;
      for (int k = 0; k < 3; ++k) vx.dims[k] = (int)ceil((hi[k] - lo[k] + 2.0 * (dist + 2.0 * v)) / v) + 1, total *= vx.dims[k];
      if (total <= MT_VOX_MAX) break;
      v *= 1.26;
    }
    for (int k = 0; k < 3; ++k) vx.org[k] = (float)(lo[k] - (dist + 2.0 * v));
    vx.inv_v = (float)(1.0 / v);
    vx.dist = dist;
    const float vf = 1.0f / vx.inv_v;  // the edge the float32 index arithmetic effectively uses
    const size_t total = (size_t)vx.dims[0] * vx.dims[1] * vx.dims[2];
    CK(cudaMalloc(&c->d_mesh_vox, total * sizeof(int)));
    MeshTables T = mesh_of(c);
    T.vox.cls = nullptr;
    // slack: float32 rounding of (x - org) * inv_v, in metres
    const float slack = (float)(4e-7 * (g.coord_max + fabs((double)vx.org[0]) + fabs((double)vx.org[1]) + fabs((double)vx.org[2])) + 1e-9);
    k_mesh_classify<<<(unsigned)((total + 255) / 256), 256>>>(T, vx, vf, slack, c->d_mesh_vox);
    CK_LAUNCH();
    CK(cudaMalloc(&c->d_mesh_vox2, sizeof(unsigned) * (total / 16 + 1)));
    k_mesh_pack2<<<(unsigned)((total / 16 + 256) / 256), 256>>>(c->d_mesh_vox, total, c->d_mesh_vox2);
    CK_LAUNCH();
    CK(cudaDeviceSynchronize());
    vx.cls = c->d_mesh_vox;
    vx.cls2 = c->d_mesh_vox2;
    c->vox = vx;
  }
  return MT_OK;
}

static MeshTables mesh_of(mt_ctx* c) {
  MeshTables T;
  T.verts = c->d_mesh_verts;
  T.verts32 = c->d_mesh_verts32;
  T.vox = c->vox;
  T.cell_start = c->d_mesh_cells;
  T.g = c->mesh;
  T.V = c->mesh_V;
  return T;
}

// weights *= (nearest-vertex distance <= invalid_dist); *num_valid += #valid
__global__ void __launch_bounds__(256) k_prune_aos(MeshTables T, const float* __restrict__ aos, long long n, double dist,
                                                  double* __restrict__ w, int* __restrict__ num_valid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    ok = mesh_within(T, aos[16 * i + 3], aos[16 * i + 7], aos[16 * i + 11], dist);
    if (w) w[i] = ok ? w[i] * 1.0 : w[i] * 0.0;  // NaN weights stay NaN like `weights *= m`
  }
  const int cnt = __syncthreads_count(ok);
  if (threadIdx.x == 0 && cnt && num_valid) atomicAdd(num_valid, cnt);
}

extern "C" int mt_prune_aos(mt_ctx* c, const float* d_poses, long long n, double invalid_dist, double* d_weights,
                            int* d_num_valid, void* stream) {
  if (!c || !c->mesh_ready) return set_err(MT_ERR_STATE, "mt_prune_aos: no mesh uploaded");
  if (n < 0 || (n && !d_poses) || !(invalid_dist >= 0.0)) return set_err(MT_ERR_ARG, "mt_prune_aos: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_num_valid) CK(cudaMemsetAsync(d_num_valid, 0, sizeof(int), st));
  if (!n) return MT_OK;
  k_prune_aos<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mesh_of(c), d_poses, n, invalid_dist, d_weights, d_num_valid);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- block helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// fixed-topology sum over a 256-thread block; result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* s8) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s8[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < MT_CHUNK / 32; ++k) t += s8[k];
  }
  return t;
}

// inclusive scan over a 256-thread block (Kogge-Stone per warp, sequential warp carry)
__device__ __forceinline__ double block_incl_scan_256(double v, double* s8) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) s8[w] = v;
  __syncthreads();
  double carry = 0.0;
#pragma unroll
  for (int k = 0; k < MT_CHUNK / 32; ++k)
    if (k < w) carry += s8[k];
  return carry + v;
}

// The same scan made monotone: float64 rounding can leave a Kogge-Stone prefix of non-negative terms an ulp below
// its predecessor; a running maximum (exact, associative) on top removes that.  Every block that runs this on the
// same inputs gets the same bits.
__device__ __forceinline__ double block_incl_scan_mono_256(double v, double* s8) {
  v = block_incl_scan_256(v, s8);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = fmax(v, t);
  }
  __syncthreads();
  if (lane == 31) s8[w] = v;
  __syncthreads();
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < MT_CHUNK / 32; ++k)
    if (k < w) m = fmax(m, s8[k]);
  return fmax(v, m);
}

// Last block standing: exclusive prefix over `n` chunk sums with one 256-thread block.
// Each thread owns a contiguous run (sequential), thread totals are block-scanned; the
// topology depends only on n, so results are reproducible run to run.
#define MT_SCAN_TILE 4096  // chunk sums staged per pass of the last block (32 KB of shared memory)
__device__ void scan_chunk_sums(const double* part, int n, double* prefix, double* total_out, double* s8) {
  __shared__ double s_tile[MT_SCAN_TILE];
  __shared__ double s_incl[MT_CHUNK];
  constexpr int PER = MT_SCAN_TILE / MT_CHUNK;  // 16 contiguous chunk sums per thread
  double carry = 0.0;
  for (int t0 = 0; t0 < n; t0 += MT_SCAN_TILE) {
    const int cnt = min(MT_SCAN_TILE, n - t0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {  // coalesced, 16 independent loads in flight per thread
      const int j = threadIdx.x + k * MT_CHUNK;
      s_tile[j] = (j < cnt) ? __ldcg(part + t0 + j) : 0.0;
    }
    __syncthreads();
    const int b = threadIdx.x * PER;
    double loc = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) loc += s_tile[b + k];
    const double incl = block_incl_scan_256(loc, s8);
    s_incl[threadIdx.x] = incl;
    __syncthreads();
    // exclusive base of this thread's run == the previous thread's inclusive value, so clamping
    // the running sums to the own inclusive value keeps the whole prefix array monotone
    double run = carry + (threadIdx.x ? s_incl[threadIdx.x - 1] : 0.0);
    const double top = carry + incl;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const double v = s_tile[b + k];
      s_tile[b + k] = fmin(run, top);
      run += v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int j = threadIdx.x + k * MT_CHUNK;
      if (j < cnt) prefix[t0 + j] = s_tile[j];
    }
    carry += s_incl[MT_CHUNK - 1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    prefix[n] = carry;
    *total_out = carry;
  }
}

// fixed-order sum of arr[0..n) with stride `stride` by one 256-thread block (valid in thread 0)
__device__ double block_sum_array_256(const double* arr, int n, int stride, double* s8) {
  double acc = 0.0;
  for (int g0 = 0; g0 < n; g0 += 8 * MT_CHUNK) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = g0 + threadIdx.x + k * MT_CHUNK;
      v[k] = (g < n) ? __ldcg(arr + (size_t)g * stride) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  return block_sum_256(acc, s8);
}

// ------------------------------------------------------------------------- cosine kernels
template <typename T>
struct VecLoad;
template <>
struct VecLoad<float> {
  static constexpr int W = 4;
  __device__ static void ld(const float* p, double* o) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  }
  // the codebook query reads every embedding once per frame: streamed past the resident tables
  __device__ static void ld_stream(const float* p, double* o) {
#if MT_L2_HINTS
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(mt_pol_stream()));
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
#else
    ld(p, o);
#endif
  }
};
template <>
struct VecLoad<double> {
  static constexpr int W = 2;
  __device__ static void ld(const double* p, double* o) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    o[0] = v.x, o[1] = v.y;
  }
  __device__ static void ld_stream(const double* p, double* o) {
#if MT_L2_HINTS
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(o[0]), "=d"(o[1]) : "l"(p), "l"(mt_pol_stream()));
#else
    ld(p, o);
#endif
  }
};

// one warp per row.  q is staged in shared memory as float64.  Traffic: rows*D*sizeof(T).
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) k_cosine_rows(const double* __restrict__ qd, double qnorm_unused,
                                                     const T* __restrict__ E, long long rows, int D,
                                                     double* __restrict__ out, double* __restrict__ out_exp,
                                                     double* __restrict__ out2) {
  extern __shared__ double sq[];
  for (int i = threadIdx.x; i < D; i += blockDim.x) sq[i] = qd[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int W = VecLoad<T>::W;
  const T* e = E + row * (long long)D;
  double dot = 0.0, nn = 0.0, qq = 0.0;
  const int nvec = D / W;
  int v = lane;
  for (; v + 32 * (UNROLL - 1) < nvec; v += 32 * UNROLL) {
    double x[UNROLL][W];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) VecLoad<T>::ld(e + (size_t)(v + 32 * u) * W, x[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < W; ++k) {
        double qv = sq[(v + 32 * u) * W + k];
        dot = fma(x[u][k], qv, dot);
        nn = fma(x[u][k], x[u][k], nn);
      }
  }
  for (; v < nvec; v += 32) {
    double x[W];
    VecLoad<T>::ld(e + (size_t)v * W, x);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      dot = fma(x[k], sq[v * W + k], dot);
      nn = fma(x[k], x[k], nn);
    }
  }
  for (int i = nvec * W + lane; i < D; i += 32) {  // ragged tail
    double x = (double)e[i];
    dot = fma(x, sq[i], dot);
    nn = fma(x, x, nn);
  }
  for (int i = lane; i < D; i += 32) qq = fma(sq[i], sq[i], qq);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    qq += __shfl_xor_sync(0xffffffffu, qq, o);
  }
  if (lane == 0) {
    // torch cosine_similarity: x/max(|x|,eps) . y/max(|y|,eps), eps = 1e-8
    double c = dot / (fmax(sqrt(qq), 1e-8) * fmax(sqrt(nn), 1e-8));
    out[row] = c;
    if (out_exp) out_exp[row] = exp(c);
    if (out2) out2[row] = c;
  }
}

template <typename TQ>
__global__ void k_to_f64(const TQ* in, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

static int launch_cosine(const double* d_q64, const void* E, int dt, long long rows, int D, double* out,
                         double* out_exp, double* out2, cudaStream_t st) {
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  const size_t sh = sizeof(double) * D;
  if (rows == 0) return MT_OK;
  if (dt == MT_DTYPE_F32) {
    if (D % 4) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 4 for float32 rows");
    k_cosine_rows<float, 4><<<grid, 256, sh, st>>>(d_q64, 0.0, (const float*)E, rows, D, out, out_exp, out2);
  } else {
    if (D % 2) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 2 for float64 rows");
    k_cosine_rows<double, 4><<<grid, 256, sh, st>>>(d_q64, 0.0, (const double*)E, rows, D, out, out_exp, out2);
  }
  CK_LAUNCH();
  return MT_OK;
}


// ---- codebook query (static codebook: row norms are precomputed at upload) ----------------------
// row norms max(|E_m|, 1e-8) (the clamp of torch.cosine_similarity), one warp per row
template <typename T>
__global__ void __launch_bounds__(256) k_row_norms(const T* __restrict__ E, int M, int D, double* __restrict__ rnorm) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* e = E + (size_t)row * D;
  double nn = 0.0;
  for (int i = lane; i < D; i += 32) {
    const double x = (double)e[i];
    nn = fma(x, x, nn);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  if (lane == 0) rnorm[row] = fmax(sqrt(nn), 1e-8);
}

// sim[m] = <q, E_m> / (|q| |E_m|), exp(sim[m]).  Persistent warps, four rows per trip: the query
// vector is read from shared memory once per four rows, sixteen 16-byte loads are in flight per
// lane, and the four dot products are reduced with a transposing butterfly (12 shuffles instead
// of 40).  Traffic: M*D*sizeof(T) + 24*M bytes.
template <typename T, typename TQ>
__global__ void __launch_bounds__(256) k_codebook_query(const TQ* __restrict__ q_in, const T* __restrict__ E,
                                                        const double* __restrict__ rnorm, int M, int D,
                                                        double* __restrict__ sim, double* __restrict__ esim,
                                                        double* __restrict__ sim2) {
  extern __shared__ double sq[];
  __shared__ double s8q[8];
  __shared__ double s_qn;
  {  // every block stages the query as float64 and computes its clamped norm (D values: cheaper than a launch)
    double acc = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      const double x = (double)q_in[i];
      sq[i] = x;
      acc = fma(x, x, acc);
    }
    const double t = block_sum_256(acc, s8q);
    if (threadIdx.x == 0) s_qn = fmax(sqrt(t), 1e-8);
    __syncthreads();
  }
  constexpr int W = VecLoad<T>::W;
  const int lane = threadIdx.x & 31;
  const int nvec = D / W;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  const double qn = s_qn;
  for (int r0 = 4 * gw; r0 < M; r0 += 4 * nw) {
    const T* e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = E + (size_t)min(r0 + j, M - 1) * D;
    double dot[4] = {0.0, 0.0, 0.0, 0.0};
    int v = lane;
    for (; v + 96 < nvec; v += 128) {  // 4 vectors x 4 rows in flight
      double x[4][4][W];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) VecLoad<T>::ld_stream(e[j] + (size_t)(v + 32 * u) * W, x[u][j]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < W; ++k) {
          const double qv = sq[(v + 32 * u) * W + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) dot[j] = fma(x[u][j][k], qv, dot[j]);
        }
    }
    for (; v < nvec; v += 32) {
      double x[4][W];
#pragma unroll
      for (int j = 0; j < 4; ++j) VecLoad<T>::ld_stream(e[j] + (size_t)v * W, x[j]);
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const double qv = sq[v * W + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) dot[j] = fma(x[j][k], qv, dot[j]);
      }
    }
    for (int i = nvec * W + lane; i < D; i += 32)  // ragged tail
#pragma unroll
      for (int j = 0; j < 4; ++j) dot[j] = fma((double)e[j][i], sq[i], dot[j]);
    // transposing butterfly: 4 values -> 2 -> 1 per lane, then a plain reduction over 8 lanes
    {
      const bool hi = lane & 16;
      const double s0 = hi ? dot[0] : dot[2], s1 = hi ? dot[1] : dot[3];
      const double r0v = __shfl_xor_sync(0xffffffffu, s0, 16), r1v = __shfl_xor_sync(0xffffffffu, s1, 16);
      double a = (hi ? dot[2] : dot[0]) + r0v, b = (hi ? dot[3] : dot[1]) + r1v;  // lanes<16: rows 0,1; lanes>=16: rows 2,3
      const bool h8 = lane & 8;
      const double snd = h8 ? a : b;
      const double rcv = __shfl_xor_sync(0xffffffffu, snd, 8);
      double c = (h8 ? b : a) + rcv;  // row = 2*(lane>=16) + (lane&8 ? 1 : 0)
      c += __shfl_xor_sync(0xffffffffu, c, 4);
      c += __shfl_xor_sync(0xffffffffu, c, 2);
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      if ((lane & 7) == 0) {
        const int row = r0 + 2 * (lane >> 4) + ((lane >> 3) & 1);
        if (row < M) {
          const double cs = c / (qn * __ldg(rnorm + row));
          sim[row] = cs;
          esim[row] = exp(cs);
          if (sim2) sim2[row] = cs;
        }
      }
    }
  }
}

static int stage_query(mt_ctx* c, const void* d_q, int q_dtype, int D, double** out, cudaStream_t st) {
  if (D > MT_MAX_D) return set_err(MT_ERR_ARG, "cosine: D exceeds MT_MAX_D (6144)");
  if (q_dtype == MT_DTYPE_F32)
    k_to_f64<float><<<(D + 255) / 256, 256, 0, st>>>((const float*)d_q, D, c->d_q64);
  else if (q_dtype == MT_DTYPE_F64)
    k_to_f64<double><<<(D + 255) / 256, 256, 0, st>>>((const double*)d_q, D, c->d_q64);
  else
    return set_err(MT_ERR_ARG, "cosine: bad query dtype");
  CK_LAUNCH();
  *out = c->d_q64;
  return MT_OK;
}

extern "C" int mt_codebook_query(mt_ctx* c, const void* d_q, int q_dtype, double* d_sim_out, void* stream) {
  if (!c || !c->cb_ready || !c->d_emb) return set_err(MT_ERR_STATE, "mt_codebook_query: no codebook");
  if (!d_q) return set_err(MT_ERR_ARG, "mt_codebook_query: null query");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = c->D, M = c->M;
  if (D > MT_MAX_D) return set_err(MT_ERR_ARG, "cosine: D exceeds MT_MAX_D (6144)");
  if (q_dtype != MT_DTYPE_F32 && q_dtype != MT_DTYPE_F64) return set_err(MT_ERR_ARG, "cosine: bad query dtype");
  const bool e32 = c->emb_dtype == MT_DTYPE_F32, q32 = q_dtype == MT_DTYPE_F32;
  if (e32 && D % 4) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 4 for float32 rows");
  if (!e32 && D % 2) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 2 for float64 rows");
  if (!c->rnorm_ready) {  // the embeddings are static between uploads: their norms are computed once
    if (e32)
      k_row_norms<float><<<(M + 7) / 8, 256, 0, st>>>((const float*)c->d_emb, M, D, c->d_rnorm);
    else
      k_row_norms<double><<<(M + 7) / 8, 256, 0, st>>>((const double*)c->d_emb, M, D, c->d_rnorm);
    CK_LAUNCH();
    c->rnorm_ready = true;
  }
  const size_t sh = sizeof(double) * D;
  if (!c->query_blocks_per_sm) {  // persistent grid = what is resident at once (no tail wave)
    int occ = 0;
    if (e32)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_codebook_query<float, double>, 256, sh);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_codebook_query<double, double>, 256, sh);
    c->query_blocks_per_sm = occ > 0 ? occ : 4;
  }
  // persistent warps take 4 rows per trip: size the grid so that every warp makes the same number of trips
  const int resident = c->sm_count * c->query_blocks_per_sm;
  const int trips_total = (M + 3) / 4;
  int grid = (trips_total + 7) / 8;
  for (int k = 2; grid > resident; ++k) grid = ((trips_total + k - 1) / k + 7) / 8;
  const float* Ef = (const float*)c->d_emb;
  const double* Ed = (const double*)c->d_emb;
  if (e32 && q32) k_codebook_query<float, float><<<grid, 256, sh, st>>>((const float*)d_q, Ef, c->d_rnorm, M, D, c->d_sim, c->d_esim, d_sim_out);
  else if (e32) k_codebook_query<float, double><<<grid, 256, sh, st>>>((const double*)d_q, Ef, c->d_rnorm, M, D, c->d_sim, c->d_esim, d_sim_out);
  else if (q32) k_codebook_query<double, float><<<grid, 256, sh, st>>>((const float*)d_q, Ed, c->d_rnorm, M, D, c->d_sim, c->d_esim, d_sim_out);
  else k_codebook_query<double, double><<<grid, 256, sh, st>>>((const double*)d_q, Ed, c->d_rnorm, M, D, c->d_sim, c->d_esim, d_sim_out);
  CK_LAUNCH();
  return MT_OK;
}

extern "C" int mt_cosine_rows(mt_ctx* c, const void* d_q, int q_dtype, const void* d_t, int t_dtype, long long rows,
                              int D, double* d_out, void* stream) {
  if (!c || !d_q || (!d_t && rows) || !d_out || D <= 0 || rows < 0) return set_err(MT_ERR_ARG, "mt_cosine_rows: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* q64;
  int r = stage_query(c, d_q, q_dtype, D, &q64, st);
  if (r) return r;
  return launch_cosine(q64, d_t, t_dtype, rows, D, d_out, nullptr, nullptr, st);
}

// Q x rows cosine in fp32: 64x64 output tile per block, K-step 16, register 4x4 micro-tile.
// Norms are accumulated alongside.  (SIMT version; the tcgen05 path is in DESIGN.md "next".)
__global__ void __launch_bounds__(256) k_cosine_batched(const float* __restrict__ Q, int nq,
                                                        const float* __restrict__ T, long long rows, int D,
                                                        float* __restrict__ out) {
  __shared__ float sQ[16][64 + 1];
  __shared__ float sT[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.x * 64;
  const int q0 = blockIdx.y * 64;
  float acc[4][4] = {};
  float nq2[4] = {}, nt2[4] = {};
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int r = i >> 4, k = i & 15;
      sQ[k][r] = (q0 + r < nq && k0 + k < D) ? Q[(size_t)(q0 + r) * D + k0 + k] : 0.f;
      sT[k][r] = (m0 + r < rows && k0 + k < D) ? T[(size_t)(m0 + r) * D + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sQ[k][ty * 4 + i], b[i] = sT[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        nq2[i] = fmaf(a[i], a[i], nq2[i]);
        nt2[i] = fmaf(b[i], b[i], nt2[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int q = q0 + ty * 4 + i;
      long long m = m0 + tx * 4 + j;
      if (q < nq && m < rows) out[(size_t)q * rows + m] = acc[i][j] / (fmaxf(sqrtf(nq2[i]), 1e-8f) * fmaxf(sqrtf(nt2[j]), 1e-8f));
    }
}

extern "C" int mt_cosine_batched(mt_ctx* c, const float* d_Q, int nq, const float* d_T, long long rows, int D,
                                 float* d_out, void* stream) {
  if (!d_Q || !d_T || !d_out || nq <= 0 || rows <= 0 || D <= 0) return set_err(MT_ERR_ARG, "mt_cosine_batched: bad argument");
  dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((nq + 63) / 64));
  k_cosine_batched<<<grid, 256, 0, (cudaStream_t)stream>>>(d_Q, nq, d_T, rows, D, d_out);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- softmax (f64)
__global__ void __launch_bounds__(256) k_minmax(const double* __restrict__ x, long long n, double* part,
                                                unsigned int* ticket, double* scal) {
  __shared__ double smx[8], smn[8];
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double mx = -DBL_MAX, mn = DBL_MAX;
  int nan = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i];
    nan |= (v != v);
    mx = fmax(mx, v);  // fmax/fmin ignore NaN; NaN is tracked separately
    mn = fmin(mn, v);
  }
  nan = __syncthreads_or(nan);
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx, smn[threadIdx.x >> 5] = mn;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) mx = fmax(mx, smx[k]), mn = fmin(mn, smn[k]);
    part[2 * blockIdx.x] = nan ? qnan : mx;
    part[2 * blockIdx.x + 1] = nan ? qnan : mn;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = -DBL_MAX, b = DBL_MAX;
    bool any_nan = false;
    for (unsigned k = 0; k < gridDim.x; ++k) {
      double pa = __ldcg(part + 2 * k), pb = __ldcg(part + 2 * k + 1);
      any_nan |= (pa != pa);
      a = fmax(a, pa);
      b = fmin(b, pb);
    }
    scal[1] = any_nan ? qnan : a;  // torch: max()/min() propagate NaN -> softmax of NaN
    scal[2] = any_nan ? qnan : b;
    *ticket = 0;
  }
}

__global__ void __launch_bounds__(256) k_expsum(const double* __restrict__ x, long long n, double* part,
                                                unsigned int* ticket, double* scal) {
  __shared__ double s8[8];
  const double mx = scal[1];
  double acc = 0.0;
  // contiguous run per block so the summation tree is a function of (n, grid) only
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per, e = (b + per < n) ? b + per : n;
  for (long long i = b + threadIdx.x; i < e; i += blockDim.x) acc += exp(x[i] - mx);
  double t = block_sum_256(acc, s8);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    part[blockIdx.x] = t;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) s += __ldcg(part + k);
    scal[3] = s;
    *ticket = 0;
  }
}

__global__ void k_softmax_apply(const double* __restrict__ x, long long n, const double* __restrict__ scal,
                                double* __restrict__ out) {
  const double mx = scal[1], mn = scal[2], den = scal[3];
  // torch.isclose(max-min, 0): |d| <= atol(1e-8) + rtol*|0|; NaN is never close -> softmax
  const bool skip = fabs(mx - mn) <= 1e-8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i];
    out[i] = skip ? v : exp(v - mx) / den;
  }
}

extern "C" int mt_softmax_f64(mt_ctx* c, const double* d_in, long long n, double* d_out, void* stream) {
  if (!c || !d_in || !d_out || n <= 0) return set_err(MT_ERR_ARG, "mt_softmax_f64: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)std::min<long long>((n + 255) / 256, std::min<long long>(c->chunk_cap, 1184));
  k_minmax<<<grid, 256, 0, st>>>(d_in, n, c->d_rm_part, c->d_ticket + 1, c->d_scal);
  CK_LAUNCH();
  k_expsum<<<grid, 256, 0, st>>>(d_in, n, c->d_part, c->d_ticket + 2, c->d_scal);
  CK_LAUNCH();
  k_softmax_apply<<<grid, 256, 0, st>>>(d_in, n, c->d_scal, d_out);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- layout
__device__ __forceinline__ void load_pose(const float4* __restrict__ soa, long long stride, long long i, float P[3][4]) {
  float4 a = soa[i], b = soa[stride + i], c = soa[2 * stride + i];
  P[0][0] = a.x, P[0][1] = a.y, P[0][2] = a.z, P[0][3] = a.w;
  P[1][0] = b.x, P[1][1] = b.y, P[1][2] = b.z, P[1][3] = b.w;
  P[2][0] = c.x, P[2][1] = c.y, P[2][2] = c.z, P[2][3] = c.w;
}
__device__ __forceinline__ void store_pose(float4* __restrict__ soa, long long stride, long long i, const float P[3][4]) {
  soa[i] = make_float4(P[0][0], P[0][1], P[0][2], P[0][3]);
  soa[stride + i] = make_float4(P[1][0], P[1][1], P[1][2], P[1][3]);
  soa[2 * stride + i] = make_float4(P[2][0], P[2][1], P[2][2], P[2][3]);
}

// the step kernels touch every particle once per launch: streamed (L2 evict_first) so that the tables stay resident
__device__ __forceinline__ void load_pose_stream(const float4* __restrict__ soa, long long stride, long long i, float P[3][4]) {
  float4 a = mt_lds(soa + i), b = mt_lds(soa + stride + i), c = mt_lds(soa + 2 * stride + i);
  P[0][0] = a.x, P[0][1] = a.y, P[0][2] = a.z, P[0][3] = a.w;
  P[1][0] = b.x, P[1][1] = b.y, P[1][2] = b.z, P[1][3] = b.w;
  P[2][0] = c.x, P[2][1] = c.y, P[2][2] = c.z, P[2][3] = c.w;
}
__device__ __forceinline__ void store_pose_stream(float4* __restrict__ soa, long long stride, long long i, const float P[3][4]) {
  mt_sts(soa + i, make_float4(P[0][0], P[0][1], P[0][2], P[0][3]));
  mt_sts(soa + stride + i, make_float4(P[1][0], P[1][1], P[1][2], P[1][3]));
  mt_sts(soa + 2 * stride + i, make_float4(P[2][0], P[2][1], P[2][2], P[2][3]));
}

__global__ void k_aos_to_soa(const float4* __restrict__ aos, long long n, float4* __restrict__ soa, long long stride) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  soa[i] = aos[4 * i];
  soa[stride + i] = aos[4 * i + 1];
  soa[2 * stride + i] = aos[4 * i + 2];
}
__global__ void k_soa_to_aos(const float4* __restrict__ soa, long long stride, long long n, float4* __restrict__ aos) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  aos[4 * i] = soa[i];
  aos[4 * i + 1] = soa[stride + i];
  aos[4 * i + 2] = soa[2 * stride + i];
  aos[4 * i + 3] = make_float4(0.f, 0.f, 0.f, 1.f);
}
extern "C" int mt_aos_to_soa(const float* d_aos, long long n, float* d_soa, long long stride, void* stream) {
  if (n < 0 || stride < n || (n && (!d_aos || !d_soa))) return set_err(MT_ERR_ARG, "mt_aos_to_soa: bad argument");
  if (!n) return MT_OK;
  k_aos_to_soa<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_aos, n, (float4*)d_soa, stride);
  CK_LAUNCH();
  return MT_OK;
}
extern "C" int mt_soa_to_aos(const float* d_soa, long long stride, long long n, float* d_aos, void* stream) {
  if (n < 0 || stride < n || (n && (!d_aos || !d_soa))) return set_err(MT_ERR_ARG, "mt_soa_to_aos: bad argument");
  if (!n) return MT_OK;
  k_soa_to_aos<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_soa, stride, n, (float4*)d_aos);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- exact 1-NN
// search algorithms: mt_nn.cuh.  Standalone index search (SE3_NN on explicit keys).
__global__ void __launch_bounds__(MT_NN_BLOCK) k_nn_grid(NNTables T, const float* __restrict__ keys, long long n,
                                                         const int* __restrict__ hint, int* __restrict__ idx,
                                                         int* fallbacks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float q[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = keys[6 * i + k];
  }
  const int r = nn_assign(T, i < n, q, (hint && i < n) ? hint[i] : -1, fallbacks);
  if (i < n) idx[i] = r;
}

// exhaustive search: codebook keys staged through shared memory in original order so that
// a strict '<' keeps the lowest index on ties.
__global__ void __launch_bounds__(256) k_nn_brute(const float4* __restrict__ keys_orig, int M,
                                                  const float* __restrict__ keys, long long n, int* __restrict__ idx) {
  __shared__ float4 sk[2 * 512];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float q[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = keys[6 * i + k];
  }
  float best_d = FLT_MAX;
  int best_i = 0;
  for (int m0 = 0; m0 < M; m0 += 512) {
    int cnt = min(512, M - m0);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * cnt; t += blockDim.x) sk[t] = keys_orig[2 * (size_t)m0 + t];
    __syncthreads();
    for (int p = 0; p < cnt; ++p) {
      float4 a = sk[2 * p], b = sk[2 * p + 1];
      float k[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
      float d = mt_key_dist(q, k);
      if (d < best_d) best_d = d, best_i = m0 + p;
    }
  }
  if (i < n) idx[i] = best_i;
}

static NNTables tables_of(mt_ctx* c) {
  NNTables T;
  T.keys_orig = c->d_keys_orig;
  T.keys_sorted = c->d_keys_sorted;
  T.bvh_leaf = c->d_bvh;
  T.bvh_l1 = c->d_bvh + 3 * (size_t)c->bvh.n_leaf;
  T.bvh_l2 = c->d_bvh + 3 * ((size_t)c->bvh.n_leaf + c->bvh.n_l1);
  T.nbr = c->d_nbr;
  T.b = c->bvh;
  T.M = c->M;
  T.K = c->nbr_k;
  return T;
}

// wt / wr: the float32 factors (1 - w) and w of R3_SE3 (tactile_tree.py:73-77); the codebook's own keys use w = 0.01
__global__ void k_se3_keys(const float4* __restrict__ soa, long long stride, long long n, float* __restrict__ keys, float wt, float wr) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float P[3][4], lg[3];
  load_pose(soa, stride, i, P);
  mt_so3_log(P, lg);
  keys[6 * i] = wt * P[0][3], keys[6 * i + 1] = wt * P[1][3], keys[6 * i + 2] = wt * P[2][3];
  keys[6 * i + 3] = wr * lg[0], keys[6 * i + 4] = wr * lg[1], keys[6 * i + 5] = wr * lg[2];
}

extern "C" int mt_se3_keys_w(const float* d_soa, long long stride, long long n, double w, float* d_keys, void* stream) {
  if (n < 0 || stride < n || (n && (!d_soa || !d_keys)) || !(w == w)) return set_err(MT_ERR_ARG, "mt_se3_keys: bad argument");
  if (!n) return MT_OK;
  // torch multiplies a float32 tensor by the Python scalars (1.0 - w) and w: both rounded to float32 first
  k_se3_keys<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_soa, stride, n, d_keys, (float)(1.0 - w), (float)w);
  CK_LAUNCH();
  return MT_OK;
}
extern "C" int mt_se3_keys(const float* d_soa, long long stride, long long n, float* d_keys, void* stream) {
  return mt_se3_keys_w(d_soa, stride, n, 0.01, d_keys, stream);
}

// SE3_NN with nn > 1 (tactile_tree.py:43-52): the k <= 64 nearest codebook keys of every query key, ascending
// (distance, index) like kneighbors(); exhaustive (every query streams the codebook).
extern "C" int mt_nn_topk(mt_ctx* c, const float* d_keys, long long n, int k, int32_t* d_idx, void* stream) {
  if (!c || !c->cb_ready) return set_err(MT_ERR_STATE, "mt_nn_topk: no codebook");
  if (n < 0 || k < 1 || k > MT_NBR_K || k > c->M || (n && (!d_keys || !d_idx)) || n > 0x7fffffffll / 8) return set_err(MT_ERR_ARG, "mt_nn_topk: bad argument (1 <= k <= min(64, M))");
  if (!n) return MT_OK;
  k_build_nbr<2><<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(c->d_keys_orig, c->M, nullptr, nullptr, d_keys, (int)n, k, d_idx);
  CK_LAUNCH();
  return MT_OK;
}

extern "C" int mt_nn_assign(mt_ctx* c, const float* d_keys, long long n, const int32_t* d_hint, int mode,
                            int32_t* d_idx, void* stream) {
  if (!c || !c->cb_ready) return set_err(MT_ERR_STATE, "mt_nn_assign: no codebook");
  if (n < 0 || (n && (!d_keys || !d_idx))) return set_err(MT_ERR_ARG, "mt_nn_assign: bad argument");
  if (!n) return MT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1)
    k_nn_brute<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->d_keys_orig, c->M, d_keys, n, d_idx);
  else
    k_nn_grid<<<(unsigned)((n + MT_NN_BLOCK - 1) / MT_NN_BLOCK), MT_NN_BLOCK, 0, st>>>(tables_of(c), d_keys, n, d_hint, d_idx, c->d_flags + 3);
  CK_LAUNCH();
  return MT_OK;
}

__global__ void k_gather_rows_f32(const float* __restrict__ table, const int* __restrict__ idx, long long n,
                                  int row_floats, float* __restrict__ out) {
  // one thread per float4 of output
  const int v = row_floats / 4;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * v) return;
  long long i = t / v;
  int k = (int)(t - i * v);
  reinterpret_cast<float4*>(out)[t] = __ldg(reinterpret_cast<const float4*>(table) + (size_t)idx[i] * v + k);
}
extern "C" int mt_gather_rows_f32(const float* d_table, const int32_t* d_idx, long long n, int row_floats,
                                  float* d_out, void* stream) {
  if (n < 0 || row_floats <= 0 || row_floats % 4 || (n && (!d_table || !d_idx || !d_out)))
    return set_err(MT_ERR_ARG, "mt_gather_rows_f32: bad argument");
  if (!n) return MT_OK;
  long long tot = n * (row_floats / 4);
  k_gather_rows_f32<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_table, d_idx, n, row_floats, d_out);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- motion
struct Affine {
  float m[3][4];
};
static Affine affine_from_host16(const float* h) {
  Affine a;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) a.m[i][j] = h[4 * i + j];
  return a;
}

__device__ __forceinline__ void draw_or_load_noise(const float* __restrict__ tn, const float* __restrict__ rot,
                                                   long long i, float sig_t, float sig_r, uint64_t seed,
                                                   uint64_t step, uint64_t gid, float t[3], float r[3]) {
  if (tn) {
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = __ldg(tn + 3 * i + k), r[k] = __ldg(rot + 3 * i + k);
  } else {
    mt_motion_normals(seed, step, gid, t, r);
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] *= sig_t, r[k] *= sig_r;
  }
}

__device__ __forceinline__ void apply_motion(const float P[3][4], const Affine& odom, const float t[3], const float r[3], float out[3][4], int euler_mode = 0, int fast = 0) {
  float Tn[3][4], G[3][4];
  if (euler_mode)
    mt_noise_affine_extrinsic(t, r, Tn);
  else
    mt_noise_affine(t, r, Tn, fast);
  mt_compose(odom.m, Tn, G);  // noisyOdom = odom @ Tn   (particle_filter.py:345)
  mt_compose(P, G, out);      // pose @ noisyOdom         (particle_filter.py:374)
}

__global__ void __launch_bounds__(256) k_motion(const float4* __restrict__ in, float4* __restrict__ out,
                                                long long stride, long long n, Affine odom,
                                                const float* __restrict__ tn, const float* __restrict__ rot,
                                                float sig_t, float sig_r, uint64_t seed, uint64_t step,
                                                uint64_t first_gid, int* invalid, int euler_mode) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float P[3][4], t[3], r[3], O[3][4];
  load_pose(in, stride, i, P);
  draw_or_load_noise(tn, rot, i, sig_t, sig_r, seed, step, first_gid + (uint64_t)i, t, r);
  apply_motion(P, odom, t, r, O, euler_mode, tn == nullptr);
  store_pose(out, stride, i, O);
  if (invalid && mt_pose_invalid(O)) atomicAdd(invalid, 1);
}

extern "C" int mt_motion(const float* d_in, float* d_out, long long stride, long long n, const float* h_odom,
                         const float* d_tn, const float* d_rot, float sig_t, float sig_r, uint64_t seed,
                         uint64_t step, uint64_t first_gid, int* d_invalid, int euler_mode, void* stream) {
  if (n < 0 || stride < n || !h_odom || (n && (!d_in || !d_out)) || ((d_tn == nullptr) != (d_rot == nullptr)))
    return set_err(MT_ERR_ARG, "mt_motion: bad argument");
  if (!n) return MT_OK;
  k_motion<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)d_in, (float4*)d_out, stride, n, affine_from_host16(h_odom), d_tn, d_rot, sig_t, sig_r, seed, step,
      first_gid, d_invalid, euler_mode);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- rmse
__device__ __forceinline__ void rmse_terms(const Affine& gt, const float P[3][4], double& et2, double& ang2) {
  float dx = gt.m[0][3] - P[0][3], dy = gt.m[1][3] - P[1][3], dz = gt.m[2][3] - P[2][3];
  float e = sqrtf(dx * dx + dy * dy + dz * dz);  // torch.norm then **2 (particle_filter.py:488-492)
  et2 = (double)(e * e);
  float a = mt_rot_err_deg(gt.m, P);
  ang2 = (double)(a * a);
}

// finalise by one thread of the last block: fixed-order sum of the chunk partials
__device__ void rmse_finalize(const double* part, int nch, long long n, float* out2) {
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nch; ++k) a += __ldcg(part + 2 * k), b += __ldcg(part + 2 * k + 1);
  out2[0] = (float)sqrt(a / (double)n);
  out2[1] = (float)sqrt(b / (double)n);
}

__global__ void __launch_bounds__(256) k_rmse(const float4* __restrict__ soa, long long stride, long long n, Affine gt,
                                              double* part, unsigned int* ticket, float* out2) {
  __shared__ double s8[8];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double et2 = 0.0, ang2 = 0.0;
  if (i < n) {
    float P[3][4];
    load_pose(soa, stride, i, P);
    rmse_terms(gt, P, et2, ang2);
  }
  double a = block_sum_256(et2, s8);
  double b = block_sum_256(ang2, s8);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    rmse_finalize(part, gridDim.x, n, out2);
    *ticket = 0;
  }
}

extern "C" int mt_rmse(mt_ctx* c, const float* d_soa, long long stride, long long n, const float* h_gt, float* d_out2,
                       void* stream) {
  if (!c || !d_soa || !h_gt || !d_out2 || n <= 0 || stride < n) return set_err(MT_ERR_ARG, "mt_rmse: bad argument");
  if ((long long)nchunks_of(n) > c->chunk_cap) return set_err(MT_ERR_CAPACITY, "mt_rmse: n exceeds context capacity");
  k_rmse<<<(unsigned)nchunks_of(n), 256, 0, (cudaStream_t)stream>>>((const float4*)d_soa, stride, n, affine_from_host16(h_gt),
                                                                    c->d_rm_part, c->d_ticket + 3, d_out2);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- gathers
__global__ void k_gather_soa(const float4* __restrict__ in, long long sin, const int* __restrict__ anc, long long n,
                             float4* __restrict__ out, long long sout) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = anc[i];
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  out[i] = a >= 0 ? in[a] : z;  // unfilled slots stay zero like the reference (particle_filter.py:290)
  out[sout + i] = a >= 0 ? in[sin + a] : z;
  out[2 * sout + i] = a >= 0 ? in[2 * sin + a] : z;
}
extern "C" int mt_gather_soa(const float* d_in, long long sin, const int32_t* d_anc, long long n, float* d_out,
                             long long sout, void* stream) {
  if (n < 0 || (n && (!d_in || !d_anc || !d_out)) || sout < n) return set_err(MT_ERR_ARG, "mt_gather_soa: bad argument");
  if (!n) return MT_OK;
  k_gather_soa<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_in, sin, d_anc, n, (float4*)d_out, sout);
  CK_LAUNCH();
  return MT_OK;
}
__global__ void k_gather_f64(const double* __restrict__ in, const int* __restrict__ anc, long long n, double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = anc[i] >= 0 ? in[anc[i]] : 0.0;
}
extern "C" int mt_gather_f64(const double* d_in, const int32_t* d_anc, long long n, double* d_out, void* stream) {
  if (n < 0 || (n && (!d_in || !d_anc || !d_out))) return set_err(MT_ERR_ARG, "mt_gather_f64: bad argument");
  if (!n) return MT_OK;
  k_gather_f64<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_anc, n, d_out);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- fused step
// MT_TRACE builds (diagnostics, scripts/step_trace.py): every step kernel records the earliest start and the latest end
// of its blocks (%globaltimer, ns) in words [8 + 2k, 8 + 2k + 1] of the context's debug buffer -- k = 0 k_step_a,
// 1 k_step_meshq, 2 k_step_meshq2, 3 k_step_nnq, 4 k_step_bw -- and k_step_bw its phase boundaries in words 24..28.
// mt_trace_read returns the buffer.  Off in the shipped library (no instructions emitted).
#ifndef MT_TRACE
#define MT_TRACE 0
#endif
__device__ __forceinline__ unsigned long long mt_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#if MT_TRACE
#define MT_TRACE_BEGIN(buf, k) if (threadIdx.x == 0) atomicMin((buf) + 8 + 2 * (k), mt_now());
#define MT_TRACE_END(buf, k) if ((threadIdx.x & 31) == 0) atomicMax((buf) + 9 + 2 * (k), mt_now());
#define MT_TRACE_MAX(buf, w) if (threadIdx.x == 0) atomicMax((buf) + (w), mt_now());
#define MT_TRACE_MIN(buf, w) if (threadIdx.x == 0) atomicMin((buf) + (w), mt_now());
#else
#define MT_TRACE_BEGIN(buf, k)
#define MT_TRACE_END(buf, k)
#define MT_TRACE_MAX(buf, w)
#define MT_TRACE_MIN(buf, w)
#endif
struct StepDev {
  float4* soa_cur;
  float4* soa_next;
  long long stride;
  int* nn_cur;
  int* nn_next;
  int* anc;
  long long n;
  Affine odom;
  const float* tn;
  const float* rot;
  float sig_t, sig_r;
  uint64_t seed, step, first_gid;
  const double* wtab;  // exp(sim) or sim
  const double* wsrc;  // explicit weights (resample-only path) or NULL
  float u;
  int has_gt;
  Affine gt;
  float* rmse2;
  int rank, world;
  long long n_global;
  const double* shard_sums;
  Xchg* const* peers;          // sharded fused step: peer exchange buffers (nullptr otherwise)
  unsigned long long xseq;     // sequence number of this step's exchange (same on every rank)
  unsigned long long* xdbg;    // timestamps of the exchange (diagnostics)
  long long* n_out;
  const long long* n_in;  // device-resident particle count (nullable): overrides n
  double prune_dist;      // > 0: drift test against the mesh (remove_invalid_particles)
  const float4* cb_poses; // (M,4,4) codebook poses for the all-drifted re-projection (nullable)
  // scratch
  double* part;
  double* prefix;
  double* rm_part;
  double* wpart;
  double* wrm;
  int* wcnt;
  float4* srec;         // search record of queue entry e, 2 float4: k0..k3 | k4, k5, best_d, best_i bits
  int* queue;
  int* queue2;          // drift tests that need the grid search (k_step_meshq -> k_step_meshq2)
  long long queue_cap;  // entries in `queue` (searches grow from the front, deferred drift tests from the back)
  unsigned int* qctl;   // [0] searches queued, [1] queue head, [2] drift tests left for the grid search, [3] drift tests queued,
                        // [4] scans to continue
  double* scal;
  unsigned int* ticket;
  int* flags;
  int nchunks;
};

// particle count seen by the step kernels: the device-resident count when there is one (sharded runs: the host only
// knows the capacity), clamped to what the grids cover; a count beyond the buffers raises MT_STAT_OVERFLOW
__device__ __forceinline__ long long step_count(const StepDev& p) {
  if (!p.n_in) return p.n;
  long long n = *p.n_in;
  if (n > p.stride) {
    n = p.stride;
    p.flags[0] = 1;
  }
  return n < 0 ? 0 : n;
}

// The measurement half of a step is three kernels (mt_step_a launches them back to back):
//
//  k_step_a     one particle per thread, 64-thread blocks, no block-level synchronisation:
//               motion, SE(3) key, drift test, hint-graph search.  Conclusive searches (~99 %)
//               store the match; the others store their best candidate and append the
//               particle to a queue.  Per-warp RMSE / on-surface partials.
//               HBM per particle: read 48 B pose + 4 B hint (+ 24 B noise when supplied),
//               write 48 B + 4 B.  Keys / neighbour lists / mesh grid are L2-resident.
//  k_step_nnq   the queue, one warp per entry (box-hierarchy search, mt_nn.cuh): the long-tailed work is
//               spread over the whole GPU instead of stalling the warp that found it.
//  k_step_sums  weight lookup w = table[match] (4 B / particle), deterministic float64 chunk
//               sums, and in the last block the chunk prefix for kernel B, the RMSE and the
//               drift flag.
#ifndef MT_A_BLOCK
#define MT_A_BLOCK 64
#endif
#ifndef MT_A_MINBLOCKS
#define MT_A_MINBLOCKS 21  // 48 registers: 42 warps per SM (measured: 16 -> 18 -> 21 blocks = 92.8 -> 91.0 -> 87.0 us)
#endif
#ifndef MT_MESH_DEFER
#define MT_MESH_DEFER 1  // undecided voxels of the drift test go to the queue instead of stalling their warp
#endif
// queue entries: particle index | what is left to do for it
#define MT_Q_NN 0x20000000    // the hint-graph search was not conclusive: box-hierarchy search
#define MT_Q_MESH 0x40000000  // the voxel class was "undecided": vertex search of the drift test
#define MT_Q_MASKED 0x08000000  // the particle is already known to be masked (off the mesh / invalid pose)
#define MT_Q_INDEX 0x07ffffff
// (Cutting the hint scan after 16 / 24 list entries and continuing the rest in a second, compacted pass was measured
// slower -- 105 / 102 us against 99 us for the whole scan: 42 % of the scans need more than 16 entries -- and is gone.)
// (Splitting the sweep into a streaming kernel -- motion, voxel class, key, partials; 48 registers -- and a search kernel
// that starts from a 32-byte key record -- 40 registers, 50 warps per SM -- was measured much slower: 116 us on the
// spread cloud and 162 us on the converged one against 91 us fused.  The search is a chain of dependent list loads;
// in the fused kernel the arithmetic of other warps fills those waits, in a search-only kernel nothing does.)
// (A persistent form of this kernel -- grid = resident blocks, grid-stride tiles, the next tile's hint / key / list head
// prefetched one tile ahead -- was measured 10 % slower, 103 vs 93 us: the hardware block scheduler balances the
// long-tailed scans better than a static tile assignment.)
// (Letting a block walk over 2 / 4 / 8 consecutive tiles with the next tile's pose + hint staged in shared memory by
// cp.async -- to take the DRAM wait of the 52 input bytes off the warps' critical path -- was measured slower still:
// 113 / 129 / 157 us against 91 us.  Every form in which a thread handles more than one particle lost; what the kernel
// needs is more independent warps, which is what the register count below buys.)
// (Re-dealing the hint scans inside the block by their expected length -- a counting sort of the block's jobs on 64
// logarithmic d_h buckets through shared memory, warp w scanning the w-th quantile and finishing the jobs it scanned:
// on the CPU model of this workload it cuts the trips a warp pays for from 12.7 to 7.9 (blocks of 128) -- was measured
// slower in every shape: 105 / 112 / 110 us for blocks of 128 / 256 / 64 against 91 us.  Three block barriers in the
// middle of the kernel cost this latency-bound kernel more than the divergence they remove.)
__global__ void __launch_bounds__(MT_A_BLOCK, MT_A_MINBLOCKS) k_step_a(StepDev p, NNTables T, MeshTables Mh) {
  const long long n = step_count(p);
  const int lane = threadIdx.x & 31;
  MT_TRACE_BEGIN(p.xdbg, 0)
  const long long i = (long long)blockIdx.x * MT_A_BLOCK + threadIdx.x;
  const bool valid = i < n;
  double et2 = 0.0, ang2 = 0.0;
  bool on_surface = valid;
  int todo = 0;
  float4 sq[2];
  if (valid) {
    float P[3][4], t[3], r[3], O[3][4], key[6];
    load_pose_stream(p.soa_cur, p.stride, i, P);
    const int hint = nn_index(mt_lds(p.nn_cur + i));
    nn_prefetch(T, hint);
    draw_or_load_noise(p.tn, p.rot, i, p.sig_t, p.sig_r, p.seed, p.step, p.first_gid + (uint64_t)i, t, r);
    apply_motion(P, p.odom, t, r, O, 0, p.tn == nullptr);
    // drift test: the voxel class is one dependent 4-byte load; it is requested here so that it travels while
    // the key is computed
    int mcls = 1;
    [[maybe_unused]] int mk = -1;
#if MT_MESH_DEFER && MT_VOX2
    if (p.prune_dist > 0.0) mcls = mesh_voxel_class2(Mh, O[0][3], O[1][3], O[2][3], p.prune_dist);
#else
    if (p.prune_dist > 0.0) mcls = mesh_voxel_class(Mh, O[0][3], O[1][3], O[2][3], p.prune_dist, &mk);
#endif
    store_pose_stream(p.soa_cur, p.stride, i, O);
    mt_se3_key(O, key);
    const bool invalid = mt_pose_invalid(O);
    if (invalid) atomicAdd(p.flags + 2, 1);  // check_quats would delete the particle (particle_filter.py:347-357)
    if (p.has_gt) rmse_terms(p.gt, O, et2, ang2);
#if MT_MESH_DEFER
    on_surface = mcls != 0;  // undecided: counted as on the surface until the queue consumer has looked
    if (mcls >= 2) todo |= MT_Q_MESH;
#else
    on_surface = (mcls >= 2) ? mesh_within_search(Mh, O[0][3], O[1][3], O[2][3], p.prune_dist, mcls == 2 ? mk : -1) : (mcls == 1);
#endif
    // hint-graph search
    float bd, dh;
    int bi, centre;
    int st = nn_hint_begin(T, key, hint, bd, bi, centre, dh);
    if (st == 0) st = nn_hint_scan_all(T, key, centre, dh, bd, bi);
    if (st <= 0) todo |= MT_Q_NN;  // no usable hint, or the list is exhausted: box-hierarchy search
    if (bi == INT_MAX) bi = -1;  // no usable hint
    const bool masked = !on_surface || invalid;
    // masked: weights *= m (particle_filter.py:398-401); a particle without any candidate yet
    // (-1) has its mask re-derived by k_step_nnq
    mt_sts(p.nn_cur + i, (bi >= 0 && masked) ? nn_masked(bi) : bi);
    if (todo & MT_Q_NN) {  // the search record travels with the queue entry: the consumer starts from it, not from the pose
      if (masked) todo |= MT_Q_MASKED;
      sq[0] = make_float4(key[0], key[1], key[2], key[3]);
      sq[1] = make_float4(key[4], key[5], bd, __int_as_float(bi));
    }
  }
  // queue what is left (one atomic per warp and queue).  Box-hierarchy searches go to the front of the queue array
  // (one warp per entry later on) together with their search record, drift tests that need nothing else to its back
  // (one thread per entry).
  const unsigned nm = __ballot_sync(0xffffffffu, (todo & MT_Q_NN) != 0);
  const unsigned mm = __ballot_sync(0xffffffffu, todo == MT_Q_MESH);
  if (nm | mm) {
    unsigned base = 0, mbase = 0;
    if (lane == 0) {
      if (nm) base = atomicAdd(p.qctl, (unsigned)__popc(nm));
      if (mm) mbase = atomicAdd(p.qctl + 3, (unsigned)__popc(mm));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    mbase = __shfl_sync(0xffffffffu, mbase, 0);
    const unsigned below = (1u << lane) - 1;
    if (todo & MT_Q_NN) {
      const unsigned e = base + __popc(nm & below);
      p.queue[e] = (int)i | (todo & (MT_Q_NN | MT_Q_MESH | MT_Q_MASKED));
      p.srec[2 * (size_t)e] = sq[0], p.srec[2 * (size_t)e + 1] = sq[1];
    } else if (todo) p.queue[p.queue_cap - 1 - (mbase + __popc(mm & below))] = (int)i;
  }
  const long long gw = i >> 5;  // global warp = 32 consecutive particles
  const unsigned on = __ballot_sync(0xffffffffu, on_surface);
  if (p.has_gt) et2 = warp_sum(et2), ang2 = warp_sum(ang2);
  if (lane == 0) {
    p.wcnt[gw] = __popc(on);
    if (p.has_gt) p.wrm[2 * gw] = et2, p.wrm[2 * gw + 1] = ang2;
  }
  MT_TRACE_END(p.xdbg, 0)
}

// queue consumer: one warp per entry.  MT_Q_NN: best-first search through the box hierarchy (nn_bvh_search)
// seeded with the candidate the hint scan left behind.  MT_Q_MESH: the vertex search of the drift test.
struct NnqEntry {
  long long i;
  float key[6];
  float bd;
  int bi;
  bool masked;
};
__device__ __forceinline__ void nnq_load(const StepDev& p, const MeshTables& Mh, unsigned e, NnqEntry& q) {
  // queue word and search record are independent loads: one memory round trip, no pose / match / key re-reads
  const int raw = p.queue[e];
  const float4 r0 = p.srec[2 * (size_t)e], r1 = p.srec[2 * (size_t)e + 1];
  const long long i = raw & MT_Q_INDEX;
  q.i = i;
  q.key[0] = r0.x, q.key[1] = r0.y, q.key[2] = r0.z, q.key[3] = r0.w, q.key[4] = r1.x, q.key[5] = r1.y;
  q.bd = r1.z;
  q.bi = __float_as_int(r1.w);
  if (q.bi < 0) q.bi = INT_MAX, q.bd = FLT_MAX;  // no candidate yet
  q.masked = (raw & MT_Q_MASKED) != 0;
  if ((raw & MT_Q_MESH) && !q.masked) {  // provisionally on the surface: look now
    const float x = p.soa_cur[i].w, y = p.soa_cur[p.stride + i].w, z = p.soa_cur[2 * p.stride + i].w;
    if (!mesh_within_warp(Mh, x, y, z, p.prune_dist)) {
      q.masked = true;
      if ((threadIdx.x & 31) == 0) atomicSub(p.wcnt + (i >> 5), 1);
    }
  }
}
#define MT_NNQ_WARPS 4
// Deferred drift tests (voxel class "undecided" in k_step_a): the vertex search of remove_invalid_particles.  The
// particle was counted as on the surface; a failed test takes that back.
//   k_step_meshq   one thread per entry: the two quick certificates (mesh_quick) settle ~3/4 of the entries with one
//                  vertex load; the rest is compacted into a second list
//   k_step_meshq2  one warp per entry of that list: grid search, the lanes over the candidate vertices
// Both kernels are chains of dependent memory round trips (count -> entry -> point -> tables -> match), ~1.2 us each,
// with next to no arithmetic: the first trip's entry is therefore requested before the count is known (a stale or
// uninitialised entry is clamped to a valid particle and dropped once the count has arrived), and the particle's point
// and stored match are requested together.  (Doing the grid search of the left-over quarter right inside k_step_meshq, one
// thread per entry, instead of compacting it for the warp-per-entry kernel: 55 us against 12.5 + 2.3 + 12 us.)
__device__ __forceinline__ void mesh_mark_off(const StepDev& p, long long i, int stored) {
  if (stored >= 0) p.nn_cur[i] = nn_masked(stored);  // (nobody else touches a queued particle's match meanwhile)
  atomicSub(p.wcnt + (i >> 5), 1);
}
__global__ void __launch_bounds__(256) k_step_meshq(StepDev p, MeshTables Mh) {
  const int lane = threadIdx.x & 31;
  const unsigned span = gridDim.x * blockDim.x;
  const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
  int raw = ((long long)first < p.queue_cap) ? p.queue[p.queue_cap - 1 - first] : 0;  // speculative
  const unsigned mn = p.qctl[3];
  MT_TRACE_BEGIN(p.xdbg, 1)
  for (unsigned e0 = first - lane; e0 < mn; e0 += span) {  // whole warps stay in the loop
    const unsigned e = e0 + lane;
    if (e0 != first - lane && e < mn) raw = p.queue[p.queue_cap - 1 - e];
    const long long n_cov = p.stride;
    const long long i = raw < 0 ? 0 : ((long long)raw >= n_cov ? n_cov - 1 : raw);
    const float x = p.soa_cur[i].w, y = p.soa_cur[p.stride + i].w, z = p.soa_cur[2 * p.stride + i].w;
    const int stored = p.nn_cur[i];
    int res = 1;  // 1 within, 0 not within, 2 grid search
    if (e < mn) {
      int k = -1;
      const int c = mesh_voxel_class(Mh, x, y, z, p.prune_dist, &k);
      res = c < 2 ? c : (c == 2 ? mesh_quick(Mh, x, y, z, p.prune_dist, k) : 2);
      if (res == 0) mesh_mark_off(p, i, stored);
    }
    const unsigned m2 = __ballot_sync(0xffffffffu, res == 2);
    if (m2) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(p.qctl + 2, (unsigned)__popc(m2));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (res == 2) p.queue2[base + __popc(m2 & ((1u << lane) - 1))] = (int)i;
    }
  }
  MT_TRACE_END(p.xdbg, 1)
}
__global__ void __launch_bounds__(256) k_step_meshq2(StepDev p, MeshTables Mh) {
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  int raw = ((long long)gw < p.queue_cap) ? p.queue2[gw] : 0;  // speculative
  const unsigned n2 = p.qctl[2];
  MT_TRACE_BEGIN(p.xdbg, 2)
  for (unsigned e = gw; e < n2; e += nw) {
    if (e != gw) raw = p.queue2[e];
    const long long i = raw < 0 ? 0 : ((long long)raw >= p.stride ? p.stride - 1 : raw);
    const float x = p.soa_cur[i].w, y = p.soa_cur[p.stride + i].w, z = p.soa_cur[2 * p.stride + i].w;
    const int stored = p.nn_cur[i];
    const bool on = mesh_search_warp(Mh, x, y, z, p.prune_dist);
    if (!on && (threadIdx.x & 31) == 0) mesh_mark_off(p, i, stored);
  }
  MT_TRACE_END(p.xdbg, 2)
}

__global__ void __launch_bounds__(32 * MT_NNQ_WARPS) k_step_nnq(StepDev p, NNTables T, MeshTables Mh) {
  const unsigned qn = p.qctl[0];
  const int lane = threadIdx.x & 31;
  MT_TRACE_BEGIN(p.xdbg, 3)
  if (qn) {
    // the search index (boxes 0.1 MB + Morton-ordered keys 32 B each) was last touched a step ago: pull it into
    // L2 with one prefetch per 128-byte line, spread over the grid, so that the dependent rounds of the searches
    // below pay L2 latency instead of DRAM latency
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    const size_t box_lines = ((size_t)(T.b.n_leaf + T.b.n_l1 + T.b.n_l2) * 48 + 127) / 128;
    const size_t key_lines = ((size_t)T.M * 32 + 127) / 128;
    for (size_t l = gtid; l < box_lines + key_lines; l += gsz) {
      const char* a = (l < box_lines) ? (const char*)T.bvh_leaf + 128 * l : (const char*)T.keys_sorted + 128 * (l - box_lines);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
  // Few entries (the usual few hundred): one BLOCK per entry, its four warps search as a team (nn_bvh_search_team); a
  // drift test that is still pending for the particle is done by the fourth warp meanwhile.  Many entries (rod-like
  // objects: hundreds of thousands): one warp per entry, pulled from a shared counter -- throughput, not the tail, matters.
  __shared__ unsigned long long s_best;
  __shared__ int s_leaves, s_masked;
  const unsigned W = gridDim.x * MT_NNQ_WARPS;
  if (qn <= gridDim.x) {
    const int warp = threadIdx.x >> 5;
    for (unsigned e = blockIdx.x; e < qn; e += gridDim.x) {  // (at most one trip; block-uniform)
#if MT_TRACE
      const unsigned long long tr0 = mt_now();
#endif
      const int raw = p.queue[e];
      const float4 r0 = p.srec[2 * (size_t)e], r1 = p.srec[2 * (size_t)e + 1];
      const long long i = raw & MT_Q_INDEX;
      const float key[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
      const bool masked0 = (raw & MT_Q_MASKED) != 0;
      const bool mesh = (raw & MT_Q_MESH) && !masked0;
      const bool nan = !(key[0] == key[0]) || !(key[1] == key[1]) || !(key[2] == key[2]) || !(key[3] == key[3]) || !(key[4] == key[4]) || !(key[5] == key[5]);
      if (threadIdx.x == 0) {
        const int bi = __float_as_int(r1.w);
        s_best = bi < 0 ? mt_dist_word(FLT_MAX, INT_MAX) : mt_dist_word(r1.z, bi);
        s_leaves = 0;
        s_masked = masked0 ? 1 : 0;
      }
      __syncthreads();
      if (mesh && warp == MT_NNQ_WARPS - 1) {
        const float x = p.soa_cur[i].w, y = p.soa_cur[p.stride + i].w, z = p.soa_cur[2 * p.stride + i].w;
        if (!mesh_within_warp(Mh, x, y, z, p.prune_dist) && lane == 0) {
          s_masked = 1;
          atomicSub(p.wcnt + (i >> 5), 1);
        }
      } else if (!nan) {
        nn_bvh_search_team(T, key, warp, mesh ? MT_NNQ_WARPS - 1 : MT_NNQ_WARPS, &s_best, &s_leaves);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const int bi = (int)(unsigned)s_best;
        const int res = (nan || bi == INT_MAX) ? 0 : bi;  // NaN / Inf query: np.argmin semantics
        p.nn_cur[i] = s_masked ? nn_masked(res) : res;
        atomicAdd(p.flags + 4, s_leaves);
        atomicMax(p.flags + 4 + 3, s_leaves);
#if MT_TRACE
        const unsigned long long tr2 = mt_now();
        atomicMax(p.xdbg + 30, tr2 - tr0), atomicAdd(p.xdbg + 31, tr2 - tr0), atomicAdd(p.xdbg + 32, 1ull), atomicMax(p.xdbg + 35, tr0);
#endif
      }
      __syncthreads();
    }
    MT_TRACE_END(p.xdbg, 3)
    return;
  }
  // the first entry of every warp is assigned statically (one per block first, so that they spread over the SMs):
  // with the usual few hundred entries nobody touches the shared counter -- thousands of warps opening with an
  // atomic on one address serialised in L2 and were most of this kernel's duration.  Further entries are pulled.
  // (Handing long queues out in runs of 4 / 8 / 16 consecutive entries per warp -- neighbours in the queue open the same
  // leaves, which a warp doing them back to back would find in L1 -- was measured slower on the cotter-pin stand-in:
  // 717 / 772 / 860 us against 681 us for 302k searches; the tail of uneven runs costs more than the L1 hits save.)
  unsigned e = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  while (e < qn) {
    NnqEntry q;
#if MT_TRACE
    const unsigned long long tr0 = mt_now();
#endif
    nnq_load(p, Mh, e, q);
#if MT_TRACE
    const unsigned long long tr1 = mt_now();
#endif
    const int res = nn_bvh_search(T, q.key, q.bd, q.bi, p.flags + 4);
    if (lane == 0) p.nn_cur[q.i] = q.masked ? nn_masked(res) : res;
#if MT_TRACE
    if (lane == 0) {  // words 30..35: slowest entry, sum over entries, entries, slowest load part, sum of load parts, latest entry start
      const unsigned long long tr2 = mt_now();
      atomicMax(p.xdbg + 30, tr2 - tr0), atomicAdd(p.xdbg + 31, tr2 - tr0), atomicAdd(p.xdbg + 32, 1ull);
      atomicMax(p.xdbg + 33, tr1 - tr0), atomicAdd(p.xdbg + 34, tr1 - tr0), atomicMax(p.xdbg + 35, tr0);
    }
#endif
    if (qn <= W) break;
    if (lane == 0) e = W + atomicAdd(p.qctl + 1, 1u);
    e = __shfl_sync(0xffffffffu, e, 0);
  }
  MT_TRACE_END(p.xdbg, 3)
}

// chunk sums of the weights + (last block) prefix for kernel B, RMSE, drift flag
__global__ void __launch_bounds__(256) k_step_sums(StepDev p) {
  __shared__ double s8[8];
  __shared__ int s_cnt;
  const long long i = (long long)blockIdx.x * MT_CHUNK + threadIdx.x;
  const long long n = step_count(p);
  const int nwarps = (int)((n + 31) >> 5);
  double e = 0.0;
  if (i < n) {
    const int stored = mt_lds(p.nn_cur + i);
    e = nn_is_masked(stored) ? 0.0 : mt_ldk(p.wtab + nn_index(stored));
  }
  const double se = block_sum_256(e, s8);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    p.part[blockIdx.x] = se;
    // fold kernel A's 8 per-warp partials of this chunk (fixed order)
    double ra = 0.0, rb = 0.0;
    int cnt = 0;
    for (int j = 0; j < 8; ++j) {
      const int gw = 8 * blockIdx.x + j;
      if (gw < nwarps) {
        cnt += p.wcnt[gw];
        if (p.has_gt) ra += p.wrm[2 * gw], rb += p.wrm[2 * gw + 1];
      }
    }
    p.rm_part[2 * blockIdx.x] = ra, p.rm_part[2 * blockIdx.x + 1] = rb;
    p.wcnt[8 * blockIdx.x] = cnt;  // slot 8c now holds the chunk count
    __threadfence();
    last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  scan_chunk_sums(p.part, (int)gridDim.x, p.prefix, p.scal, s8);
  double sa = 0.0, sb = 0.0;
  if (p.has_gt) {
    sa = block_sum_array_256(p.rm_part, (int)gridDim.x, 2, s8);
    sb = block_sum_array_256(p.rm_part + 1, (int)gridDim.x, 2, s8);
  }
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int cnt = 0;
  for (int c = threadIdx.x; c < (int)gridDim.x; c += MT_CHUNK) cnt += __ldcg(p.wcnt + 8 * (size_t)c);
  if (cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.has_gt) {
      p.rmse2[0] = (float)sqrt(sa / (double)n);
      p.rmse2[1] = (float)sqrt(sb / (double)n);
    }
    p.flags[6] = s_cnt;                               // particles on the surface this step
    p.flags[5] = (p.prune_dist > 0.0 && s_cnt == 0);  // drifted (particle_filter.py:402)
    p.flags[3] += (int)p.qctl[0];                     // searches that needed the box hierarchy (cumulative)
    p.flags[MT_STAT_MESH_DEFERRED] += (int)p.qctl[3];
    p.qctl[0] = 0, p.qctl[1] = 0, p.qctl[2] = 0, p.qctl[3] = 0;
    *p.ticket = 0;
  }
}

// chunk sums + prefix for explicit weights (resampler on caller-provided weights)
__global__ void __launch_bounds__(256) k_weight_sums(StepDev p) {
  __shared__ double s8[8];
  const long long i = (long long)blockIdx.x * MT_CHUNK + threadIdx.x;
  double e = (i < p.n) ? p.wsrc[i] : 0.0;
  double se = block_sum_256(e, s8);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    p.part[blockIdx.x] = se;
    __threadfence();
    last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    scan_chunk_sums(p.part, (int)gridDim.x, p.prefix, p.scal, s8);
    if (threadIdx.x == 0) *p.ticket = 0;
  }
}

// Kernel B.  One chunk per block.  CDF value of particle i:
//     C_i = (A_rank + prefix[chunk] + incl_i) / S     (float64, fixed topology)
// with chunk ends pinned to (A_rank + prefix[chunk+1]) / S so neighbouring blocks and
// neighbouring GPUs agree bit-for-bit on the slot boundary.  Slots owned by particle i are
// [cnt(C_{i-1}), cnt(C_i)), cnt(C) = #{j : loc_j < C}; each parent writes its own children
// (reads and writes are both contiguous up to the child-count jitter).
//   HBM per particle: read 4 B nn + 48 B pose, write 48 B pose + 4 B nn + 4 B ancestor.
// One 256-particle chunk of kernel B: CDF values from (base, inclusive scan, endv), slot ownership,
// scatter of the children.  base / endv are the chunk's CDF interval in un-normalised units
// (A + prefix[c], A + prefix[c+1]); S the global normaliser; A this shard's offset.
// the per-particle inputs of one chunk of kernel B (loaded ahead of their use)
struct ChunkIn {
  int nn;
  double e;
  float P[3][4];
};
template <bool FROM_TABLE, bool SCATTER>
__device__ __forceinline__ void step_b_load(const StepDev& p, const int c, const long long n, ChunkIn& in) {
  const long long i = (long long)c * MT_CHUNK + threadIdx.x;
  in.nn = 0;
  in.e = 0.0;
  if (i < n) {
    if (FROM_TABLE) {
      const int stored = mt_lds(p.nn_cur + i);
      in.nn = nn_index(stored);
      in.e = nn_is_masked(stored) ? 0.0 : mt_ldk(p.wtab + in.nn);
    } else {
      in.e = p.wsrc[i];
    }
    if (SCATTER) load_pose_stream(p.soa_cur, p.stride, i, in.P);
  }
}

template <bool FROM_TABLE, bool SCATTER>
__device__ __forceinline__ void step_b_chunk(const StepDev& p, const int c, const long long n, const double S, const double A,
                                             const double base, const double endv, double* s8, long long* s_cnt, ChunkIn& in,
                                             const long long slot_base_in = -1) {
  const long long i = (long long)c * MT_CHUNK + threadIdx.x;
  const bool valid = i < n;
  const int nn = in.nn;
  const double e = in.e;
  float (&P)[3][4] = in.P;
  const long long N = p.n_global;
  const double dN = (double)N;
  const double off = (double)(p.u / (float)N);  // float32 division, then promoted (particle_filter.py:260)
  const bool bad = !(S > 0.0) || !(S <= DBL_MAX);  // all-zero / NaN / Inf weights: identity (237-241)
  const double incl = block_incl_scan_256(e, s8);
  long long cnt;
  if (bad) {
    // the reference returns the particles unchanged (237-241); when every particle has drifted off
    // the mesh it first re-projects them onto the codebook (filter.py:176-179)
    if (SCATTER && FROM_TABLE && valid && S == 0.0 && p.prune_dist > 0.0 && p.cb_poses) {
      const float4 a = __ldg(p.cb_poses + 4 * (size_t)nn), b = __ldg(p.cb_poses + 4 * (size_t)nn + 1),
                   c2 = __ldg(p.cb_poses + 4 * (size_t)nn + 2);
      P[0][0] = a.x, P[0][1] = a.y, P[0][2] = a.z, P[0][3] = a.w;
      P[1][0] = b.x, P[1][1] = b.y, P[1][2] = b.z, P[1][3] = b.w;
      P[2][0] = c2.x, P[2][1] = c2.y, P[2][2] = c2.z, P[2][3] = c2.w;
    }
    cnt = valid ? i + 1 : n;
    if (threadIdx.x == 0) s_cnt[0] = (long long)c * MT_CHUNK, p.flags[1] = 1;
  } else {
    const bool is_end = (threadIdx.x == MT_CHUNK - 1) || (i == n - 1);
    double C = is_end ? endv / S : (base + incl) / S;
    cnt = valid ? mt_count_below(C, N, dN, off) : 0;
    if (threadIdx.x == 0) s_cnt[0] = mt_count_below(base / S, N, dN, off);
  }
  s_cnt[threadIdx.x + 1] = cnt;
  __syncthreads();
  // first slot of this shard (sharded runs); callers that loop over chunks pass it in
  const long long slot_base = (bad || p.world == 1) ? 0 : (slot_base_in >= 0 ? slot_base_in : mt_count_below(A / S, N, dN, off));
  if (valid && i == n - 1 && p.n_out) *p.n_out = cnt - slot_base;
  long long prev = s_cnt[threadIdx.x];
  long long kids = valid ? (cnt - prev) : 0;
  if (kids < 0) kids = 0;
  long long dst = prev - slot_base;
  const long long cap = p.stride;
  // light parents write their own children; heavy parents are spread over the warp
  const int lane = threadIdx.x & 31;
  const bool heavy = kids > 8;
  if (!heavy) {
    for (long long k = 0; k < kids; ++k) {
      long long s = dst + k;
      if (s >= cap) {
        p.flags[0] = 1;
        break;
      }
      if (SCATTER) {
        store_pose_stream(p.soa_next, p.stride, s, P);
        if (FROM_TABLE) mt_sts(p.nn_next + s, nn);
      }
      if (p.anc) mt_sts(p.anc + s, (int)i);
    }
  }
  unsigned hm = __ballot_sync(0xffffffffu, heavy);
  while (hm) {
    const int src = __ffs(hm) - 1;
    hm &= hm - 1;
    float Q[3][4];
    if (SCATTER) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Q[a][b] = __shfl_sync(0xffffffffu, P[a][b], src);
    }
    const long long d0 = __shfl_sync(0xffffffffu, dst, src);
    const long long kn = __shfl_sync(0xffffffffu, kids, src);
    const int pn = __shfl_sync(0xffffffffu, nn, src);
    const long long pi = __shfl_sync(0xffffffffu, i, src);
    for (long long k = lane; k < kn; k += 32) {
      long long s = d0 + k;
      if (s >= cap) {
        p.flags[0] = 1;
        break;
      }
      if (SCATTER) {
        store_pose_stream(p.soa_next, p.stride, s, Q);
        if (FROM_TABLE) mt_sts(p.nn_next + s, pn);
      }
      if (p.anc) mt_sts(p.anc + s, (int)pi);
    }
  }
}

template <bool FROM_TABLE, bool SCATTER>
__global__ void __launch_bounds__(256) k_step_b(StepDev p) {
  __shared__ double s8[8];
  __shared__ long long s_cnt[MT_CHUNK + 1];
  const int c = blockIdx.x;
  const long long n = step_count(p);
  // global normaliser and this shard's CDF offset (sequential, identical on every GPU)
  double S = 0.0, A = 0.0;
  if (p.world > 1) {
    for (int r = 0; r < p.world; ++r) {
      if (r == p.rank) A = S;
      S += p.shard_sums[r];
    }
  } else {
    S = p.prefix[p.nchunks];
  }
  const double base = A + p.prefix[c];
  const double endv = (c + 1 == p.nchunks) ? (A + p.prefix[p.nchunks]) : (A + p.prefix[c + 1]);
  if (n == 0 && c == 0 && threadIdx.x == 0 && p.n_out) *p.n_out = 0;
  ChunkIn in;
  step_b_load<FROM_TABLE, SCATTER>(p, c, n, in);
  step_b_chunk<FROM_TABLE, SCATTER>(p, c, n, S, A, base, endv, s8, s_cnt, in);
}

// Single-GPU fused form of k_step_sums + k_step_b: a persistent cooperative grid.  Phase 1: every
// block sums the weights of its contiguous run of chunks; grid barrier; phase 2: every block adds
// the block totals before it in the same sequential order (so neighbouring blocks agree bit for bit
// on their common boundary), then resamples its chunks.  Saves a launch, the 4 B/particle re-read
// and the serial last-block scan.
#define MT_BW_MAX_PER 32
#define MT_BW_FAST_PER 10  // chunks per block whose weights stay in shared memory (30 KB)
#define MT_BW_MAX_GRID 1184
#ifndef MT_BW_PREFETCH
#define MT_BW_PREFETCH 0
#endif
#ifndef MT_BW_FLAT
#define MT_BW_FLAT 1  // barrier-free phase 2 when the block's weights fit in shared memory
#endif
#ifndef MT_BW_L2PREFETCH
#define MT_BW_L2PREFETCH 1
#endif
#ifndef MT_BW_P1
#define MT_BW_P1 4  // chunks whose look-ups phase 1 issues together
#endif
__global__ void __launch_bounds__(256, 4) k_step_bw(StepDev p, unsigned long long* bar, unsigned long long bar_target,
                                                 double* __restrict__ blocktot /* 3 x grid */, int* __restrict__ blockcnt) {
  __shared__ double s8[8];
  __shared__ long long s_cnt[MT_CHUNK + 1];
  __shared__ double s_part[MT_BW_MAX_PER];
  __shared__ double s_tot[MT_BW_MAX_GRID];
  __shared__ double s_bc[3];
  __shared__ double s_e[MT_BW_FAST_PER * MT_CHUNK];  // weight of every particle of the block's chunks
  __shared__ int s_nn[MT_BW_FAST_PER * MT_CHUNK];    // and its match (phase 2 does not go back to memory for them)
  __shared__ double s_x2[MT_CHUNK + 1];              // flat form: exclusive prefix of the threads' runs of weights
  const int G = gridDim.x, g = blockIdx.x;
  const long long n = step_count(p);
  const int nwarps = (int)((n + 31) >> 5);
  // chunks per block from the ACTUAL particle count (sharded runs size buffers and grids from the capacity, 1.5 x the
  // nominal shard: dividing the capacity's chunks would leave a third of the blocks idle and the others with 10 chunks)
  const int nch = (int)min((long long)p.nchunks, (n + MT_CHUNK - 1) / MT_CHUNK);
  const int per = max(1, (nch + G - 1) / G);
  const int c_lo = min(g * per, nch), c_hi = min(c_lo + per, nch);
  MT_TRACE_BEGIN(p.xdbg, 4)
  // ---- phase 1: weights of this block's chunks.  Up to MT_BW_FAST_PER chunks keep (match, weight) of
  // every particle in shared memory for phase 2; the look-ups are issued four chunks at a time so that
  // their DRAM / L2 latencies overlap.
  const bool cached = per <= MT_BW_FAST_PER;
  double tot = 0.0, ra = 0.0, rb = 0.0;
  int cnt = 0;
#if MT_BW_FLAT
  if (cached) {
    // flat form: only the block total is needed.  All look-ups of the thread are issued together, the thread adds its
    // weights in chunk order and ONE fixed-topology block sum follows; kernel A's per-warp partials of the block's
    // chunks are folded by the first 8 * per threads the same way.  While these latencies run, the poses the scatter
    // of phase 2 will read are requested into L2 (one prefetch per 32-byte sector).
    int stv[MT_BW_FAST_PER];
#pragma unroll
    for (int k = 0; k < MT_BW_FAST_PER; ++k) {
      const long long i = (long long)(c_lo + k) * MT_CHUNK + threadIdx.x;
      stv[k] = (k < per && c_lo + k < c_hi && i < n) ? mt_lds(p.nn_cur + i) : -1;
    }
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < MT_BW_FAST_PER; ++k) {
      if (k < per && c_lo + k < c_hi) {
        const double e = (stv[k] >= 0) ? mt_ldk(p.wtab + stv[k]) : 0.0;  // masked (< -1) and absent (-1): 0
        s_e[k * MT_CHUNK + threadIdx.x] = e, s_nn[k * MT_CHUNK + threadIdx.x] = nn_index(stv[k]);
        mine += e;
      }
    }
#if MT_BW_L2PREFETCH
    if (!(threadIdx.x & 1)) {
#pragma unroll
      for (int k = 0; k < MT_BW_FAST_PER; ++k) {
        const long long i = (long long)(c_lo + k) * MT_CHUNK + threadIdx.x;
        if (k < per && c_lo + k < c_hi && i < n) {
#pragma unroll
          for (int r = 0; r < 3; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.soa_cur + (size_t)r * p.stride + i));
        }
      }
    }
#endif
    int mycnt = 0;
    double mra = 0.0, mrb = 0.0;
    {
      const int gw = 8 * c_lo + (int)threadIdx.x;
      if ((int)threadIdx.x < 8 * (c_hi - c_lo) && gw < nwarps) {
        mycnt = p.wcnt[gw];
        if (p.has_gt) mra = p.wrm[2 * gw], mrb = p.wrm[2 * gw + 1];
      }
    }
    tot = block_sum_256(mine, s8);
    if (p.has_gt) {
      ra = block_sum_256(mra, s8);
      rb = block_sum_256(mrb, s8);
    }
    cnt = (int)block_sum_256((double)mycnt, s8);  // exact: at most 2560 particles
    // The in-block part of phase 2's scan needs neither the global sum nor the block's offset, so it runs here, while
    // the slower blocks are still on their way to the grid barrier: thread t owns `per` consecutive particles of the
    // block (transposed view of s_e), replaces their weights by the thread-local inclusive prefix, and the thread
    // totals are block-scanned (monotone form) into s_x2.
    {
      const long long left = n - (long long)c_lo * MT_CHUNK;
      const int nloc = (int)(left < 0 ? 0 : (left < (long long)(c_hi - c_lo) * MT_CHUNK ? left : (long long)(c_hi - c_lo) * MT_CHUNK));
      const int j0 = threadIdx.x * per;
      double acc = 0.0;
      __syncthreads();  // s_e is complete
#pragma unroll
      for (int k = 0; k < MT_BW_FAST_PER; ++k) {
        if (k < per && j0 + k < nloc) {
          acc += s_e[j0 + k];
          s_e[j0 + k] = acc;
        }
      }
      const double I = block_incl_scan_mono_256(acc, s8);
      s_x2[threadIdx.x + 1] = I;
      if (threadIdx.x == 0) s_x2[0] = 0.0;
    }
  } else
#endif
  for (int c0 = c_lo; c0 < c_hi; c0 += MT_BW_P1) {
    int st4[MT_BW_P1];
    double e4[MT_BW_P1];
#pragma unroll
    for (int k = 0; k < MT_BW_P1; ++k) {
      const long long i = (long long)(c0 + k) * MT_CHUNK + threadIdx.x;
      st4[k] = (c0 + k < c_hi && i < n) ? mt_lds(p.nn_cur + i) : -1;
    }
#pragma unroll
    for (int k = 0; k < MT_BW_P1; ++k) e4[k] = (st4[k] >= 0) ? mt_ldk(p.wtab + st4[k]) : 0.0;  // masked (< -1) and absent (-1): 0
#pragma unroll
    for (int k = 0; k < MT_BW_P1; ++k) {
      const int c = c0 + k;
      if (c >= c_hi) break;
      if (cached) s_e[(c - c_lo) * MT_CHUNK + threadIdx.x] = e4[k], s_nn[(c - c_lo) * MT_CHUNK + threadIdx.x] = nn_index(st4[k]);
      const double se = block_sum_256(e4[k], s8);
      if (threadIdx.x == 0) {
        s_part[c - c_lo] = se;
        tot += se;
        for (int j = 0; j < 8; ++j) {  // kernel A's per-warp partials of this chunk (fixed order)
          const int gw = 8 * c + j;
          if (gw < nwarps) {
            cnt += p.wcnt[gw];
            if (p.has_gt) ra += p.wrm[2 * gw], rb += p.wrm[2 * gw + 1];
          }
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    blocktot[g] = tot, blocktot[G + g] = ra, blocktot[2 * G + g] = rb;
    blockcnt[g] = cnt;
    __threadfence();
    MT_TRACE_MAX(p.xdbg, 24)
    atomicAdd(bar, 1ull);
    while (*(volatile unsigned long long*)bar < bar_target) {
    }
    __threadfence();
    MT_TRACE_MIN(p.xdbg, 25)
    MT_TRACE_MAX(p.xdbg, 26)
  }
  __syncthreads();
  // ---- phase 2
  for (int j = threadIdx.x; j < G; j += blockDim.x) s_tot[j] = __ldcg(blocktot + j);
  __syncthreads();
  {
    // prefix of the G block totals, the same code (hence bit-identical values) in every block, by ONE warp (no block
    // barriers): lane l owns a contiguous run of totals; the run sums are scanned across the warp (Kogge-Stone plus a
    // running maximum, which removes the ulp-sized inversions float64 rounding can leave) into E[0..32]; the values
    // inside run l are E[l] + (local prefix), clamped to E[l+1]; a run's last end IS E[l+1], which is also where the
    // next run starts -- so neighbouring blocks agree on their common boundary and everything is monotone
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int perb = (G + 31) / 32;  // <= 37
      const int b0 = lane * perb;
      double loc = 0.0;
      for (int k = 0; k < perb; ++k)
        if (b0 + k < G) loc += s_tot[b0 + k];
      double I = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, I, o);
        if (lane >= o) I += t;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, I, o);
        if (lane >= o) I = fmax(I, t);
      }
      double E = __shfl_up_sync(0xffffffffu, I, 1);
      if (lane == 0) E = 0.0;
      const double En = I;
      if (lane == 31) s_bc[2] = I;
      if (g >= b0 && g < b0 + perb) {
        double lp = 0.0;
        for (int j = b0; j < g; ++j) lp += s_tot[j];
        const double start = (g == b0) ? E : fmin(E + lp, En);
        lp += s_tot[g];
        const double end = (g == b0 + perb - 1 || g == G - 1) ? En : fmin(E + lp, En);
        s_bc[0] = start, s_bc[1] = end;
      }
    }
  }
  __syncthreads();
  double base_g = s_bc[0], next_g = s_bc[1], S = s_bc[2], A = 0.0;
  if (p.world > 1) {
    // sharded: this GPU's total goes straight into every peer's exchange buffer (remote stores over
    // NVLink), then every block waits for all the totals of this step to have arrived locally
    __shared__ double s_x[2];
    const int par = (int)(p.xseq & 1);
    const unsigned seq = (unsigned)p.xseq;
    if (g == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.xdbg[0]));
    {  // while the sums travel: pull the poses of this block's first chunk towards L2
      const long long i0 = (long long)c_lo * MT_CHUNK + threadIdx.x;
      if (c_lo < c_hi && i0 < n) {
#pragma unroll
        for (int r = 0; r < 3; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.soa_cur + (size_t)r * p.stride + i0));
      }
    }
    if (g == 0 && threadIdx.x < p.world) {
      Xchg* peer = p.peers[threadIdx.x];
      const unsigned long long bits = (unsigned long long)__double_as_longlong(S);
      volatile unsigned long long* dst = peer->w[par][p.rank];
      dst[0] = (bits << 32) | seq;
      dst[1] = (bits & 0xffffffff00000000ull) | seq;
      if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.xdbg[1]));
    }
    if (threadIdx.x == 0) {
      Xchg* own = p.peers[p.rank];
      double tot = 0.0, off = 0.0;
      bool ok = true;
      for (int r = 0; r < p.world; ++r) {
        volatile unsigned long long* src = own->w[par][r];
        unsigned long long t0 = 0, w0, w1;
        int spins = 0;
        for (;;) {
          w0 = src[0], w1 = src[1];
          if ((unsigned)w0 == seq && (unsigned)w1 == seq) break;
          if ((++spins & 1023) == 0) {  // a peer that never arrives (10 s) is flagged instead of hanging the GPU
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            if (t - t0 > 10000000000ull) {
              ok = false;
              break;
            }
          }
        }
        if (r == p.rank) off = tot;
        tot += __longlong_as_double((long long)((w0 >> 32) | (w1 & 0xffffffff00000000ull)));  // sequential, identical on every GPU
      }
      if (!ok) {  // a peer never arrived: S is not trustworthy -> NaN, i.e. the "keep the particles" path of every chunk
        p.flags[0] = 2;
        tot = __longlong_as_double(0x7ff8000000000000LL);
      }
      if (g == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.xdbg[2]));
      s_x[0] = off, s_x[1] = tot;
    }
    __syncthreads();
    A = s_x[0];
    S = s_x[1];
    base_g += A, next_g += A;
  }
  long long slot_base = 0;
  if (p.world > 1 && S > 0.0 && S <= DBL_MAX) {
    const long long N = p.n_global;
    slot_base = mt_count_below(A / S, N, (double)N, (double)(p.u / (float)N));
  }
  MT_TRACE_MAX(p.xdbg, 27)
#if MT_BW_FLAT
  if (cached) {
    // ---- phase 2, flat form: one scan over ALL of the block's weights (they sit in shared memory), then a scatter
    // loop without a single block barrier.  The chunk-by-chunk form below synchronises the block five times per
    // chunk, and a third of its stall cycles were barrier waits.
    //  (a) transposed pass: thread t owns `per` consecutive particles of the block; sequential local prefix, monotone
    //      block scan of the thread totals, CDF value = fmin((base_g + X_t) + local, base_g + X_{t+1}, next_g) -- a
    //      non-decreasing sequence that starts at base_g and is pinned to next_g at the block's last particle, so
    //      neighbouring blocks (and GPUs) still agree bit for bit on their boundary; slot counts replace the weights
    //      in shared memory
    //  (b) coalesced pass: particle (chunk, thread) reads its own and its predecessor's count and writes its children
    const long long i0 = (long long)c_lo * MT_CHUNK;
    const long long left = n - i0;
    const int nloc = (int)(left < 0 ? 0 : (left < (long long)(c_hi - c_lo) * MT_CHUNK ? left : (long long)(c_hi - c_lo) * MT_CHUNK));
    const long long N = p.n_global;
    const double dN = (double)N;
    const double off = (double)(p.u / (float)N);  // float32 division, then promoted (particle_filter.py:260)
    const bool bad = !(S > 0.0) || !(S <= DBL_MAX);  // all-zero / NaN / Inf weights: identity (237-241)
    long long* s_cntall = reinterpret_cast<long long*>(s_e);
    __shared__ long long s_cbase;
    if (!bad) {
      // C = x * (1 / S): one division per thread instead of one per particle; every block forms the same products of
      // the same values, so the pinned interval ends still agree bit for bit
      const double rS = 1.0 / S;
      const int j0 = threadIdx.x * per;
      if (threadIdx.x == 0) s_cbase = mt_count_below(base_g * rS, N, dN, off);
      const double lo = base_g + s_x2[threadIdx.x], hi = fmin(base_g + s_x2[threadIdx.x + 1], next_g);
#pragma unroll
      for (int k = 0; k < MT_BW_FAST_PER; ++k) {
        if (k < per && j0 + k < nloc) {
          const bool is_end = (j0 + k == nloc - 1);
          const double C = (is_end ? next_g : fmin(lo + s_e[j0 + k], hi)) * rS;
          s_cntall[j0 + k] = mt_count_below(C, N, dN, off);
        }
      }
    } else if (threadIdx.x == 0) {
      p.flags[1] = 1;
    }
    __syncthreads();
    const long long cap = p.stride;
    const int lane = threadIdx.x & 31;
    for (int c = c_lo; c < c_hi; ++c) {
      const int il = (c - c_lo) * MT_CHUNK + threadIdx.x;
      const long long i = i0 + il;
      const bool valid = il < nloc;
      float P[3][4];
      int nn = 0;
      long long cnt = 0, prev = 0;
      if (valid) {
        load_pose_stream(p.soa_cur, p.stride, i, P);
        nn = s_nn[il];
        if (bad) {
          // the reference returns the particles unchanged (237-241); when every particle has drifted off the mesh it
          // first re-projects them onto the codebook (filter.py:176-179)
          if (S == 0.0 && p.prune_dist > 0.0 && p.cb_poses) {
            const float4 a = __ldg(p.cb_poses + 4 * (size_t)nn), b = __ldg(p.cb_poses + 4 * (size_t)nn + 1),
                         c2 = __ldg(p.cb_poses + 4 * (size_t)nn + 2);
            P[0][0] = a.x, P[0][1] = a.y, P[0][2] = a.z, P[0][3] = a.w;
            P[1][0] = b.x, P[1][1] = b.y, P[1][2] = b.z, P[1][3] = b.w;
            P[2][0] = c2.x, P[2][1] = c2.y, P[2][2] = c2.z, P[2][3] = c2.w;
          }
          cnt = i + 1, prev = i;
        } else {
          cnt = s_cntall[il];
          prev = il > 0 ? s_cntall[il - 1] : s_cbase;
        }
        if (i == n - 1 && p.n_out) *p.n_out = cnt - (bad ? 0 : slot_base);
      }
      long long kids = valid ? (cnt - prev) : 0;
      if (kids < 0) kids = 0;
      const long long dst = prev - (bad ? 0 : slot_base);
      const bool heavy = kids > 8;  // light parents write their own children; heavy parents are spread over the warp
      if (!heavy) {
        for (long long k = 0; k < kids; ++k) {
          const long long sl = dst + k;
          if (sl >= cap) {
            p.flags[0] = 1;
            break;
          }
          store_pose_stream(p.soa_next, p.stride, sl, P);
          mt_sts(p.nn_next + sl, nn);
          if (p.anc) mt_sts(p.anc + sl, (int)i);
        }
      }
      unsigned hm = __ballot_sync(0xffffffffu, heavy);
      while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        float Q[3][4];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) Q[a][b] = __shfl_sync(0xffffffffu, P[a][b], src);
        const long long d0 = __shfl_sync(0xffffffffu, dst, src);
        const long long kn = __shfl_sync(0xffffffffu, kids, src);
        const int pn = __shfl_sync(0xffffffffu, nn, src);
        const long long pi = __shfl_sync(0xffffffffu, i, src);
        for (long long k = lane; k < kn; k += 32) {
          const long long sl = d0 + k;
          if (sl >= cap) {
            p.flags[0] = 1;
            break;
          }
          store_pose_stream(p.soa_next, p.stride, sl, Q);
          mt_sts(p.nn_next + sl, pn);
          if (p.anc) mt_sts(p.anc + sl, (int)pi);
        }
      }
#if MT_TRACE
      if (c == c_lo) MT_TRACE_MAX(p.xdbg, 28)
#endif
    }
  } else
#endif
  {
  double run = base_g;
#if MT_BW_PREFETCH
  // the poses of chunk c + 1 are requested before chunk c is scanned and scattered: the block's chunks are a chain of
  // dependent steps (scan -> slot counts -> scatter), and its next loads travel underneath
  ChunkIn nxt;
  if (cached && c_lo < c_hi) {
    const long long i = (long long)c_lo * MT_CHUNK + threadIdx.x;
    if (i < n) load_pose_stream(p.soa_cur, p.stride, i, nxt.P);
  }
#endif
  for (int c = c_lo; c < c_hi; ++c) {
    const double base = fmin(run, next_g);
    run += s_part[c - c_lo];
    // the block's last chunk, or the chunk holding the shard's last particle (sharded runs size the grid
    // from a host-side bound), ends exactly on the block boundary
    const double endv = (c + 1 == c_hi || (long long)(c + 1) * MT_CHUNK >= n) ? next_g : fmin(run, next_g);
    ChunkIn cur;
    if (cached) {  // only the pose comes from memory, and it is not needed before the scatter
      const long long i = (long long)c * MT_CHUNK + threadIdx.x;
      cur.nn = s_nn[(c - c_lo) * MT_CHUNK + threadIdx.x];
      cur.e = s_e[(c - c_lo) * MT_CHUNK + threadIdx.x];
#if MT_BW_PREFETCH
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cur.P[a][b] = nxt.P[a][b];
      const long long i2 = i + MT_CHUNK;
      if (c + 1 < c_hi && i2 < n) load_pose_stream(p.soa_cur, p.stride, i2, nxt.P);
#else
      if (i < n) load_pose_stream(p.soa_cur, p.stride, i, cur.P);
#endif
    } else {
      step_b_load<true, true>(p, c, n, cur);
    }
    __syncthreads();  // s8 / s_cnt of the previous chunk are free
    step_b_chunk<true, true>(p, c, n, S, A, base, endv, s8, s_cnt, cur, slot_base);
#if MT_TRACE
    if (c == c_lo) MT_TRACE_MAX(p.xdbg, 28)
#endif
  }
  }
  MT_TRACE_END(p.xdbg, 4)
  if (g == 0) {  // RMSE, drift flag, bookkeeping (what the last block of k_step_sums does)
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    if (p.has_gt) {
      sa = block_sum_array_256(blocktot + G, G, 1, s8);
      sb = block_sum_array_256(blocktot + 2 * G, G, 1, s8);
    }
    int ctot = 0;
    for (int j = threadIdx.x; j < G; j += blockDim.x) ctot += __ldcg(blockcnt + j);
    __shared__ int s_on;
    if (threadIdx.x == 0) s_on = 0;
    __syncthreads();
    if (ctot) atomicAdd(&s_on, ctot);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (p.has_gt) {
        p.rmse2[0] = (float)sqrt(sa / (double)n);
        p.rmse2[1] = (float)sqrt(sb / (double)n);
      }
      if (n == 0 && p.n_out) *p.n_out = 0;  // a shard without particles has no children (nobody owns i == n - 1)
      p.prefix[p.nchunks] = s_bc[2];  // this GPU's weight sum
      p.scal[0] = s_bc[2];
      p.flags[6] = s_on;
      p.flags[5] = (p.prune_dist > 0.0 && s_on == 0);
      p.flags[3] += (int)p.qctl[0];
      p.flags[MT_STAT_MESH_DEFERRED] += (int)p.qctl[3];
      p.qctl[0] = 0, p.qctl[1] = 0, p.qctl[2] = 0, p.qctl[3] = 0;
    }
  }
}

// strictly sequential float64 prefix (parity mode): one thread reproduces torch.cumsum on
// CPU bit-for-bit, including slots the reference leaves unfilled (-1).
__global__ void k_resample_seq(StepDev p) {
  if (threadIdx.x || blockIdx.x) return;
  const long long N = p.n;
  const double dN = (double)N;
  const double off = (double)(p.u / (float)N);
  const double S = p.prefix[p.nchunks];
  if (!(S > 0.0) || !(S <= DBL_MAX)) {
    p.flags[1] = 1;
    for (long long i = 0; i < N; ++i) p.anc[i] = (int)i;
    return;
  }
  double run = 0.0;
  long long prev = 0;
  for (long long i = 0; i < N; ++i) {
    run += p.wsrc[i] / S;
    long long cnt = mt_count_below(run, N, dN, off);
    for (long long s = prev; s < cnt; ++s) p.anc[s] = (int)i;
    if (cnt > prev) prev = cnt;
  }
  for (long long s = prev; s < N; ++s) p.anc[s] = -1;
}

// particles the grids must cover: with a device-resident count (d_n_in) the host only knows the capacity
static long long step_cover(const mt_step_args* a) { return a->d_n_in ? a->stride : a->n; }
// blocks of k_step_a: one per tile of MT_A_BLOCK particles
static unsigned step_a_grid(const mt_ctx* c, long long cover) {
  long long g = (cover + MT_A_BLOCK - 1) / MT_A_BLOCK;
  (void)c;
  return (unsigned)std::max(g, 1ll);
}

static int fill_step(mt_ctx* c, const mt_step_args* a, StepDev* d) {
  if (!c || !a) return set_err(MT_ERR_ARG, "step: null argument");
  if (a->n <= 0 || (size_t)a->n > c->cap || a->stride < a->n) return set_err(MT_ERR_CAPACITY, "step: n/stride out of range");
  if (a->d_n_in && (size_t)a->stride > c->cap) return set_err(MT_ERR_CAPACITY, "step: with a device-resident count the stride must not exceed the context capacity");
  if (a->n > 0x0fffffffLL) return set_err(MT_ERR_CAPACITY, "step: at most 2^28 - 1 particles per GPU (queue entries carry flag bits)");
  memset(d, 0, sizeof(*d));
  d->soa_cur = (float4*)a->d_soa_cur;
  d->soa_next = (float4*)a->d_soa_next;
  d->stride = a->stride;
  d->nn_cur = a->d_nn_cur;
  d->nn_next = a->d_nn_next;
  d->anc = a->d_anc;
  d->n = a->n;
  d->odom = affine_from_host16(a->odom);
  d->tn = a->d_tn;
  d->rot = a->d_rot;
  d->sig_t = a->sig_t;
  d->sig_r = a->sig_r;
  d->seed = a->seed;
  d->step = a->step;
  d->first_gid = a->first_gid;
  d->wtab = a->softmax ? c->d_esim : c->d_sim;
  d->u = a->u;
  d->has_gt = a->gt != nullptr;
  if (a->gt) d->gt = affine_from_host16(a->gt);
  d->rmse2 = a->d_rmse2;
  d->rank = a->world > 1 ? a->rank : 0;
  d->world = a->world > 1 ? a->world : 1;
  // single GPU: n_global > 0 sets the number of children to draw (annealing changes the particle count,
  // particle_filter.py:405-447); 0 keeps it at n
  d->n_global = (a->world > 1 || a->n_global > 0) ? a->n_global : a->n;
  d->shard_sums = a->d_shard_sums;
  d->peers = (c->peers_world == d->world && d->world > 1) ? c->d_peers : nullptr;
  d->xseq = c->xchg_count + 1;
  d->xdbg = c->d_xdbg;
  d->n_out = a->d_n_out;
  d->n_in = a->d_n_in;
  d->prune_dist = (c->mesh_ready && a->prune_dist > 0.0) ? a->prune_dist : 0.0;
  d->cb_poses = (const float4*)a->d_cb_poses;
  d->part = c->d_part;
  d->prefix = c->d_prefix;
  d->rm_part = c->d_rm_part;
  d->wpart = c->d_wpart;
  d->wrm = c->d_wrm;
  d->wcnt = c->d_wcnt;
  d->srec = c->d_rec;
  d->queue = c->d_queue;
  d->queue2 = c->d_queue2;
  d->queue_cap = (long long)c->cap + 32;
  d->qctl = c->d_qctl;
  d->scal = c->d_scal;
  d->ticket = c->d_ticket;
  d->flags = c->d_flags;
  d->nchunks = (int)nchunks_of(step_cover(a));
  return MT_OK;
}

extern "C" int mt_dist_export(mt_ctx* c, void* h_handle64) {
  if (!c || !h_handle64) return set_err(MT_ERR_ARG, "mt_dist_export: null");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->d_xchg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(h_handle64, &h, 64);
  return MT_OK;
}

extern "C" int mt_dist_import(mt_ctx* c, int rank, int world, const void* h_handles) {
  if (!c || !h_handles || world < 2 || world > MT_MAX_WORLD || rank < 0 || rank >= world)
    return set_err(MT_ERR_ARG, "mt_dist_import: bad argument (2 <= world <= 16)");
  CK(cudaSetDevice(c->device));
  for (int r = 0; r < c->peers_world; ++r) {  // re-import: drop the previous mappings
    if (c->h_peers[r] && c->h_peers[r] != c->d_xchg) cudaIpcCloseMemHandle(c->h_peers[r]);
    c->h_peers[r] = nullptr;
  }
  c->peers_world = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->h_peers[r] = c->d_xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)h_handles + 64 * r, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->h_peers[r] = (Xchg*)ptr;
  }
  CK(cudaMemcpy(c->d_peers, c->h_peers, sizeof(Xchg*) * world, cudaMemcpyHostToDevice));
  // Every rank (re)imports collectively, so the exchange sequence restarts everywhere: counter 0, buffer cleared (a
  // stale word could otherwise carry a sequence number that becomes valid again).  The caller must put a barrier
  // between this call and the first sharded step (FilterEngine.connect_peers: the all-reduce of the "ok" flag).
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(c->d_xchg, 0, sizeof(Xchg)));
  CK(cudaDeviceSynchronize());
  c->xchg_count = 0;
  c->peers_world = world;
  return MT_OK;
}

// single GPU, resampling requested, few enough chunks per block: k_step_sums + k_step_b run as one
// persistent cooperative kernel (k_step_bw)
static bool step_fused(mt_ctx* c, const mt_step_args* a) {
  if (!a->fuse_sums || !a->resample) return false;
  if (a->world > 1 && c->peers_world != a->world) return false;  // sharded: needs mt_dist_import
  if (c->bw_blocks_per_sm == 0) {
    int coop = 0, occ = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    if (coop) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step_bw, 256, 0);
    c->bw_blocks_per_sm = (coop && occ > 0) ? std::min(occ, 4) : -1;
  }
  if (c->bw_blocks_per_sm < 0) return false;
  // sharded: every rank must take the same decision, so it is taken on the capacity (equal on all ranks,
  // and chunks per block only grow with the particle count), not on this rank's current count
  const long long nb = (a->world > 1 || a->d_n_in) ? std::max(a->stride, a->n) : a->n;
  const int grid = std::min(std::min((int)nchunks_of(nb), c->sm_count * c->bw_blocks_per_sm), MT_BW_MAX_GRID);
  return (nchunks_of(nb) + grid - 1) / grid <= MT_BW_MAX_PER;
}


extern "C" int mt_dist_debug(mt_ctx* c, unsigned long long* h_out3) {
  if (!c || !h_out3) return set_err(MT_ERR_ARG, "mt_dist_debug: null");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(h_out3, c->d_xdbg, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost));
  return MT_OK;
}

extern "C" int mt_trace_read(mt_ctx* c, unsigned long long* h_out64, int reset) {
  if (!c || !h_out64) return set_err(MT_ERR_ARG, "mt_trace_read: null");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_out64, c->d_xdbg, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost));
  if (reset) {  // "earliest" words (even trace slots, word 25) restart at the maximum, "latest" words at 0
    unsigned long long init[64];
    memcpy(init, h_out64, sizeof(init));
    for (int k = 8; k < 64; ++k) init[k] = 0;
    for (int k = 8; k < 24; k += 2) init[k] = ~0ull;
    init[25] = ~0ull;
    CK(cudaMemcpy(c->d_xdbg, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  h_out64[63] = MT_TRACE;
  return MT_OK;
}

extern "C" int mt_step_is_fused(mt_ctx* c, const mt_step_args* a, int* h_fused) {
  if (!c || !a || !h_fused) return set_err(MT_ERR_ARG, "mt_step_is_fused: null");
  *h_fused = step_fused(c, a) ? 1 : 0;
  return MT_OK;
}

extern "C" int mt_step_a(mt_ctx* c, const mt_step_args* a, void* stream) {
  StepDev d;
  int r = fill_step(c, a, &d);
  if (r) return r;
  if (!c->cb_ready) return set_err(MT_ERR_STATE, "mt_step_a: no codebook");
  if (!a->d_soa_cur || !a->d_nn_cur) return set_err(MT_ERR_ARG, "mt_step_a: null particle buffers");
  if ((a->d_tn == nullptr) != (a->d_rot == nullptr)) return set_err(MT_ERR_ARG, "mt_step_a: tn/rot must both be given");
  if (a->gt && !a->d_rmse2) return set_err(MT_ERR_ARG, "mt_step_a: gt without rmse output");
  if (a->prune_dist > 0.0 && !c->mesh_ready) return set_err(MT_ERR_STATE, "mt_step_a: prune_dist given but no mesh uploaded");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->timing[0]) CK(cudaEventRecord(c->timing[0], st));
  k_step_a<<<step_a_grid(c, step_cover(a)), MT_A_BLOCK, 0, st>>>(d, tables_of(c), mesh_of(c));
  CK_LAUNCH();
  if (c->timing[1]) CK(cudaEventRecord(c->timing[1], st));
  if (d.prune_dist > 0.0) {  // (a particle is in at most one of the two queues)
    k_step_meshq<<<c->sm_count * 4, 256, 0, st>>>(d, mesh_of(c));
    CK_LAUNCH();
    k_step_meshq2<<<c->sm_count * 8, 256, 0, st>>>(d, mesh_of(c));
    CK_LAUNCH();
  }
  k_step_nnq<<<c->sm_count * 12, 32 * MT_NNQ_WARPS, 0, st>>>(d, tables_of(c), mesh_of(c));
  CK_LAUNCH();
  if (c->timing[2]) CK(cudaEventRecord(c->timing[2], st));
  if (!step_fused(c, a)) {  // otherwise the sums are folded into mt_step_b's kernel
    if (a->table_ready_event) CK(cudaStreamWaitEvent(st, (cudaEvent_t)a->table_ready_event, 0));
    k_step_sums<<<d.nchunks, MT_CHUNK, 0, st>>>(d);
    CK_LAUNCH();
  }
  if (c->timing[3]) CK(cudaEventRecord(c->timing[3], st));
  return MT_OK;
}

extern "C" int mt_step_local_sum_ptr(mt_ctx* c, double** d_sum) {
  if (!c || !d_sum) return set_err(MT_ERR_ARG, "mt_step_local_sum_ptr: null");
  *d_sum = c->d_scal;
  return MT_OK;
}

extern "C" int mt_step_b(mt_ctx* c, const mt_step_args* a, void* stream) {
  StepDev d;
  int r = fill_step(c, a, &d);
  if (r) return r;
  if (!a->d_soa_cur || !a->d_soa_next || !a->d_nn_cur || !a->d_nn_next) return set_err(MT_ERR_ARG, "mt_step_b: null particle buffers");
  if (a->world > 1 && a->n_global <= 0) return set_err(MT_ERR_ARG, "mt_step_b: sharded step needs n_global");
  if (a->world > 1 && !step_fused(c, a) && !a->d_shard_sums) return set_err(MT_ERR_ARG, "mt_step_b: sharded step needs shard sums");
  cudaStream_t st = (cudaStream_t)stream;
  if (step_fused(c, a)) {
    if (a->table_ready_event) CK(cudaStreamWaitEvent(st, (cudaEvent_t)a->table_ready_event, 0));
    const int grid = std::min(std::min(d.nchunks, c->sm_count * c->bw_blocks_per_sm), MT_BW_MAX_GRID);
    unsigned long long* bar = c->d_bar;
    unsigned long long target = c->bar_target + (unsigned long long)grid;  // the barrier counter is monotone
    double* bw = c->d_bw;
    int* bwc = c->d_bwcnt;
    void* args[] = {&d, &bar, &target, &bw, &bwc};
    CK(cudaLaunchCooperativeKernel((const void*)k_step_bw, dim3(grid), dim3(256), args, 0, st));
    c->bar_target = target;  // only once the launch was accepted
    if (d.world > 1) c->xchg_count += 1;
    return MT_OK;
  }
  k_step_b<true, true><<<d.nchunks, MT_CHUNK, 0, st>>>(d);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- one call per step, CUDA graph
// mt_step = codebook query + k_step_a + queue consumers + the cooperative resampling kernel.  Stream form: the query
// runs on the context's side stream, concurrently with k_step_a.  Graph form: the same kernels as nodes of a CUDA
// graph (query | a -> {meshq -> meshq2, nnq} -> bw), instantiated once per configuration (buffer parity) and
// replayed; the by-value kernel arguments that change every step (odometry, noise key, offset, ground truth,
// barrier / exchange counters) are patched into the instantiated graph with cudaGraphExecKernelNodeSetParams.
#define MT_STEP_GRAPHS 4
struct StepGraph {
  bool valid;
  unsigned long long key[20];
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaGraphNode_t n_query, n_a, n_meshq, n_meshq2, n_nnq, n_bw;
  bool has_mesh;
};
static void step_graphs_free(mt_ctx* c) {
  if (!c->graphs) return;
  for (int k = 0; k < MT_STEP_GRAPHS; ++k)
    if (c->graphs[k].valid) {
      cudaGraphExecDestroy(c->graphs[k].exec);
      cudaGraphDestroy(c->graphs[k].graph);
    }
  delete[] c->graphs;
  c->graphs = nullptr;
}

struct QueryLaunch {
  const void* func;
  int grid;
  size_t smem;
  const void* q;
  const void* E;
  const double* rnorm;
  int M, D;
  double *sim, *esim, *sim2;
};
// launch parameters of k_codebook_query for this context (row norms are computed on `st` if they are not cached yet)
static int query_plan(mt_ctx* c, const void* d_q, int q_dtype, QueryLaunch* Q, cudaStream_t st) {
  if (!c->cb_ready || !c->d_emb) return set_err(MT_ERR_STATE, "mt_step: no codebook");
  if (!d_q) return set_err(MT_ERR_ARG, "mt_step: null query");
  const int D = c->D, M = c->M;
  if (D > MT_MAX_D) return set_err(MT_ERR_ARG, "cosine: D exceeds MT_MAX_D (6144)");
  if (q_dtype != MT_DTYPE_F32 && q_dtype != MT_DTYPE_F64) return set_err(MT_ERR_ARG, "cosine: bad query dtype");
  const bool e32 = c->emb_dtype == MT_DTYPE_F32, q32 = q_dtype == MT_DTYPE_F32;
  if (e32 && D % 4) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 4 for float32 rows");
  if (!e32 && D % 2) return set_err(MT_ERR_ARG, "cosine: D must be a multiple of 2 for float64 rows");
  if (!c->rnorm_ready) {
    if (e32)
      k_row_norms<float><<<(M + 7) / 8, 256, 0, st>>>((const float*)c->d_emb, M, D, c->d_rnorm);
    else
      k_row_norms<double><<<(M + 7) / 8, 256, 0, st>>>((const double*)c->d_emb, M, D, c->d_rnorm);
    CK_LAUNCH();
    c->rnorm_ready = true;
  }
  Q->smem = sizeof(double) * D;
  Q->func = e32 ? (q32 ? (const void*)k_codebook_query<float, float> : (const void*)k_codebook_query<float, double>)
                : (q32 ? (const void*)k_codebook_query<double, float> : (const void*)k_codebook_query<double, double>);
  if (!c->query_blocks_per_sm) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, Q->func, 256, Q->smem);
    c->query_blocks_per_sm = occ > 0 ? occ : 4;
  }
  const int resident = c->sm_count * c->query_blocks_per_sm;
  const int trips_total = (M + 3) / 4;
  int grid = (trips_total + 7) / 8;
  for (int k = 2; grid > resident; ++k) grid = ((trips_total + k - 1) / k + 7) / 8;
  Q->grid = grid;
  Q->q = d_q, Q->E = c->d_emb, Q->rnorm = c->d_rnorm, Q->M = M, Q->D = D;
  Q->sim = c->d_sim, Q->esim = c->d_esim, Q->sim2 = nullptr;
  return MT_OK;
}


extern "C" int mt_step(mt_ctx* c, const mt_step_args* a, const void* d_q, int q_dtype, int use_graph, void* stream) {
  if (!c || !a) return set_err(MT_ERR_ARG, "mt_step: null argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = a->resample && step_fused(c, a);
  if (a->world > 1 && a->resample && !fused)
    return set_err(MT_ERR_STATE, "mt_step: sharded steps need mt_dist_import (or mt_step_a + all-gather + mt_step_b)");
  if (!use_graph || !fused) {
    CK(cudaEventRecord(c->ev_fork, st));
    CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    int r = mt_codebook_query(c, d_q, q_dtype, nullptr, c->side);
    if (r) return r;
    CK(cudaEventRecord(c->ev_join, c->side));
    mt_step_args b = *a;
    b.table_ready_event = c->ev_join;
    r = mt_step_a(c, &b, st);
    if (r) return r;
    if (a->resample) return mt_step_b(c, &b, st);
    CK(cudaStreamWaitEvent(st, c->ev_join, 0));
    return MT_OK;
  }
  // ---- graph form
  StepDev d;
  int r = fill_step(c, a, &d);
  if (r) return r;
  if (!a->d_soa_cur || !a->d_soa_next || !a->d_nn_cur || !a->d_nn_next) return set_err(MT_ERR_ARG, "mt_step: null particle buffers");
  if ((a->d_tn == nullptr) != (a->d_rot == nullptr)) return set_err(MT_ERR_ARG, "mt_step: tn/rot must both be given");
  if (a->gt && !a->d_rmse2) return set_err(MT_ERR_ARG, "mt_step: gt without rmse output");
  if (a->prune_dist > 0.0 && !c->mesh_ready) return set_err(MT_ERR_STATE, "mt_step: prune_dist given but no mesh uploaded");
  if (a->world > 1 && a->n_global <= 0) return set_err(MT_ERR_ARG, "mt_step: sharded step needs n_global");
  QueryLaunch Q;
  r = query_plan(c, d_q, q_dtype, &Q, st);
  if (r) return r;
  NNTables T = tables_of(c);
  MeshTables Mh = mesh_of(c);
  const long long cover = step_cover(a);
  const int grid_bw = std::min(std::min((int)nchunks_of(cover), c->sm_count * c->bw_blocks_per_sm), MT_BW_MAX_GRID);
  unsigned long long* bar = c->d_bar;
  unsigned long long target = c->bar_target + (unsigned long long)grid_bw;
  double* bw = c->d_bw;
  int* bwc = c->d_bwcnt;
  void* args_q[] = {&Q.q, &Q.E, &Q.rnorm, &Q.M, &Q.D, &Q.sim, &Q.esim, &Q.sim2};
  void* args_a[] = {&d, &T, &Mh};
  void* args_m[] = {&d, &Mh};
  void* args_bw[] = {&d, &bar, &target, &bw, &bwc};
  const bool has_mesh = d.prune_dist > 0.0;
  const void* f_a = (const void*)k_step_a;
  const dim3 grid_a(step_a_grid(c, cover)), block_a(MT_A_BLOCK);
  const size_t smem_a = 0;
  // configuration key: everything that is baked into the graph's topology or launch geometry
  unsigned long long key[20] = {(unsigned long long)a->d_soa_cur, (unsigned long long)a->d_soa_next, (unsigned long long)a->d_nn_cur,
                                (unsigned long long)a->d_nn_next, (unsigned long long)a->d_anc, (unsigned long long)a->stride,
                                (unsigned long long)cover, (unsigned long long)has_mesh, (unsigned long long)a->world,
                                (unsigned long long)d_q, (unsigned long long)q_dtype, (unsigned long long)c->d_emb,
                                (unsigned long long)c->emb_dtype, (unsigned long long)grid_bw, (unsigned long long)Q.grid,
                                (unsigned long long)c->M, (unsigned long long)c->D, 0, 0, 0};
  if (!c->graphs) {
    c->graphs = new StepGraph[MT_STEP_GRAPHS];
    memset(c->graphs, 0, sizeof(StepGraph) * MT_STEP_GRAPHS);
  }
  StepGraph* G = nullptr;
  for (int k = 0; k < MT_STEP_GRAPHS; ++k)
    if (c->graphs[k].valid && !memcmp(c->graphs[k].key, key, sizeof(key))) G = &c->graphs[k];
  auto kparams = [](const void* f, dim3 g, dim3 b, size_t sm, void** args) {
    cudaKernelNodeParams p;
    memset(&p, 0, sizeof(p));
    p.func = (void*)f, p.gridDim = g, p.blockDim = b, p.sharedMemBytes = (unsigned)sm, p.kernelParams = args, p.extra = nullptr;
    return p;
  };
  cudaKernelNodeParams P_q = kparams(Q.func, dim3(Q.grid), dim3(256), Q.smem, args_q);
  cudaKernelNodeParams P_a = kparams(f_a, grid_a, block_a, smem_a, args_a);
  cudaKernelNodeParams P_m1 = kparams((const void*)k_step_meshq, dim3(c->sm_count * 4), dim3(256), 0, args_m);
  cudaKernelNodeParams P_m2 = kparams((const void*)k_step_meshq2, dim3(c->sm_count * 8), dim3(256), 0, args_m);
  cudaKernelNodeParams P_n = kparams((const void*)k_step_nnq, dim3(c->sm_count * 12), dim3(32 * MT_NNQ_WARPS), 0, args_a);
  cudaKernelNodeParams P_b = kparams((const void*)k_step_bw, dim3(grid_bw), dim3(256), 0, args_bw);
  if (!G) {
    G = &c->graphs[c->graph_next];
    c->graph_next = (c->graph_next + 1) % MT_STEP_GRAPHS;
    if (G->valid) {
      cudaGraphExecDestroy(G->exec);
      cudaGraphDestroy(G->graph);
      G->valid = false;
    }
    CK(cudaGraphCreate(&G->graph, 0));
    CK(cudaGraphAddKernelNode(&G->n_query, G->graph, nullptr, 0, &P_q));
    CK(cudaGraphAddKernelNode(&G->n_a, G->graph, nullptr, 0, &P_a));
    cudaGraphNode_t* after_a = &G->n_a;
    std::vector<cudaGraphNode_t> deps_bw = {G->n_query};
    if (has_mesh) {
      CK(cudaGraphAddKernelNode(&G->n_meshq, G->graph, after_a, 1, &P_m1));
      CK(cudaGraphAddKernelNode(&G->n_meshq2, G->graph, &G->n_meshq, 1, &P_m2));
      deps_bw.push_back(G->n_meshq2);
    }
    CK(cudaGraphAddKernelNode(&G->n_nnq, G->graph, after_a, 1, &P_n));
    deps_bw.push_back(G->n_nnq);
    CK(cudaGraphAddKernelNode(&G->n_bw, G->graph, deps_bw.data(), deps_bw.size(), &P_b));
    cudaKernelNodeAttrValue coop;
    memset(&coop, 0, sizeof(coop));
    coop.cooperative = 1;
    CK(cudaGraphKernelNodeSetAttribute(G->n_bw, cudaKernelNodeAttributeCooperative, &coop));
    CK(cudaGraphInstantiate(&G->exec, G->graph, 0));
    memcpy(G->key, key, sizeof(key));
    G->has_mesh = has_mesh;
    G->valid = true;
  } else {
    CK(cudaGraphExecKernelNodeSetParams(G->exec, G->n_a, &P_a));
    if (G->has_mesh) {
      CK(cudaGraphExecKernelNodeSetParams(G->exec, G->n_meshq, &P_m1));
      CK(cudaGraphExecKernelNodeSetParams(G->exec, G->n_meshq2, &P_m2));
    }
    CK(cudaGraphExecKernelNodeSetParams(G->exec, G->n_nnq, &P_n));
    CK(cudaGraphExecKernelNodeSetParams(G->exec, G->n_bw, &P_b));
  }
  CK(cudaGraphLaunch(G->exec, st));
  c->bar_target = target;
  if (d.world > 1) c->xchg_count += 1;
  c->graph_replays += 1;
  return MT_OK;
}

extern "C" int mt_step_graph_info(mt_ctx* c, long long* h_replays, int* h_cached) {
  if (!c) return set_err(MT_ERR_ARG, "mt_step_graph_info: null");
  if (h_replays) *h_replays = (long long)c->graph_replays;
  if (h_cached) {
    int n = 0;
    for (int k = 0; c->graphs && k < MT_STEP_GRAPHS; ++k) n += c->graphs[k].valid;
    *h_cached = n;
  }
  return MT_OK;
}

__global__ void k_step_weights(StepDev p, double* __restrict__ w) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= step_count(p)) return;
  double S = 0.0;
  if (p.world > 1)
    for (int r = 0; r < p.world; ++r) S += p.shard_sums[r];
  else
    S = p.prefix[p.nchunks];
  const int stored = p.nn_cur[i];
  w[i] = nn_is_masked(stored) ? 0.0 / S : __ldg(p.wtab + nn_index(stored)) / S;
}

extern "C" int mt_step_weights(mt_ctx* c, const mt_step_args* a, double* d_w, void* stream) {
  StepDev d;
  int r = fill_step(c, a, &d);
  if (r) return r;
  if (!d_w || !a->d_nn_cur) return set_err(MT_ERR_ARG, "mt_step_weights: null");
  k_step_weights<<<(unsigned)((step_cover(a) + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, d_w);
  CK_LAUNCH();
  return MT_OK;
}

extern "C" int mt_resample_systematic(mt_ctx* c, const double* d_w, long long n, float u, int seq, int32_t* d_anc,
                                      int* d_status, void* stream) {
  if (!c || !d_w || !d_anc || n <= 0) return set_err(MT_ERR_ARG, "mt_resample_systematic: bad argument");
  if ((size_t)n > c->cap) return set_err(MT_ERR_CAPACITY, "mt_resample_systematic: n exceeds context capacity");
  cudaStream_t st = (cudaStream_t)stream;
  StepDev d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.n_global = n;
  d.stride = n;
  d.world = 1;
  d.wsrc = d_w;
  d.u = u;
  d.anc = d_anc;
  d.part = c->d_part;
  d.prefix = c->d_prefix;
  d.scal = c->d_scal + 4;
  d.ticket = c->d_ticket;
  d.flags = c->d_flags;
  d.nchunks = (int)nchunks_of(n);
  CK(cudaMemsetAsync(c->d_flags + 1, 0, sizeof(int), st));
  k_weight_sums<<<d.nchunks, MT_CHUNK, 0, st>>>(d);
  CK_LAUNCH();
  if (seq)
    k_resample_seq<<<1, 32, 0, st>>>(d);
  else
    k_step_b<false, false><<<d.nchunks, MT_CHUNK, 0, st>>>(d);
  CK_LAUNCH();
  if (d_status) CK(cudaMemcpyAsync(d_status, c->d_flags + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return MT_OK;
}

// ------------------------------------------------------------------------- multinomial resampling
// resampler(..., "weighted_random") (particle_filter.py:243-250: WeightedRandomSampler = torch.multinomial with
// replacement): n_draws independent categorical draws.  The inclusive CDF is built with the chunk prefix of
// k_weight_sums plus a strictly sequential sum inside each 256-chunk (so it is monotone and an item of zero weight
// adds exactly nothing -- pruned particles can never be drawn), pinned to the next chunk's base at its end.
__global__ void __launch_bounds__(256) k_cdf_chunks(StepDev p, double* __restrict__ cdf) {
  __shared__ double s_w[MT_CHUNK];
  const int c = blockIdx.x;
  const long long i = (long long)c * MT_CHUNK + threadIdx.x;
  s_w[threadIdx.x] = (i < p.n) ? p.wsrc[i] : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = p.prefix[c];
    const double cap = p.prefix[c + 1];  // == the next chunk's base
    int last_pos = -1;
    for (int k = 0; k < MT_CHUNK; ++k) {
      if (s_w[k] > 0.0) last_pos = k;
      run += s_w[k];
      s_w[k] = fmin(run, cap);
    }
    // the chunk ends exactly on the next chunk's base, and it gets there on its last item of positive weight:
    // the CDF is flat across every zero-weight item, also across chunk boundaries
    if (last_pos >= 0)
      for (int k = last_pos; k < MT_CHUNK; ++k) s_w[k] = cap;
  }
  __syncthreads();
  if (i < p.n) cdf[i] = s_w[threadIdx.x];
}

__global__ void k_multinomial(StepDev p, const double* __restrict__ cdf, long long n_draws, int32_t* __restrict__ idx) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_draws) return;
  const double S = p.prefix[p.nchunks];
  if (!(S > 0.0) || !(S <= DBL_MAX)) {  // all-zero / NaN / Inf weights: the caller keeps the particles (237-241)
    if (j == 0) p.flags[1] = 1;
    idx[j] = (int32_t)(j < p.n ? j : p.n - 1);
    return;
  }
  const mt_u4 ctr = {(uint32_t)j, (uint32_t)((uint64_t)j >> 32), (uint32_t)p.step, (uint32_t)(p.step >> 32)};
  const mt_u4 r = mt_philox(ctr, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  idx[j] = (int32_t)mt_cdf_draw(cdf, p.n, S, mt_u01_53(r.x, r.y));
}

extern "C" int mt_resample_multinomial(mt_ctx* c, const double* d_w, long long n, long long n_draws, uint64_t seed,
                                       uint64_t stream_id, double* d_cdf_scratch, int32_t* d_idx, int* d_status, void* stream) {
  if (!c || !d_w || !d_idx || !d_cdf_scratch || n <= 0 || n_draws <= 0)
    return set_err(MT_ERR_ARG, "mt_resample_multinomial: bad argument");
  if ((size_t)n > c->cap) return set_err(MT_ERR_CAPACITY, "mt_resample_multinomial: n exceeds context capacity");
  if (n > 0x7fffffffLL) return set_err(MT_ERR_ARG, "mt_resample_multinomial: indices are int32");
  cudaStream_t st = (cudaStream_t)stream;
  StepDev d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.n_global = n;
  d.stride = n;
  d.world = 1;
  d.wsrc = d_w;
  d.seed = seed;
  d.step = stream_id;
  d.part = c->d_part;
  d.prefix = c->d_prefix;
  d.scal = c->d_scal + 4;
  d.ticket = c->d_ticket;
  d.flags = c->d_flags;
  d.nchunks = (int)nchunks_of(n);
  CK(cudaMemsetAsync(c->d_flags + 1, 0, sizeof(int), st));
  k_weight_sums<<<d.nchunks, MT_CHUNK, 0, st>>>(d);
  CK_LAUNCH();
  k_cdf_chunks<<<d.nchunks, MT_CHUNK, 0, st>>>(d, d_cdf_scratch);
  CK_LAUNCH();
  k_multinomial<<<(unsigned)((n_draws + 255) / 256), 256, 0, st>>>(d, d_cdf_scratch, n_draws, d_idx);
  CK_LAUNCH();
  if (d_status) CK(cudaMemcpyAsync(d_status, c->d_flags + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return MT_OK;
}

// ------------------------------------------------------------------------- batched codebook query (tensor cores)
#include "mt_gemm_tma.cuh"

static int ensure_scratch(mt_ctx* c, size_t bytes);
extern "C" int mt_codebook_query_batched(mt_ctx* c, const float* d_Q, int nq, float* d_out, void* stream) {
  if (!c || !c->cb_ready || !c->d_emb) return set_err(MT_ERR_STATE, "mt_codebook_query_batched: no codebook");
  if (!d_Q || !d_out || nq <= 0) return set_err(MT_ERR_ARG, "mt_codebook_query_batched: bad argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = c->M, D = c->D;
  const bool e32 = c->emb_dtype == MT_DTYPE_F32;
  if (D % 4) return set_err(MT_ERR_ARG, "mt_codebook_query_batched: D must be a multiple of 4 (16-byte rows)");
  if (!c->rnorm_ready) {
    if (e32)
      k_row_norms<float><<<(M + 7) / 8, 256, 0, st>>>((const float*)c->d_emb, M, D, c->d_rnorm);
    else
      k_row_norms<double><<<(M + 7) / 8, 256, 0, st>>>((const double*)c->d_emb, M, D, c->d_rnorm);
    CK_LAUNCH();
    c->rnorm_ready = true;
  }
  if (!c->planes_ready) {  // split planes of the codebook: once per upload
    cudaFree(c->d_plane_big), cudaFree(c->d_plane_small);
    c->d_plane_big = c->d_plane_small = nullptr;
    CK(cudaMalloc(&c->d_plane_big, sizeof(float) * (size_t)M * D));
    CK(cudaMalloc(&c->d_plane_small, sizeof(float) * (size_t)M * D));
    if (e32)
      k_split_planes<float><<<(M + 7) / 8, 256, 0, st>>>((const float*)c->d_emb, M, D, c->d_plane_big, c->d_plane_small, nullptr);
    else
      k_split_planes<double><<<(M + 7) / 8, 256, 0, st>>>((const double*)c->d_emb, M, D, c->d_plane_big, c->d_plane_small, nullptr);
    CK_LAUNCH();
    if (!g2_make_map(&c->tm_a_big, c->d_plane_big, M, D) || !g2_make_map(&c->tm_a_small, c->d_plane_small, M, D))
      return set_err(MT_ERR_CUDA, "mt_codebook_query_batched: cuTensorMapEncodeTiled failed");
    c->planes_ready = true;
  }
  // planes + norms of the queries (per call)
  const size_t qbytes = sizeof(float) * (size_t)nq * D;
  int r = ensure_scratch(c, 2 * qbytes + sizeof(float) * (size_t)nq + 256);
  if (r) return r;
  float* q_big = (float*)c->d_scratch;
  float* q_small = q_big + (size_t)nq * D;
  float* qinv = q_small + (size_t)nq * D;
  k_split_planes<float><<<(nq + 7) / 8, 256, 0, st>>>(d_Q, nq, D, q_big, q_small, qinv);
  CK_LAUNCH();
  CUtensorMap tm_b_big, tm_b_small;
  if (!g2_make_map(&tm_b_big, q_big, nq, D) || !g2_make_map(&tm_b_small, q_small, nq, D))
    return set_err(MT_ERR_CUDA, "mt_codebook_query_batched: cuTensorMapEncodeTiled failed");
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_codebook_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
    attr_set = true;
  }
  const int ntiles = ((M + G2_BM - 1) / G2_BM) * ((nq + G2_BN - 1) / G2_BN);
  k_codebook_gemm_tma<<<std::min(ntiles, c->sm_count), G2_THREADS, G2_SMEM_BYTES, st>>>(c->tm_a_big, c->tm_a_small, tm_b_big, tm_b_small, c->d_rnorm,
                                                                                       qinv, M, D, nq, d_out);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- cluster centres / annealing
#include "mt_cluster.cuh"

static int ensure_scratch(mt_ctx* c, size_t bytes) {
  if (bytes <= c->scratch_bytes) return MT_OK;
  cudaFree(c->d_scratch);
  c->d_scratch = nullptr, c->scratch_bytes = 0;
  CK(cudaMalloc(&c->d_scratch, bytes));
  c->scratch_bytes = bytes;
  return MT_OK;
}

extern "C" int mt_cluster_centers(mt_ctx* c, const float* d_poses, const double* d_weights, const int32_t* d_labels, long long n,
                                  int K, int method, float* d_centers, float* d_stds, void* stream) {
  if (method != 0 && method != 1) return set_err(MT_ERR_ARG, "mt_cluster_centers: method 0 (quat_avg) or 1 (logmap)");
  if (!c || !d_poses || !d_weights || !d_labels || !d_centers || !d_stds || n <= 0 || K <= 0 || K > MT_MAX_CLUSTERS)
    return set_err(MT_ERR_ARG, "mt_cluster_centers: bad argument (1 <= K <= 16)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)((n + 255) / 256);
  const size_t mom = sizeof(double) * (size_t)nb * K * MT_CL_VALS, mm = sizeof(float) * (size_t)nb * K * 2;
  int r = ensure_scratch(c, mom + mm + sizeof(int) * MT_MAX_CLUSTERS + 64);
  if (r) return r;
  double* part = (double*)c->d_scratch;
  float* mpart = (float*)((char*)c->d_scratch + mom);
  int* uniform = (int*)((char*)c->d_scratch + mom + mm);
  k_cluster_minmax<<<nb, 256, 0, st>>>(d_weights, d_labels, n, K, mpart);
  CK_LAUNCH();
  k_cluster_minmax_final<<<K, 256, 0, st>>>(mpart, nb, K, uniform);
  CK_LAUNCH();
  k_cluster_moments<<<nb, 256, 0, st>>>((const float4*)d_poses, d_weights, d_labels, n, K, uniform, part, method);
  CK_LAUNCH();
  k_cluster_final<<<K, 256, 0, st>>>(part, nb, K, d_centers, d_stds, method);
  CK_LAUNCH();
  return MT_OK;
}

extern "C" int mt_select_k(mt_ctx* c, const double* d_w, long long n, long long k, int largest, int32_t* d_sel, int32_t* d_keep,
                           void* stream) {
  if (!c || !d_w || n <= 0 || k <= 0 || k > n) return set_err(MT_ERR_ARG, "mt_select_k: bad argument (1 <= k <= n)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)((n + 255) / 256);
  int r = ensure_scratch(c, sizeof(int) * 2 * (size_t)nb + 256 * sizeof(unsigned int) + 64);
  if (r) return r;
  unsigned long long* state = (unsigned long long*)c->d_scratch;
  unsigned int* hist = (unsigned int*)((char*)c->d_scratch + 32);
  int* blk = (int*)((char*)c->d_scratch + 32 + 256 * sizeof(unsigned int));
  const unsigned long long init[2] = {0ull, (unsigned long long)(k - 1)};
  CK(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned int), st));
  const int hgrid = std::min(nb, c->sm_count * 8);
  for (int pass = 0; pass < 8; ++pass) {
    k_select_hist<<<hgrid, 256, 0, st>>>(d_w, n, largest, pass, state, hist);
    CK_LAUNCH();
    k_select_pick<<<1, 32, 0, st>>>(hist, pass, state);
    CK_LAUNCH();
  }
  k_select_count<<<nb, 256, 0, st>>>(d_w, n, largest, state, blk);
  CK_LAUNCH();
  k_select_scan<<<1, 1024, 0, st>>>(blk, nb);
  CK_LAUNCH();
  k_select_scatter<<<nb, 256, 0, st>>>(d_w, n, largest, state, blk, d_sel, d_keep);
  CK_LAUNCH();
  return MT_OK;
}

// ------------------------------------------------------------------------- cluster_particles (DBSCAN)
#include "mt_dbscan.cuh"

extern "C" int mt_dbscan(mt_ctx* c, const float* d_poses, long long n, double eps, long long min_samples, long long* d_labels,
                         int* h_n_clusters, void* stream) {
  if (!c || !d_poses || !d_labels || n <= 0 || !(eps > 0.0) || min_samples < 1) return set_err(MT_ERR_ARG, "mt_dbscan: bad argument");
  if (n > 0x7fffffffLL) return set_err(MT_ERR_ARG, "mt_dbscan: indices are int32");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = (int)n, nb = (N + MT_DB_TILE - 1) / MT_DB_TILE;
  int r = ensure_scratch(c, sizeof(float4) * (size_t)N + 3 * sizeof(int) * (size_t)N + 64);
  if (r) return r;
  float4* pts = (float4*)c->d_scratch;
  int* la = (int*)(pts + N);
  int* lb = la + N;
  int* cid = lb + N;
  int* ctl = cid + N;  // [0] changed, [1] clusters
  DbTest t;
  t.eps2 = eps * eps;
  t.lo2 = (float)(t.eps2 * (1.0 - 1e-4));
  t.hi2 = (float)(t.eps2 * (1.0 + 1e-4)) + 1e-30f;
  k_db_points<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_poses, n, pts);
  CK_LAUNCH();
  const int ms = (int)std::min<long long>(min_samples, 0x7fffffffLL);
  k_db_count<<<nb, MT_DB_TILE, 0, st>>>(pts, N, t, ms, la);
  CK_LAUNCH();
  for (int it = 0; it < 4096; ++it) {  // label propagation + pointer jumping until a pass changes nothing
    CK(cudaMemsetAsync(ctl, 0, sizeof(int), st));
    k_db_propagate<<<nb, MT_DB_TILE, 0, st>>>(pts, N, t, la, lb, ctl);
    CK_LAUNCH();
    k_db_jump<<<(N + 255) / 256, 256, 0, st>>>(lb, N);
    CK_LAUNCH();
    std::swap(la, lb);
    int changed = 0;
    CK(cudaMemcpyAsync(&changed, ctl, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!changed) break;
  }
  k_db_number<<<1, 1024, 0, st>>>(la, N, cid, ctl + 1);
  CK_LAUNCH();
  k_db_final<<<nb, MT_DB_TILE, 0, st>>>(pts, N, t, la, cid, d_labels);
  CK_LAUNCH();
  if (h_n_clusters) {
    CK(cudaMemcpyAsync(h_n_clusters, ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return MT_OK;
}

// ------------------------------------------------------------------------- tactile code network
#include "mt_tcn.cuh"
