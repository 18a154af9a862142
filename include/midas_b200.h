/* libmidas_b200 -- C ABI of the B200-native particle-filter hot path for MidasTouch.
 *
 * Every entry point replaces one Python-level call of the reference (paths relative to
 * the reference tree); the reference has no FFI of its own (it is pure Python over
 * torch), so the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (tensor.data_ptr());
 *     h_* is a host pointer.  The library owns only mt_ctx scratch.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden
 *     synchronisation unless stated.
 *   - return value: 0 = ok, negative = error (mt_last_error() gives the text).  Nothing
 *     throws or aborts.
 *   - particle poses inside the engine are SoA "3 x float4": three arrays of `stride`
 *     float4, array r holding row r of [R|t] = (R[r][0], R[r][1], R[r][2], t[r]);
 *     the constant bottom row of the reference's (N,4,4) tensor is implicit.
 *     d_soa points at 3*stride float4.  mt_aos_to_soa / mt_soa_to_aos convert from/to
 *     the reference layout (N,4,4) float32 row-major (particle_filter.py:33-58).
 */
#ifndef MIDAS_B200_H
#define MIDAS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mt_ctx mt_ctx;

#define MT_OK 0
#define MT_ERR_ARG -1
#define MT_ERR_CUDA -2
#define MT_ERR_STATE -3
#define MT_ERR_CAPACITY -4

#define MT_DTYPE_F32 0
#define MT_DTYPE_F64 1

const char* mt_last_error(void);
int mt_version(void);

/* ---- context ------------------------------------------------------------------- */
/* capacity = max particles resident on this GPU; M,D = codebook rows / embedding width. */
int mt_ctx_create(int device, size_t capacity, int M, int D, mt_ctx** out);
int mt_ctx_destroy(mt_ctx* ctx);

/* ---- codebook: tactile_tree.__init__ / init_tree (tactile_tree.py:13-41) ---------- */
/* h_keys: (M,6) float32 R3_SE3 keys on the HOST (a per-key neighbour graph plus a 32-ary
 * bounding-box hierarchy over the keys in 6-D Morton order replace the nanoflann tree); d_emb: (M,D) embeddings on the device in emb_dtype
 * (the reference stores float64, build_codebook.py:72-74).  The library keeps the pointer,
 * it does not copy the embeddings; their row norms are cached by the first query, so call
 * mt_codebook_upload again if the embeddings change. */
int mt_codebook_upload(mt_ctx* ctx, const float* h_keys, const void* d_emb, int emb_dtype);
/* search index introspection (tests): Morton cell edge, dims = {leaves, level-1 nodes, level-2 nodes},
 * occupied = leaves (32 keys each) */
int mt_codebook_grid_info(mt_ctx* ctx, float* h, int dims[3], int* occupied);

/* neighbour-graph introspection (tests): device pointer to the (M, k, 8) float32 table
 * [key(6), delta, index bits] of every key's k nearest other keys, ascending.  k = 64, or 128 / 256 on dense codebooks
 * (median 64th-neighbour distance below 1e-2 / 8e-3 key units; chosen by mt_codebook_upload, environment variable
 * MIDAS_B200_NBR_K = 64 | 128 | 256 overrides). */
int mt_codebook_nbr_info(mt_ctx* ctx, const float** d_nbr, int* k);
/* d_rank[m] = position of codebook row m in the library's spatial (grid-cell) order; sorting
 * particles by the rank of their match keeps neighbouring threads on neighbouring keys. */
int mt_codebook_rank(mt_ctx* ctx, int32_t* d_rank, void* stream);
/* instrumentation: four cudaEvent_t (before k_step_a, after it, after k_step_nnq, after
 * k_step_sums) recorded on the step's stream by every following mt_step_a; NULL switches it off. */
int mt_ctx_set_timing_events(mt_ctx* ctx, void* const* events4);
/* status / statistics words of the context (synchronises).  h_out[MT_STAT_*], MT_STAT_COUNT entries;
 * reset != 0 clears the cumulative slots (0..4, 7..). */
#define MT_STAT_COUNT 16
#define MT_STAT_OVERFLOW 0      /* 1: children did not fit the destination buffer (sharded steps);
                                   2: a peer's weight sum never arrived (fused sharded step timed out) */
#define MT_STAT_RESAMPLE_SKIP 1 /* a resampling saw all-zero / NaN weights and kept the particles */
#define MT_STAT_INVALID_POSES 2 /* poses check_quats would prune (cumulative) */
#define MT_STAT_NN_FALLBACKS 3  /* queries that left the hint graph for the box-hierarchy search (cumulative) */
#define MT_STAT_GRID_ROWS 4     /* leaves (32 keys each) visited by those searches (cumulative) */
#define MT_STAT_GRID_ROWS_MAX 7 /* most leaves visited by a single search */
#define MT_STAT_DRIFTED 5       /* last mt_step_a: every particle failed the drift test */
#define MT_STAT_ON_SURFACE 6    /* last mt_step_a: particles that passed the drift test */
#define MT_STAT_MESH_DEFERRED 8 /* drift tests whose voxel was undecided and that ran the vertex search (cumulative) */
#define MT_STAT_SCAN_DEFERRED 9 /* unused (always 0): the two-pass hint scan was removed */
int mt_ctx_stats(mt_ctx* ctx, long long* h_out, int reset);

/* ---- mesh: particle_filter.__init__ (particle_filter.py:108-110) ------------------- */
/* h_vertices: (V,3) float64 down-sampled mesh vertices (mesh.vertices[::10]) on the HOST; a
 * uniform grid of edge `cell` (>= the default invalid_dist, tdn.render.pen.max) replaces the
 * sklearn KDTree. */
int mt_mesh_upload(mt_ctx* ctx, const double* h_vertices, long long V, double cell);
/* remove_invalid_particles (particle_filter.py:379-403) on (n,4,4) float32 poses:
 * d_weights[i] *= (distance to the nearest vertex <= invalid_dist), float64, in place
 * (nullable); *d_num_valid = number of particles that passed (device int, nullable). */
int mt_prune_aos(mt_ctx* ctx, const float* d_poses, long long n, double invalid_dist, double* d_weights,
                 int* d_num_valid, void* stream);

/* cos(q, E_m) for all M rows -> ctx-resident float64 tables sim[M] and exp(sim)[M]
 * (get_similarity(code, heatmap_embeddings, softmax=False), filter.py:213-215, and the
 * per-particle weights of filter.py:170-173 by table lookup).  d_q: (D,) in q_dtype.
 * d_sim_out (nullable): (M,) float64 copy of the table. */
int mt_codebook_query(mt_ctx* ctx, const void* d_q, int q_dtype, double* d_sim_out, void* stream);
/* nq queries at once against the uploaded codebook: d_out[q*M + m] = cos(Q_q, E_m), float32
 * (the batched form of get_similarity, eval/single_touch_test.py:35-73).  tcgen05 tensor-core
 * GEMM with TMEM accumulators, 3xTF32 split operands (~1e-6 relative).  d_Q: (nq, D) float32. */
int mt_codebook_query_batched(mt_ctx* ctx, const float* d_Q, int nq, float* d_out, void* stream);
/* general form: cos(q, T_n) for an explicit (rows, D) target matrix
 * (get_similarity(queries, targets), particle_filter.py:449-457). */
int mt_cosine_rows(mt_ctx* ctx, const void* d_q, int q_dtype, const void* d_targets, int t_dtype, long long rows,
                   int D, double* d_out, void* stream);
/* Q queries at once: out[q*rows + m] = cos(Q_q, T_m) in float32 (the M x M retrieval of
 * eval/single_touch_test.py:35-73). */
int mt_cosine_batched(mt_ctx* ctx, const float* d_Q, int nq, const float* d_T, long long rows, int D, float* d_out,
                      void* stream);
/* softmax over n float64 scores unless (max-min) is within 1e-8 of 0, then a copy
 * (particle_filter.py:459-468).  d_out may alias d_in. */
int mt_softmax_f64(mt_ctx* ctx, const double* d_in, long long n, double* d_out, void* stream);

/* ---- layout converters --------------------------------------------------------- */
int mt_aos_to_soa(const float* d_aos, long long n, float* d_soa, long long stride, void* stream);
int mt_soa_to_aos(const float* d_soa, long long stride, long long n, float* d_aos, void* stream);

/* ---- SE3_NN (tactile_tree.py:43-58) --------------------------------------------- */
/* R3_SE3 keys of n poses -> d_keys (n,6) float32 */
int mt_se3_keys(const float* d_soa, long long stride, long long n, float* d_keys, void* stream);
/* the same with R3_SE3's weight w as an argument (tactile_tree.py:73: key = [(1 - w) t, w Log(R)]; mt_se3_keys is w = 0.01,
 * the value the codebook's own keys are built with) */
int mt_se3_keys_w(const float* d_soa, long long stride, long long n, double w, float* d_keys, void* stream);
/* SE3_NN(nn = k > 1) (tactile_tree.py:43-52): d_idx (n,k) int32 = the k <= 64 nearest codebook keys of every query key,
 * ascending (distance, index) like kneighbors(); exhaustive search, meant for small n */
int mt_nn_topk(mt_ctx* ctx, const float* d_keys, long long n, int k, int32_t* d_idx, void* stream);
/* exact L2 1-NN of n keys in the codebook; ties -> lowest index.  d_hint (nullable):
 * a codebook index per query that seeds the search bound (any valid index is correct).
 * mode 0 = hint graph + grid search (mt_nn.cuh), 1 = exhaustive (tiled).  d_idx: (n,) int32. */
int mt_nn_assign(mt_ctx* ctx, const float* d_keys, long long n, const int32_t* d_hint, int mode, int32_t* d_idx,
                 void* stream);
/* out[i] = table rows gathered by index: poses (M,4,4) f32 -> AoS (n,4,4) */
int mt_gather_rows_f32(const float* d_table, const int32_t* d_idx, long long n, int row_floats, float* d_out,
                       void* stream);

/* ---- motionModel (particle_filter.py:319-377) ----------------------------------- */
/* pose_n <- pose_n @ (odom @ [Rzyx(rot_n) | tn_n]).  h_odom: 16 floats (4,4) row-major on
 * the host.  d_tn/d_rot: (n,3) float32 noise (translation m / rotation deg) drawn by the
 * caller -- the reference's CPU RNG contract (326-335); if both are NULL the kernel draws
 * Philox4x32-10 normals keyed by (seed, step, first_gid+n) scaled by sig_t/sig_r.
 * d_invalid_count (nullable): incremented for poses check_quats would prune (347-357).
 * euler_mode 0: Rn = Rz Ry Rx (euler_angles_to_matrix "ZYX", motionModel); 1: Rn = Rx Ry Rz
 * (scipy from_euler("zyx"), init_filter at particle_filter.py:129-145). */
int mt_motion(const float* d_soa_in, float* d_soa_out, long long stride, long long n, const float* h_odom,
              const float* d_tn, const float* d_rot, float sig_t, float sig_r, uint64_t seed, uint64_t step,
              uint64_t first_gid, int* d_invalid_count, int euler_mode, void* stream);

/* ---- particle_rmse (particle_filter.py:472-496) --------------------------------- */
/* d_out2: {rmse_t, rmse_r} float32 on the device */
int mt_rmse(mt_ctx* ctx, const float* d_soa, long long stride, long long n, const float* h_gt, float* d_out2,
            void* stream);

/* ---- resampler("low_var") (particle_filter.py:230-261, 288-307) ------------------ */
/* systematic resampling of n float64 weights: normalise by their sum, inclusive float64
 * prefix, slot j <- first i with loc_j < C_i, loc_j = j/n + float32(u)/n.
 * d_anc: (n,) int32 ancestor of every slot.  seq != 0 selects the strictly sequential
 * float64 prefix (bit-identical to torch.cumsum on CPU; one thread, tests only).
 * d_status (nullable): set to 1 when the reference would skip resampling (all-zero or
 * NaN weights, 237-241); the ancestors are then the identity. */
int mt_resample_systematic(mt_ctx* ctx, const double* d_w, long long n, float u, int seq, int32_t* d_anc,
                           int* d_status, void* stream);

/* ---- resampler("weighted_random") (particle_filter.py:243-250) -------------------- */
/* n_draws independent categorical draws with replacement from n float64 weights (the reference's
 * WeightedRandomSampler / torch.multinomial(replacement=True)): d_idx[j] = first i with C_i > u_j * sum(w),
 * C the inclusive float64 prefix of the weights, u_j a 53-bit uniform from Philox4x32-10 keyed by
 * (seed; j, stream_id).  Items of zero weight are never drawn.  d_cdf_scratch: (n,) float64 work space owned by
 * the caller.  d_status as for mt_resample_systematic (indices are then the identity). */
int mt_resample_multinomial(mt_ctx* ctx, const double* d_w, long long n, long long n_draws, uint64_t seed,
                            uint64_t stream_id, double* d_cdf_scratch, int32_t* d_idx, int* d_status, void* stream);
int mt_gather_soa(const float* d_soa_in, long long stride_in, const int32_t* d_anc, long long n, float* d_soa_out,
                  long long stride_out, void* stream);
int mt_gather_f64(const double* d_in, const int32_t* d_anc, long long n, double* d_out, void* stream);

/* ---- fused engine step (filter.py:152-190 without prune/cluster/anneal) ----------- */
typedef struct mt_step_args {
  /* particle state, SoA, ping-pong */
  float* d_soa_cur;      /* in: poses at t-1 (moved in place) */
  float* d_soa_next;     /* out: resampled poses */
  long long stride;      /* float4 elements per row array, >= capacity */
  int32_t* d_nn_cur;     /* in: NN hint per particle (-1 = none); out: NN of the moved pose */
  int32_t* d_nn_next;    /* out: NN index inherited by every child (next step's hint) */
  int32_t* d_anc;        /* out (nullable): ancestor (local index) of every child */
  long long n;           /* particles on this GPU */
  /* motion */
  float odom[16];
  const float* d_tn;     /* nullable -> Philox */
  const float* d_rot;
  float sig_t, sig_r;
  uint64_t seed, step, first_gid;
  /* measurement */
  int softmax;           /* 1: w = exp(cos) (filter.py:171-173); 0: raw cosine (filter_real.py:208-210) */
  /* resampling */
  float u;               /* the single systematic offset torch.rand(1) (particle_filter.py:260) */
  int resample;          /* 0: stop after weighting (kernel A only) */
  /* metric (nullable gt): rmse of the moved poses vs gt (filter.py:164) */
  const float* gt;       /* host, 16 floats, or NULL */
  float* d_rmse2;        /* device, 2 floats */
  /* sharding: this GPU holds ranks' slice `rank` of `world`; d_shard_sums (world doubles,
   * device) holds every rank's weight sum, written by the caller's collective between
   * mt_step_a and mt_step_b; NULL for a single GPU. */
  int rank, world;
  long long n_global;    /* sharded: particles over all GPUs.  Single GPU: > 0 = number of children to draw (<= stride; the
                          * particle count changes under annealing, particle_filter.py:405-447), 0 = n */
  double* d_shard_sums;
  long long* d_n_out;    /* device (nullable): number of children written on this GPU */
  const long long* d_n_in; /* device (nullable): particle count read by the kernels instead of n (no host sync
                            * between steps); the grids then cover `stride` particles and n is informational.
                            * A count beyond stride raises MT_STAT_OVERFLOW. */
  /* drift pruning (remove_invalid_particles, filter.py:176-179): > 0 zeroes the weight of particles
   * further than this from the uploaded mesh; if all drift, mt_step_b re-projects the particles onto
   * d_cb_poses[(M,4,4) float32 codebook poses, nullable] instead of resampling. */
  double prune_dist;
  const float* d_cb_poses;
  /* optional cudaEvent_t: mt_step_a makes `stream` wait for it before the weight lookup (its third
   * kernel), so that mt_codebook_query may run concurrently on another stream with the motion /
   * SE3_NN kernels.  NULL: the query was enqueued on `stream` before mt_step_a. */
  void* table_ready_event;
  /* != 0 (single GPU, resample != 0): mt_step_a leaves the weight sums to mt_step_b, which then runs
   * sums + resampling as one persistent cooperative kernel; mt_step_weights / the local sum are only
   * valid after mt_step_b in that mode. */
  int fuse_sums;
} mt_step_args;

/* ---- sharded runs without a collective call ---------------------------------------------------
 * Every context owns a small exchange buffer.  mt_dist_export writes its CUDA IPC handle (64 bytes,
 * host) -- exchange the handles between the ranks (e.g. torch.distributed.all_gather) -- and
 * mt_dist_import(rank, world, handles[world][64]) maps the peers' buffers.  A step with fuse_sums != 0
 * then runs sums + exchange + resampling as ONE cooperative kernel per GPU: each GPU stores its weight
 * sum into its peers' buffers over NVLink and spins on its own buffer for theirs. */
int mt_dist_export(mt_ctx* ctx, void* h_handle64);
int mt_dist_import(mt_ctx* ctx, int rank, int world, const void* h_handles);
/* diagnostics of the last fused sharded step (synchronises): %globaltimer (ns) of block 0 at the end of the local
 * phase, after its sums were sent, and after all peers' sums had arrived */
int mt_dist_debug(mt_ctx* ctx, unsigned long long* h_out3);
/* diagnostics (synchronises): the context's 64-word timestamp buffer.  In a library built with -DMT_TRACE=1 the step
 * kernels record the earliest block start / latest block end (%globaltimer, ns) in words 8 + 2k / 9 + 2k
 * (k = 0 k_step_a, 1 k_step_meshq, 2 k_step_meshq2, 3 k_step_nnq, 4 k_step_bw) and k_step_bw its phase boundaries
 * in words 24..28; h_out64[63] = 1 for such a build, 0 for the shipped one (which records nothing).  reset != 0
 * re-arms the buffer for the next step. */
int mt_trace_read(mt_ctx* ctx, unsigned long long* h_out64, int reset);
/* *h_fused = 1 when mt_step_a/mt_step_b will run this step in the fused form (sums + exchange + resampling in
 * one cooperative kernel; a sharded caller then skips its all-gather of the weight sums), else 0 */
int mt_step_is_fused(mt_ctx* ctx, const mt_step_args* a, int* h_fused);

/* One filter step as ONE call: mt_codebook_query(d_q) concurrently with mt_step_a, then mt_step_b -- the loop body
 * filter.py:154-190 (motionModel, SE3_NN, get_similarity, remove_invalid_particles, resampler "low_var").
 * use_graph != 0 (and a step that runs in the fused form, see mt_step_is_fused): the kernels are nodes of a CUDA
 * graph (query | motion+SE3_NN -> queue consumers -> cooperative resampling kernel) that is instantiated once per
 * configuration (buffer parity, query pointer, grid sizes) and replayed; only the by-value arguments that change
 * from step to step are patched.  d_q must then be the same device buffer every step (copy the code into it).
 * a->table_ready_event is ignored.  Sharded steps (world > 1) need mt_dist_import; without it use
 * mt_step_a + all-gather + mt_step_b. */
int mt_step(mt_ctx* ctx, const mt_step_args* a, const void* d_q, int q_dtype, int use_graph, void* stream);
/* graph replays so far / instantiated graphs held by the context */
int mt_step_graph_info(mt_ctx* ctx, long long* h_replays, int* h_cached);

/* kernel A: motion + key + exact NN + weight lookup + deterministic weight sums */
int mt_step_a(mt_ctx* ctx, const mt_step_args* a, void* stream);
/* device pointer to this GPU's local weight sum (float64), valid after mt_step_a */
int mt_step_local_sum_ptr(mt_ctx* ctx, double** d_sum);
/* kernel B: normalise + prefix + systematic draw + scatter of children */
int mt_step_b(mt_ctx* ctx, const mt_step_args* a, void* stream);
/* normalised float64 weights of the current particles (after mt_step_a) */
int mt_step_weights(mt_ctx* ctx, const mt_step_args* a, double* d_w, void* stream);

/* ---- get_cluster_centers (particle_filter.py:153-206; xyz_quat_averaged pose.py:112-147, log_map_averaged 101-109) ---- */
/* d_poses (n,4,4) float32, d_weights (n,) float64 (cast to float32 like the reference), d_labels (n,)
 * int32 cluster ids 0..K-1 (negative = ignored).  method 0 = "quat_avg" (what filter.py:184-186 passes), 1 = "logmap"
 * (the reference's default: weighted mean of the SE(3) tangents, exponentiated).  d_centers (K,4,4), d_stds (K,3) float32. */
int mt_cluster_centers(mt_ctx* ctx, const float* d_poses, const double* d_weights, const int32_t* d_labels, long long n, int K,
                       int method, float* d_centers, float* d_stds, void* stream);
/* ---- cluster_particles(method="euclidean") (particle_filter.py:208-228) -------------------- */
/* sklearn DBSCAN(eps, min_samples) on the translations of d_poses (n,4,4) float32, with sklearn's semantics: float64
 * squared distance <= eps^2 (the point itself counts), clusters numbered by their lowest-index core point, border
 * points to the lowest-numbered cluster in reach, noise -1.  d_labels: (n,) int64.  h_n_clusters (host, nullable):
 * number of clusters.  Synchronises `stream` (the label propagation iterates until nothing changes). */
int mt_dbscan(mt_ctx* ctx, const float* d_poses, long long n, double eps, long long min_samples, long long* d_labels,
              int* h_n_clusters, void* stream);
/* ---- annealing (particle_filter.py:405-447): order statistics of the weights ---------------- */
/* the k smallest (largest != 0: largest) of n float64 weights: d_sel (k,) their indices in ascending
 * index order, d_keep (n-k,) the remaining indices in ascending order (either may be NULL).  Ties at
 * the threshold are resolved towards the lowest index. */
int mt_select_k(mt_ctx* ctx, const double* d_w, long long n, long long k, int largest, int32_t* d_sel, int32_t* d_keep,
                void* stream);

/* ---- TCN: tactile code network forward (contrib/tcn_minkloc/{tcn,minkloc,minkfpn}.py) ------- */
/* MinkLoc3D (sparse 3-D FPN + GeM) for the shipped topology (config/tcn/default.yaml): conv0 k5,
 * three stride-2 stages with one BasicBlock each, one top-down block, GeM, L2 normalisation.
 * Parameters are passed in MinkowskiEngine's layout: kernels (kvol, cin, cout) float32 with
 * offset index x-fastest; BatchNorm as (weight, bias, running_mean, running_var).
 * conv ids: 0 conv0 | 1+s convs[s] | 4+2s, 5+2s blocks[s].conv1/conv2 | 10+s blocks[s].downsample
 *           | 13 conv1x1[0] | 14 tconvs[0] | 15 conv1x1[1]
 * bn ids:   0 bn0 | 1+s bn[s] | 4+2s, 5+2s blocks[s].norm1/norm2 | 10+s blocks[s].downsample.1 */
typedef struct mt_tcn mt_tcn;
int mt_tcn_create(int device, int max_points, int max_batch, mt_tcn** out);
int mt_tcn_destroy(mt_tcn* t);
int mt_tcn_set_conv(mt_tcn* t, int conv_id, const float* h_kernel, int kvol, int cin, int cout);
int mt_tcn_set_bn(mt_tcn* t, int bn_id, const float* h_weight, const float* h_bias, const float* h_mean, const float* h_var,
                  int c, float eps);
int mt_tcn_set_gem(mt_tcn* t, float p, float eps);
/* d_keys: n 64-bit coordinates of the quantised clouds of `batch` frames, batch-major
 * (batch << 54 | (x + 2^17) << 36 | (y + 2^17) << 18 | (z + 2^17); ME.utils.sparse_quantize +
 * batched_coordinates, tcn.py:124-131; duplicates are merged).  d_out: (batch, feature) float64
 * descriptors, L2-normalised when normalize != 0 (tcn.py:138-148).  d_counts (nullable): active
 * points per level (4 ints). */
int mt_tcn_forward(mt_tcn* t, const unsigned long long* d_keys, int n, int batch, int normalize, double* d_out,
                   int* d_counts, void* stream);
/* The same from the sampled, scaled clouds of tcn.py:96-123: d_clouds (batch, points, 3) float32;
 * voxel = floor(x * inv_q) with inv_q = the float32 reciprocal torch multiplies by when it evaluates
 * `cloud / q` on CUDA (tcn.py:124-130; float32(1 / q) in torch 2.x) -- duplicate voxels merged in the library (no
 * sort, no host synchronisation; rows = voxels in order of their first point).  A voxel coordinate
 * beyond +-131071 (or NaN) turns the descriptors of the call into NaN. */
int mt_tcn_embed(mt_tcn* t, const float* d_clouds, int batch, int points, float inv_q, int normalize, double* d_out,
                 int* d_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif
