// TEST INFRASTRUCTURE ONLY: compiles midastouch_b200/csrc/mt_math.cuh (the per-particle
// arithmetic shared by the CUDA kernels) for the host with g++ so the no-GPU test tier can
// check it against the oracle.  Never loaded by the product.
#include "../midastouch_b200/csrc/mt_math.cuh"
#include "../midastouch_b200/csrc/mt_nn.cuh"
#include "../midastouch_b200/csrc/mt_cluster.cuh"

extern "C" {
void h_se3_keys(const float* aos, long long n, float* keys) {
  for (long long i = 0; i < n; ++i) {
    float P[3][4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P[r][c] = aos[16 * i + 4 * r + c];
    mt_se3_key(P, keys + 6 * i);
  }
}
void h_motion(const float* aos, long long n, const float* odom16, const float* tn, const float* rot, float* out) {
  float O[3][4];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) O[r][c] = odom16[4 * r + c];
  for (long long i = 0; i < n; ++i) {
    float P[3][4], Tn[3][4], G[3][4], R[3][4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P[r][c] = aos[16 * i + 4 * r + c];
    mt_noise_affine(tn + 3 * i, rot + 3 * i, Tn);
    mt_compose(O, Tn, G);
    mt_compose(P, G, R);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) out[16 * i + 4 * r + c] = R[r][c];
    out[16 * i + 12] = out[16 * i + 13] = out[16 * i + 14] = 0.f;
    out[16 * i + 15] = 1.f;
  }
}
void h_key_dist(const float* keys, long long m, const float* q, float* out) {
  for (long long i = 0; i < m; ++i) out[i] = mt_key_dist(q, keys + 6 * i);
}
// systematic ancestors from an explicit CDF (sequential reference of kernel B's slot logic)
void h_ancestors_from_cdf(const double* C, long long n, float u, long long* anc) {
  double dN = (double)n, off = (double)(u / (float)n);
  long long prev = 0;
  for (long long i = 0; i < n; ++i) {
    long long cnt = mt_count_below(C[i], n, dN, off);
    for (long long s = prev; s < cnt; ++s) anc[s] = i;
    if (cnt > prev) prev = cnt;
  }
  for (long long s = prev; s < n; ++s) anc[s] = -1;
}
// mt_div_rn(k, dN, 1 / dN) for k = k0 .. k0 + cnt - 1
void h_div_rn(long long k0, long long cnt, long long n, double* out) {
  const double dN = (double)n, rN = 1.0 / dN;
  for (long long k = 0; k < cnt; ++k) out[k] = mt_div_rn((double)(k0 + k), dN, rN);
}
void h_locs(long long n, float u, double* out) {
  double dN = (double)n, off = (double)(u / (float)n);
  for (long long k = 0; k < n; ++k) out[k] = mt_loc(k, dN, off);
}
void h_normals(unsigned long long seed, unsigned long long step, long long n, float* out6) {
  for (long long i = 0; i < n; ++i) mt_motion_normals(seed, step, (uint64_t)i, out6 + 6 * i, out6 + 6 * i + 3);
}
void h_rot_err(const float* gt16, const float* aos, long long n, float* out) {
  float G[3][4];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) G[r][c] = gt16[4 * r + c];
  for (long long i = 0; i < n; ++i) {
    float P[3][4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P[r][c] = aos[16 * i + 4 * r + c];
    out[i] = mt_rot_err_deg(G, P);
  }
}
// hint-graph search (mt_nn.cuh) on the host: keys (M,6), nbr (M,K,8), queries (n,6), hints (n,)
// -> idx (n,), ok (n,) [1 = proven exact by the list, 0 = would go to the grid search]
void h_hint_scan(const float* keys, long long M, const float* nbr, int K, const float* q, long long n, const int* hint,
                 int* idx, int* ok, float* dist) {
  for (long long i = 0; i < n; ++i) {
    float bd;
    int bi;
    bool r = mt_hint_scan(q + 6 * i, keys + 6 * hint[i], hint[i], nbr + (size_t)hint[i] * K * 8, K, bd, bi);
    idx[i] = bi;
    ok[i] = r ? 1 : 0;
    dist[i] = bd;
  }
}
int h_cell_coord(float x, float org, float inv_h, int dim) { return mt_cell_coord(x, org, inv_h, dim); }
// slot ownership of one shard (the arithmetic of k_step_b for rank r of a sharded step): this
// shard holds weights w[0..n) whose CDF interval is [A, A + sum(w)) of the global total S over N
// slots.  anc_local[k] = local parent of the shard's k-th child; returns the number of children
// and *slot_base = global index of its first slot.
long long h_shard_children(const double* w, long long n, double A, double S, long long N, float u, long long* anc_local,
                           long long* slot_base) {
  const double dN = (double)N, off = (double)(u / (float)N);
  const long long base = mt_count_below(A / S, N, dN, off);
  long long prev = base;
  double run = 0.0;
  for (long long i = 0; i < n; ++i) {
    run += w[i];
    long long cnt = mt_count_below((A + run) / S, N, dN, off);
    for (long long s = prev; s < cnt; ++s) anc_local[s - base] = i;
    if (cnt > prev) prev = cnt;
  }
  *slot_base = base;
  return prev - base;
}
void h_so3_to_quat(const float* aos, long long n, float* out4) {
  for (long long i = 0; i < n; ++i) {
    float P[3][4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P[r][c] = aos[16 * i + 4 * r + c];
    mt_so3_to_quat(P, out4 + 4 * i);
  }
}
// search index of the fallback path (mt_nn.cuh): build on the host exactly as mt_codebook_upload does, then run
// the host restatement of the box-hierarchy search for n queries.  seeds: candidate index per query or -1.
int h_bvh_search(const float* keys, long long M, const float* q, long long n, const int* seeds, int* idx, int* visited,
                 int* dims3) {
  MtBvhHost B;
  if (!mt_bvh_build(keys, (int)M, B)) return -1;
  dims3[0] = B.bp.n_leaf, dims3[1] = B.bp.n_l1, dims3[2] = B.bp.n_l2;
  for (long long i = 0; i < n; ++i) {
    const int s = seeds ? seeds[i] : -1;
    const float sd = s >= 0 ? mt_key_dist(q + 6 * i, keys + 6 * (long long)s) : 0.f;
    idx[i] = mt_bvh_search_host(B, (int)M, q + 6 * i, sd, s, visited + i);
  }
  return 0;
}
// leaf order and boxes of the search index: order (M), leaf boxes (n_leaf x 12), level-1 boxes (n_l1 x 12)
int h_bvh_layout(const float* keys, long long M, int* order, float* leaf, float* l1) {
  MtBvhHost B;
  if (!mt_bvh_build(keys, (int)M, B)) return -1;
  for (long long m = 0; m < M; ++m) order[m] = B.order[m];
  for (size_t k = 0; k < B.leaf.size(); ++k) leaf[k] = B.leaf[k];
  for (size_t k = 0; k < B.l1.size(); ++k) l1[k] = B.l1[k];
  return 0;
}
// categorical draws from an inclusive CDF (mt_cdf_draw) with the uniforms the kernel would use
void h_cdf_draws(const double* C, long long n, double S, unsigned long long seed, unsigned long long stream_id, long long n_draws,
                 int* idx, double* u_out) {
  for (long long j = 0; j < n_draws; ++j) {
    const mt_u4 ctr = {(uint32_t)j, (uint32_t)((uint64_t)j >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
    const mt_u4 r = mt_philox(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = mt_u01_53(r.x, r.y);
    if (u_out) u_out[j] = u;
    idx[j] = (int)mt_cdf_draw(C, n, S, u);
  }
}
}
