// Drift test of remove_invalid_particles (particle_filter.py:379-403): a particle is invalid
// when the float64 Euclidean distance from its translation to the nearest vertex of the
// down-sampled mesh (sklearn KDTree over vertices[::10], particle_filter.py:108-110) exceeds
// invalid_dist.  Only the boolean is needed, so the k-d tree becomes a uniform grid over the
// vertices (cell >= the default invalid_dist): scan the cells overlapping the +-dist box and
// stop at the first vertex within range.  Arithmetic follows sklearn's EuclideanDistance:
// sequential float64 sum of squared differences, sqrt, compare -- no FMA contraction.
#pragma once
#include "mt_math.cuh"
#ifndef MT_VOX2_KEEP
#define MT_VOX2_KEEP 0
#endif
#ifndef MT_VOX2
#define MT_VOX2 1
#endif

struct MeshGrid {
  double org[3];
  double cell, inv_cell;
  double coord_max;  // max |coordinate| (sizes the float32 filter band)
  float orgf[3], inv_cellf;  // float32 copies for the (conservatively inflated) search box
  int dims[3];
};

MT_HD int mt_mesh_cellf(float x, float org, float inv_cell, int dim) {
  float f = floorf((x - org) * inv_cell);
  return (f < 0.f) ? 0 : (f >= (float)dim ? dim - 1 : (int)f);
}

MT_HD int mt_mesh_cell(double x, double org, double inv_cell, int dim) {
  double f = floor((x - org) * inv_cell);
  return (f < 0.0) ? 0 : (f >= (double)dim ? dim - 1 : (int)f);
}

#if defined(__CUDACC__)
// Voxel classification for the default invalid_dist (built at upload): -1 = every point of the
// voxel is within the distance of some vertex, -2 = no point of the voxel is, v >= 0 = undecided,
// v = the vertex nearest to the voxel centre (tested first; the full search only runs if that one
// is out of range).  Particles on the surface are answered by one 4-byte load.
#define MT_VOX_IN (-1)
#define MT_VOX_OUT (-2)
struct MeshVoxels {
  const int* cls;  // nullptr: not built
  const unsigned* cls2;  // the same classes packed 2 bits per voxel (0 = out, 1 = in, 2 = undecided): 1/16 of the bytes,
                         // small enough to stay in L2 -- what the particle sweep consults (nullptr: not built)
  float org[3], inv_v;
  int dims[3];
  double dist;               // the distance the classes were computed for
};

struct MeshTables {
  const double* verts;    // V x 3 float64, sorted by cell (exact test)
  const float4* verts32;  // V x (x,y,z,0) float32 copies (filter)
  const int* cell_start;  // ncells + 1
  MeshGrid g;
  MeshVoxels vox;
  int V;
};

// exact float64 test of one vertex, sklearn's arithmetic
__device__ __forceinline__ bool mesh_vertex_within(const MeshTables& T, int p, double x, double y, double z, double dist) {
  const double dx = __dsub_rn(x, mt_ldk(T.verts + 3 * (size_t)p));
  const double dy = __dsub_rn(y, mt_ldk(T.verts + 3 * (size_t)p + 1));
  const double dz = __dsub_rn(z, mt_ldk(T.verts + 3 * (size_t)p + 2));
  double d2 = __dmul_rn(dx, dx);
  d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
  d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
  return !(sqrt(d2) > dist);
}

// Voxel class of a point for the default distance: 1 = within, 0 = not within, 2 = undecided (needs
// mesh_within_search; *k_out = the vertex nearest to the voxel centre), 3 = no voxel table for this distance.
// NaN coordinates -> 0.
__device__ __forceinline__ int mesh_voxel_class(const MeshTables& T, float xf, float yf, float zf, double dist, int* k_out) {
  if (!(xf == xf) || !(yf == yf) || !(zf == zf)) return 0;
  if (!(T.vox.cls && dist == T.vox.dist)) return 3;
  const float fx = (xf - T.vox.org[0]) * T.vox.inv_v, fy = (yf - T.vox.org[1]) * T.vox.inv_v, fz = (zf - T.vox.org[2]) * T.vox.inv_v;
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)T.vox.dims[0] && fy < (float)T.vox.dims[1] && fz < (float)T.vox.dims[2]))
    return 0;  // the voxel grid covers the vertices' bounding box inflated by more than dist
  const int k = __ldg(T.vox.cls + ((size_t)(int)fz * T.vox.dims[1] + (int)fy) * T.vox.dims[0] + (int)fx);
  if (k < 0) return k == MT_VOX_IN;
  *k_out = k;
  return 2;
}

__device__ __forceinline__ int mesh_quick(const MeshTables& T, float xf, float yf, float zf, double dist, int k);
// Voxel class from the packed table (particle sweep): 1 = within, 0 = not within, 2 = undecided, 3 = no table.
__device__ __forceinline__ int mesh_voxel_class2(const MeshTables& T, float xf, float yf, float zf, double dist) {
  if (!(xf == xf) || !(yf == yf) || !(zf == zf)) return 0;
  if (!(T.vox.cls2 && dist == T.vox.dist)) return 3;
  const float fx = (xf - T.vox.org[0]) * T.vox.inv_v, fy = (yf - T.vox.org[1]) * T.vox.inv_v, fz = (zf - T.vox.org[2]) * T.vox.inv_v;
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)T.vox.dims[0] && fy < (float)T.vox.dims[1] && fz < (float)T.vox.dims[2]))
    return 0;
  const size_t idx = ((size_t)(int)fz * T.vox.dims[1] + (int)fy) * T.vox.dims[0] + (int)fx;
#if MT_VOX2_KEEP
  return (int)((mt_ldk(T.vox.cls2 + (idx >> 4)) >> (2 * (unsigned)(idx & 15))) & 3u);
#else
  return (int)((__ldg(T.vox.cls2 + (idx >> 4)) >> (2 * (unsigned)(idx & 15))) & 3u);
#endif
}
__global__ void __launch_bounds__(256) k_mesh_pack2(const int* __restrict__ cls, size_t total, unsigned* __restrict__ out) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 16 >= total) return;
  unsigned v = 0;
  for (int k = 0; k < 16; ++k) {
    const size_t i = w * 16 + k;
    const int c = i < total ? cls[i] : MT_VOX_OUT;
    v |= (c == MT_VOX_IN ? 1u : (c == MT_VOX_OUT ? 0u : 2u)) << (2 * k);
  }
  out[w] = v;
}

// the search behind an undecided voxel: the remembered vertex first (k >= 0), then the uniform vertex grid.
// Vertices are filtered in float32 (squared distance against dist^2 with a relative band that covers the
// rounding of the float32 copies); only vertices inside the band take the exact float64 test, so the answer
// is the float64 one.
__device__ __noinline__ bool mesh_within_search(const MeshTables& T, float xf, float yf, float zf, double dist, int k) {
  const double x = (double)xf, y = (double)yf, z = (double)zf;
  if (k >= 0) {
    const int q = mesh_quick(T, xf, yf, zf, dist, k);
    if (q < 2) return q == 1;
  }
  const MeshGrid& g = T.g;
  // search box in float32, radius inflated by 1 % (>> the 1e-4-cell rounding of the float32 cell
  // coordinates) so that it covers every cell holding a vertex within `dist`
  const float r = (float)dist * 1.01f + 1e-30f;
  const int xlo = mt_mesh_cellf(xf - r, g.orgf[0], g.inv_cellf, g.dims[0]), xhi = mt_mesh_cellf(xf + r, g.orgf[0], g.inv_cellf, g.dims[0]);
  const int ylo = mt_mesh_cellf(yf - r, g.orgf[1], g.inv_cellf, g.dims[1]), yhi = mt_mesh_cellf(yf + r, g.orgf[1], g.inv_cellf, g.dims[1]);
  const int zlo = mt_mesh_cellf(zf - r, g.orgf[2], g.inv_cellf, g.dims[2]), zhi = mt_mesh_cellf(zf + r, g.orgf[2], g.inv_cellf, g.dims[2]);
  const float d2 = (float)(dist * dist);
  const float band = 1e-3f + (float)(1e-6 * g.coord_max / dist);  // float32 copies are off by <= 6e-8 |coordinate|
  const float lo2 = d2 * (1.f - band), hi2 = d2 * (1.f + band) + 1e-30f;
  // The +-r box spans at most three cells per axis (cell >= dist), i.e. at most nine (z, y) rows, each one contiguous
  // vertex range.  All eighteen range ends are requested together, then the vertices stream through in batches of
  // eight independent loads: ~6 dependent memory round trips per search instead of ~36 (this search runs after an L2's
  // worth of particle data has streamed through the cache, so every round trip is a DRAM access).
  const int cy0 = min(max(mt_mesh_cellf(yf, g.orgf[1], g.inv_cellf, g.dims[1]), ylo), yhi);
  const int cz0 = min(max(mt_mesh_cellf(zf, g.orgf[2], g.inv_cellf, g.dims[2]), zlo), zhi);
  if (zhi - zlo <= 2 && yhi - ylo <= 2) {
    int rs[9], re[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {  // rows nearest to the particle first: the common answer is found early
      const int oz = r / 3, oy = r % 3;
      const int cz = cz0 + (oz == 0 ? 0 : (oz == 1 ? -1 : 1)), cy = cy0 + (oy == 0 ? 0 : (oy == 1 ? -1 : 1));
      const bool ok = cz >= zlo && cz <= zhi && cy >= ylo && cy <= yhi;
      const int rb = ok ? (cz * g.dims[1] + cy) * g.dims[0] : 0;
      rs[r] = ok ? mt_ldk(T.cell_start + rb + xlo) : 0;
      re[r] = ok ? mt_ldk(T.cell_start + rb + xhi + 1) : 0;
    }
    // when cz0 (cy0) sits at the low or high end of its range, the offsets -1 / +1 miss the cell two steps away
    bool covered = (cz0 - zlo <= 1) && (zhi - cz0 <= 1) && (cy0 - ylo <= 1) && (yhi - cy0 <= 1);
    if (covered) {
#pragma unroll 1
      for (int r = 0; r < 9; ++r) {
        const int sb = rs[r], e = re[r];
        for (int p = sb; p < e; p += 8) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = mt_ldk(T.verts32 + min(p + u, e - 1));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float dx = xf - v[u].x, dy = yf - v[u].y, dz = zf - v[u].z;
            const float q2 = dx * dx + dy * dy + dz * dz;
            if (q2 < lo2) return true;
            if (q2 <= hi2 && mesh_vertex_within(T, min(p + u, e - 1), x, y, z, dist)) return true;
          }
        }
      }
      return false;
    }
  }
  // general form (cells smaller than the search radius, or the particle's own cell at the edge of the box)
  const int kzn = 2 * max(cz0 - zlo, zhi - cz0), kyn = 2 * max(cy0 - ylo, yhi - cy0);
  for (int kz = 0; kz <= kzn; ++kz) {
    const int cz = cz0 + ((kz & 1) ? -((kz + 1) >> 1) : (kz >> 1));  // cz0, cz0-1, cz0+1, ...
    if (cz < zlo || cz > zhi) continue;
    for (int ky = 0; ky <= kyn; ++ky) {
      const int cy = cy0 + ((ky & 1) ? -((ky + 1) >> 1) : (ky >> 1));
      if (cy < ylo || cy > yhi) continue;
      const int rb = (cz * g.dims[1] + cy) * g.dims[0];
      const int s = mt_ldk(T.cell_start + rb + xlo), e = mt_ldk(T.cell_start + rb + xhi + 1);
      for (int p = s; p < e; p += 4) {  // four independent loads per trip (the last ones clamped: harmless repeats)
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = mt_ldk(T.verts32 + min(p + u, e - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float dx = xf - v[u].x, dy = yf - v[u].y, dz = zf - v[u].z;
          const float q2 = dx * dx + dy * dy + dz * dz;
          if (q2 < lo2) return true;
          if (q2 <= hi2 && mesh_vertex_within(T, min(p + u, e - 1), x, y, z, dist)) return true;
        }
      }
    }
  }
  return false;
}

// true when some vertex lies within `dist` of (x,y,z) (i.e. the particle has NOT drifted).
__device__ __forceinline__ bool mesh_within(const MeshTables& T, float xf, float yf, float zf, double dist) {
  int k = -1;
  const int c = mesh_voxel_class(T, xf, yf, zf, dist, &k);
  if (c < 2) return c == 1;
  return mesh_within_search(T, xf, yf, zf, dist, c == 2 ? k : -1);
}

// The quick certificates of an undecided voxel (k = the vertex nearest to the centre c of the particle's voxel, at
// distance dmin from it):  1 = within (the particle is within `dist` of vertex k itself), 0 = not within (every
// vertex w has |x - w| >= |c - w| - |x - c| >= dmin - |x - c| > dist), 2 = still undecided: grid search.
__device__ __forceinline__ int mesh_quick(const MeshTables& T, float xf, float yf, float zf, double dist, int k) {
  const double x = (double)xf, y = (double)yf, z = (double)zf;
  const double vx = mt_ldk(T.verts + 3 * (size_t)k), vy = mt_ldk(T.verts + 3 * (size_t)k + 1), vz = mt_ldk(T.verts + 3 * (size_t)k + 2);
  {
    const double dx = __dsub_rn(x, vx), dy = __dsub_rn(y, vy), dz = __dsub_rn(z, vz);
    double d2 = __dmul_rn(dx, dx);
    d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
    d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
    if (!(sqrt(d2) > dist)) return 1;  // sklearn's arithmetic, as mesh_vertex_within
  }
  const float fx = (xf - T.vox.org[0]) * T.vox.inv_v, fy = (yf - T.vox.org[1]) * T.vox.inv_v, fz = (zf - T.vox.org[2]) * T.vox.inv_v;
  const double v = (double)(1.0f / T.vox.inv_v);  // the edge k_mesh_classify used
  const double cx = (double)T.vox.org[0] + ((int)fx + 0.5) * v, cy = (double)T.vox.org[1] + ((int)fy + 0.5) * v,
               cz = (double)T.vox.org[2] + ((int)fz + 0.5) * v;
  const double dmin = sqrt((cx - vx) * (cx - vx) + (cy - vy) * (cy - vy) + (cz - vz) * (cz - vz));
  const double e = sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz));
  if (dmin - e > dist * (1.0 + 1e-9) + 1e-9) return 0;  // 1 nm of slack against the rounding of c, dmin, e
  return 2;
}

// Grid search by a whole warp for ONE point (arguments warp-uniform, all 32 lanes must call): the (z, y) rows of
// the +-r box are contiguous vertex ranges; the lanes fetch the range ends of up to 32 rows at once, then take one
// vertex each of the concatenated ranges, 32 per round trip.  Same float32 filter + exact float64 test as
// mesh_within_search, hence the same answer.
__device__ __forceinline__ bool mesh_search_warp(const MeshTables& T, float xf, float yf, float zf, double dist) {
  const int lane = threadIdx.x & 31;
  const double x = (double)xf, y = (double)yf, z = (double)zf;
  const MeshGrid& g = T.g;
  const float r = (float)dist * 1.01f + 1e-30f;
  const int xlo = mt_mesh_cellf(xf - r, g.orgf[0], g.inv_cellf, g.dims[0]), xhi = mt_mesh_cellf(xf + r, g.orgf[0], g.inv_cellf, g.dims[0]);
  const int ylo = mt_mesh_cellf(yf - r, g.orgf[1], g.inv_cellf, g.dims[1]), yhi = mt_mesh_cellf(yf + r, g.orgf[1], g.inv_cellf, g.dims[1]);
  const int zlo = mt_mesh_cellf(zf - r, g.orgf[2], g.inv_cellf, g.dims[2]), zhi = mt_mesh_cellf(zf + r, g.orgf[2], g.inv_cellf, g.dims[2]);
  const float d2 = (float)(dist * dist);
  const float band = 1e-3f + (float)(1e-6 * g.coord_max / dist);
  const float lo2 = d2 * (1.f - band), hi2 = d2 * (1.f + band) + 1e-30f;
  const int ny = yhi - ylo + 1, nrows = ny * (zhi - zlo + 1);
  for (int r0 = 0; r0 < nrows; r0 += 32) {
    const int row = r0 + lane;
    int s = 0, cnt = 0;
    if (row < nrows) {
      const int rb = ((zlo + row / ny) * g.dims[1] + (ylo + row % ny)) * g.dims[0];
      s = mt_ldk(T.cell_start + rb + xlo);
      cnt = mt_ldk(T.cell_start + rb + xhi + 1) - s;
    }
    int total = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const int nr = min(32, nrows - r0);
    for (int t0 = 0; t0 < total; t0 += 32) {
      int t = t0 + lane, p = -1;
      for (int rr = 0; rr < nr; ++rr) {  // which row holds flattened index t
        const int c = __shfl_sync(0xffffffffu, cnt, rr), sr = __shfl_sync(0xffffffffu, s, rr);
        if (p < 0 && t < c) p = sr + t;
        t -= c;
      }
      bool hit = false;
      if (p >= 0 && t0 + lane < total) {
        const float4 v = mt_ldk(T.verts32 + p);
        const float dx = xf - v.x, dy = yf - v.y, dz = zf - v.z;
        const float q2 = dx * dx + dy * dy + dz * dz;
        hit = (q2 < lo2) || (q2 <= hi2 && mesh_vertex_within(T, p, x, y, z, dist));
      }
      if (__any_sync(0xffffffffu, hit)) return true;
    }
  }
  return false;
}

// mesh_within for a warp-uniform point, all 32 lanes calling
__device__ __forceinline__ bool mesh_within_warp(const MeshTables& T, float xf, float yf, float zf, double dist) {
  int k = -1;
  const int c = mesh_voxel_class(T, xf, yf, zf, dist, &k);
  if (c < 2) return c == 1;
  if (c == 2) {
    const int q = mesh_quick(T, xf, yf, zf, dist, k);
    if (q < 2) return q == 1;
  }
  return mesh_search_warp(T, xf, yf, zf, dist);
}

// upload-time kernel: classify every voxel (edge v) against distance `dist`.  hd = half diagonal of
// the voxel inflated by 1 % + an absolute slack: a particle that the float32 index arithmetic of
// mesh_within assigns to this voxel lies within hd of its centre.
__global__ void __launch_bounds__(256) k_mesh_classify(MeshTables T, MeshVoxels V, float v, float slack, int* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)V.dims[0] * V.dims[1] * V.dims[2];
  if (idx >= total) return;
  const int ix = (int)(idx % V.dims[0]), iy = (int)((idx / V.dims[0]) % V.dims[1]), iz = (int)(idx / ((size_t)V.dims[0] * V.dims[1]));
  const double cx = (double)V.org[0] + (ix + 0.5) * (double)v, cy = (double)V.org[1] + (iy + 0.5) * (double)v,
               cz = (double)V.org[2] + (iz + 0.5) * (double)v;
  const double hd = 0.8660254037844386 * (double)v * 1.01 + (double)slack;
  const double R = V.dist + hd;
  const MeshGrid& g = T.g;
  const int xlo = mt_mesh_cell(cx - R, g.org[0], g.inv_cell, g.dims[0]), xhi = mt_mesh_cell(cx + R, g.org[0], g.inv_cell, g.dims[0]);
  const int ylo = mt_mesh_cell(cy - R, g.org[1], g.inv_cell, g.dims[1]), yhi = mt_mesh_cell(cy + R, g.org[1], g.inv_cell, g.dims[1]);
  const int zlo = mt_mesh_cell(cz - R, g.org[2], g.inv_cell, g.dims[2]), zhi = mt_mesh_cell(cz + R, g.org[2], g.inv_cell, g.dims[2]);
  double dmin2 = 1e300;
  int amin = -1;
  for (int z = zlo; z <= zhi; ++z)
    for (int y = ylo; y <= yhi; ++y) {
      const int rb = (z * g.dims[1] + y) * g.dims[0];
      const int s = T.cell_start[rb + xlo], e = T.cell_start[rb + xhi + 1];
      for (int p = s; p < e; ++p) {
        const double dx = cx - T.verts[3 * (size_t)p], dy = cy - T.verts[3 * (size_t)p + 1], dz = cz - T.verts[3 * (size_t)p + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < dmin2) dmin2 = d2, amin = p;
      }
    }
  const double dmin = sqrt(dmin2);
  int k = amin;  // undecided: remember the nearest vertex
  if (amin < 0 || dmin - hd > V.dist * (1.0 + 1e-9)) k = MT_VOX_OUT;
  else if (dmin + hd <= V.dist * (1.0 - 1e-9)) k = MT_VOX_IN;
  out[idx] = k;
}
#endif
