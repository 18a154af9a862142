// Drift test of remove_invalid_particles (particle_filter.py:379-403): a particle is invalid
// when the float64 Euclidean distance from its translation to the nearest vertex of the
// down-sampled mesh (sklearn KDTree over vertices[::10], particle_filter.py:108-110) exceeds
// invalid_dist.  Only the boolean is needed, so the k-d tree becomes a uniform grid over the
// vertices (cell >= the default invalid_dist): scan the cells overlapping the +-dist box and
// stop at the first vertex within range.  Arithmetic follows sklearn's EuclideanDistance:
// sequential float64 sum of squared differences, sqrt, compare -- no FMA contraction.
#pragma once
#include "mt_math.cuh"

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
  const double dx = __dsub_rn(x, __ldg(T.verts + 3 * (size_t)p));
  const double dy = __dsub_rn(y, __ldg(T.verts + 3 * (size_t)p + 1));
  const double dz = __dsub_rn(z, __ldg(T.verts + 3 * (size_t)p + 2));
  double d2 = __dmul_rn(dx, dx);
  d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
  d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
  return !(sqrt(d2) > dist);
}

// true when some vertex lies within `dist` of (x,y,z) (i.e. the particle has NOT drifted).
// Vertices are filtered in float32 (squared distance against dist^2 with a relative band
// that covers the rounding of the float32 copies); only vertices inside the band take the
// exact float64 test, so the answer is the float64 one.  NaN coordinates -> false.
__device__ __forceinline__ bool mesh_within(const MeshTables& T, float xf, float yf, float zf, double dist) {
  if (!(xf == xf) || !(yf == yf) || !(zf == zf)) return false;
  if (T.vox.cls && dist == T.vox.dist) {
    const float fx = (xf - T.vox.org[0]) * T.vox.inv_v, fy = (yf - T.vox.org[1]) * T.vox.inv_v, fz = (zf - T.vox.org[2]) * T.vox.inv_v;
    if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)T.vox.dims[0] && fy < (float)T.vox.dims[1] && fz < (float)T.vox.dims[2]))
      return false;  // the voxel grid covers the vertices' bounding box inflated by more than dist
    const int k = __ldg(T.vox.cls + ((size_t)(int)fz * T.vox.dims[1] + (int)fy) * T.vox.dims[0] + (int)fx);
    if (k < 0) return k == MT_VOX_IN;
    if (mesh_vertex_within(T, k, (double)xf, (double)yf, (double)zf, dist)) return true;
  }
  const double x = (double)xf, y = (double)yf, z = (double)zf;
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
  // rows nearest to the particle first: the common answer (on the surface) is found early
  const int cy0 = min(max(mt_mesh_cellf(yf, g.orgf[1], g.inv_cellf, g.dims[1]), ylo), yhi);
  const int cz0 = min(max(mt_mesh_cellf(zf, g.orgf[2], g.inv_cellf, g.dims[2]), zlo), zhi);
  const int kzn = 2 * max(cz0 - zlo, zhi - cz0), kyn = 2 * max(cy0 - ylo, yhi - cy0);
  for (int kz = 0; kz <= kzn; ++kz) {
    const int cz = cz0 + ((kz & 1) ? -((kz + 1) >> 1) : (kz >> 1));  // cz0, cz0-1, cz0+1, ...
    if (cz < zlo || cz > zhi) continue;
    for (int ky = 0; ky <= kyn; ++ky) {
      const int cy = cy0 + ((ky & 1) ? -((ky + 1) >> 1) : (ky >> 1));
      if (cy < ylo || cy > yhi) continue;
      const int rb = (cz * g.dims[1] + cy) * g.dims[0];
      const int s = __ldg(T.cell_start + rb + xlo), e = __ldg(T.cell_start + rb + xhi + 1);
      for (int p = s; p < e; ++p) {
        const float4 v = __ldg(T.verts32 + p);
        const float dx = xf - v.x, dy = yf - v.y, dz = zf - v.z;
        const float q2 = dx * dx + dy * dy + dz * dz;
        if (q2 < lo2) return true;
        if (q2 <= hi2 && mesh_vertex_within(T, p, x, y, z, dist)) return true;
      }
    }
  }
  return false;
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
