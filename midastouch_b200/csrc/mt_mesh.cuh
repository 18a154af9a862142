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
  int dims[3];
};

MT_HD int mt_mesh_cell(double x, double org, double inv_cell, int dim) {
  double f = floor((x - org) * inv_cell);
  return (f < 0.0) ? 0 : (f >= (double)dim ? dim - 1 : (int)f);
}

#if defined(__CUDACC__)
struct MeshTables {
  const double* verts;    // V x 3, sorted by cell
  const int* cell_start;  // ncells + 1
  MeshGrid g;
  int V;
};

// true when some vertex lies within `dist` of (x,y,z) (i.e. the particle is NOT drifted).
// NaN coordinates -> false (the reference's `dist > invalid_dist` is False for NaN distances
// only if sklearn returned NaN; sklearn raises on NaN input, the engine treats such poses as
// invalid already through check_quats).
__device__ __forceinline__ bool mesh_within(const MeshTables& T, float xf, float yf, float zf, double dist) {
  const double x = (double)xf, y = (double)yf, z = (double)zf;
  if (!(x == x) || !(y == y) || !(z == z)) return false;
  const MeshGrid& g = T.g;
  const double r = dist * (1.0 + 1e-9) + 1e-300;
  const int xlo = mt_mesh_cell(x - r, g.org[0], g.inv_cell, g.dims[0]), xhi = mt_mesh_cell(x + r, g.org[0], g.inv_cell, g.dims[0]);
  const int ylo = mt_mesh_cell(y - r, g.org[1], g.inv_cell, g.dims[1]), yhi = mt_mesh_cell(y + r, g.org[1], g.inv_cell, g.dims[1]);
  const int zlo = mt_mesh_cell(z - r, g.org[2], g.inv_cell, g.dims[2]), zhi = mt_mesh_cell(z + r, g.org[2], g.inv_cell, g.dims[2]);
  for (int cz = zlo; cz <= zhi; ++cz)
    for (int cy = ylo; cy <= yhi; ++cy) {
      const int rb = (cz * g.dims[1] + cy) * g.dims[0];
      const int s = __ldg(T.cell_start + rb + xlo), e = __ldg(T.cell_start + rb + xhi + 1);
      for (int p = s; p < e; ++p) {
        const double dx = __dsub_rn(x, __ldg(T.verts + 3 * (size_t)p));
        const double dy = __dsub_rn(y, __ldg(T.verts + 3 * (size_t)p + 1));
        const double dz = __dsub_rn(z, __ldg(T.verts + 3 * (size_t)p + 2));
        double d2 = __dmul_rn(dx, dx);
        d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
        d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
        if (!(sqrt(d2) > dist)) return true;
      }
    }
  return false;
}
#endif
