// TEST INFRASTRUCTURE ONLY (oracle/): the two affine-tetrahedron helpers of UFC's
// ufc_geometry.h that the reference's generated tabulate_tensor bodies call
// (include/EAFE.h:4719-4726). Conventions [EXT, UFC 2016.1]: coordinate_dofs is
// vertex-major (x0 y0 z0 x1 y1 z1 ...); J[3*i+j] = d x_i / d X_j = x_{j+1,i} - x_{0,i};
// K = J^{-1} row-major; det = det J (signed).
#ifndef PNPB200_ORACLE_SHIM_UFC_GEOMETRY_H
#define PNPB200_ORACLE_SHIM_UFC_GEOMETRY_H
#include <cmath>
#include <cstddef>

inline void compute_jacobian_tetrahedron_3d(double* J, const double* coordinate_dofs)
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      J[3 * i + j] = coordinate_dofs[3 * (j + 1) + i] - coordinate_dofs[i];
}

inline void compute_jacobian_inverse_tetrahedron_3d(double* K, double& det, const double* J)
{
  const double d_00 = J[4] * J[8] - J[5] * J[7];
  const double d_01 = J[5] * J[6] - J[3] * J[8];
  const double d_02 = J[3] * J[7] - J[4] * J[6];
  const double d_10 = J[2] * J[7] - J[1] * J[8];
  const double d_11 = J[0] * J[8] - J[2] * J[6];
  const double d_12 = J[1] * J[6] - J[0] * J[7];
  const double d_20 = J[1] * J[5] - J[2] * J[4];
  const double d_21 = J[2] * J[3] - J[0] * J[5];
  const double d_22 = J[0] * J[4] - J[1] * J[3];
  det = J[0] * d_00 + J[3] * d_10 + J[6] * d_20;
  K[0] = d_00 / det; K[1] = d_10 / det; K[2] = d_20 / det;
  K[3] = d_01 / det; K[4] = d_11 / det; K[5] = d_21 / det;
  K[6] = d_02 / det; K[7] = d_12 / det; K[8] = d_22 / det;
}

// shortest / longest edge of facet `facet` (opposite local vertex `facet`); the interior-facet kernels of
// benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.h:17253-17259 call them (their forms do not use the values)
inline void facet_edge_lengths_tetrahedron_3d(double& lo, double& hi, std::size_t facet, const double* x)
{
  static const int fv[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  lo = 1e300; hi = 0.0;
  for (int e = 0; e < 3; ++e) {
    const int a = fv[facet][e], b = fv[facet][(e + 1) % 3];
    double l2 = 0.0;
    for (int d = 0; d < 3; ++d) l2 += (x[3 * a + d] - x[3 * b + d]) * (x[3 * a + d] - x[3 * b + d]);
    const double l = std::sqrt(l2);
    if (l < lo) lo = l;
    if (l > hi) hi = l;
  }
}
inline void compute_min_facet_edge_length_tetrahedron_3d(double& v, std::size_t facet, const double* x)
{ double hi; facet_edge_lengths_tetrahedron_3d(v, hi, facet, x); }
inline void compute_max_facet_edge_length_tetrahedron_3d(double& v, std::size_t facet, const double* x)
{ double lo; facet_edge_lengths_tetrahedron_3d(lo, v, facet, x); }
#endif
