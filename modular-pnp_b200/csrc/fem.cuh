// pnp-b200 — P1 tetrahedron geometry and the quadrature rules of the reference's generated
// forms, as device constants. Values are the 15-digit numbers FFC printed
// (benchmarks/pnp_exact_solution/vector_linear_pnp_forms.h:7941-7968 W24/FE0,
//  :8409-8428 W15/FE0); the first barycentric coordinate is formed as 1-x-y-z like FE0.
#pragma once
#include "common.cuh"

namespace pnpb200 {

#define PNP_PT(x, y, z) {1.0 - (x) - (y) - (z), (x), (y), (z)}

#define PNP_A0 0.214602871259152
#define PNP_A1 0.356191386222545
#define PNP_B0 0.0406739585346113
#define PNP_B1 0.877978124396166
#define PNP_C0 0.322337890142276
#define PNP_C1 0.0329863295731731
#define PNP_D0 0.0636610018750175
#define PNP_D1 0.269672331458316
#define PNP_D2 0.603005664791649

static __constant__ double c_W24[24] = {
  0.00665379170969465, 0.00665379170969465, 0.00665379170969465, 0.00665379170969465,
  0.00167953517588678, 0.00167953517588678, 0.00167953517588678, 0.00167953517588678,
  0.0092261969239424,  0.0092261969239424,  0.0092261969239424,  0.0092261969239424,
  0.00803571428571428, 0.00803571428571428, 0.00803571428571428, 0.00803571428571428,
  0.00803571428571428, 0.00803571428571428, 0.00803571428571428, 0.00803571428571428,
  0.00803571428571428, 0.00803571428571428, 0.00803571428571428, 0.00803571428571428};

static __constant__ double c_P24[24][4] = {
  PNP_PT(PNP_A1, PNP_A0, PNP_A0), PNP_PT(PNP_A0, PNP_A0, PNP_A0), PNP_PT(PNP_A0, PNP_A0, PNP_A1), PNP_PT(PNP_A0, PNP_A1, PNP_A0),
  PNP_PT(PNP_B1, PNP_B0, PNP_B0), PNP_PT(PNP_B0, PNP_B0, PNP_B0), PNP_PT(PNP_B0, PNP_B0, PNP_B1), PNP_PT(PNP_B0, PNP_B1, PNP_B0),
  PNP_PT(PNP_C1, PNP_C0, PNP_C0), PNP_PT(PNP_C0, PNP_C0, PNP_C0), PNP_PT(PNP_C0, PNP_C0, PNP_C1), PNP_PT(PNP_C0, PNP_C1, PNP_C0),
  PNP_PT(PNP_D1, PNP_D0, PNP_D0), PNP_PT(PNP_D0, PNP_D1, PNP_D0), PNP_PT(PNP_D0, PNP_D0, PNP_D1),
  PNP_PT(PNP_D2, PNP_D0, PNP_D0), PNP_PT(PNP_D0, PNP_D2, PNP_D0), PNP_PT(PNP_D0, PNP_D0, PNP_D2),
  PNP_PT(PNP_D0, PNP_D1, PNP_D2), PNP_PT(PNP_D1, PNP_D2, PNP_D0), PNP_PT(PNP_D2, PNP_D0, PNP_D1),
  PNP_PT(PNP_D0, PNP_D2, PNP_D1), PNP_PT(PNP_D1, PNP_D0, PNP_D2), PNP_PT(PNP_D2, PNP_D1, PNP_D0)};

#define PNP_T3 0.333333333333333
#define PNP_E0 0.0909090909090909
#define PNP_E1 0.727272727272727
#define PNP_F0 0.0665501535736643
#define PNP_F1 0.433449846426336

static __constant__ double c_W15[15] = {
  0.0302836780970892,
  0.00602678571428572, 0.00602678571428572, 0.00602678571428572, 0.00602678571428572,
  0.011645249086029, 0.011645249086029, 0.011645249086029, 0.011645249086029,
  0.0109491415613864, 0.0109491415613864, 0.0109491415613864,
  0.0109491415613864, 0.0109491415613864, 0.0109491415613864};

static __constant__ double c_P15[15][4] = {
  PNP_PT(0.25, 0.25, 0.25),
  PNP_PT(0.0, PNP_T3, PNP_T3), PNP_PT(PNP_T3, PNP_T3, PNP_T3), PNP_PT(PNP_T3, PNP_T3, 0.0), PNP_PT(PNP_T3, 0.0, PNP_T3),
  PNP_PT(PNP_E1, PNP_E0, PNP_E0), PNP_PT(PNP_E0, PNP_E0, PNP_E0), PNP_PT(PNP_E0, PNP_E0, PNP_E1), PNP_PT(PNP_E0, PNP_E1, PNP_E0),
  PNP_PT(PNP_F1, PNP_F0, PNP_F0), PNP_PT(PNP_F0, PNP_F1, PNP_F0), PNP_PT(PNP_F0, PNP_F0, PNP_F1),
  PNP_PT(PNP_F0, PNP_F1, PNP_F1), PNP_PT(PNP_F1, PNP_F0, PNP_F1), PNP_PT(PNP_F1, PNP_F1, PNP_F0)};

// gradients of the four barycentric functions and |det J| of an affine tetrahedron given its
// vertex coordinates x[4][3] (UFC conventions: J_ij = x_{j+1,i} - x_{0,i}, K = J^{-1})
__device__ __forceinline__ void tet_geom(const double x[4][3], double g[4][3], double& absdet)
{
  const double J0 = x[1][0] - x[0][0], J1 = x[2][0] - x[0][0], J2 = x[3][0] - x[0][0];
  const double J3 = x[1][1] - x[0][1], J4 = x[2][1] - x[0][1], J5 = x[3][1] - x[0][1];
  const double J6 = x[1][2] - x[0][2], J7 = x[2][2] - x[0][2], J8 = x[3][2] - x[0][2];
  const double c00 = J4 * J8 - J5 * J7, c01 = J5 * J6 - J3 * J8, c02 = J3 * J7 - J4 * J6;
  const double c10 = J2 * J7 - J1 * J8, c11 = J0 * J8 - J2 * J6, c12 = J1 * J6 - J0 * J7;
  const double c20 = J1 * J5 - J2 * J4, c21 = J2 * J3 - J0 * J5, c22 = J0 * J4 - J1 * J3;
  const double det = J0 * c00 + J3 * c10 + J6 * c20;
  g[1][0] = c00 / det; g[1][1] = c10 / det; g[1][2] = c20 / det;
  g[2][0] = c01 / det; g[2][1] = c11 / det; g[2][2] = c21 / det;
  g[3][0] = c02 / det; g[3][1] = c12 / det; g[3][2] = c22 / det;
#pragma unroll
  for (int d = 0; d < 3; ++d) g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
  absdet = fabs(det);
}

__device__ __forceinline__ double dot3(const double* a, const double* b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// index of (r,s), r <= s, in the packed upper triangle of a symmetric 4x4
__device__ __forceinline__ int sym4(int r, int s)
{
  const int a = r < s ? r : s, b = r < s ? s : r;
  return a * 4 - (a * (a - 1)) / 2 + (b - a);
}

// position of vertex v in a cell, -1 if absent
__device__ __forceinline__ int local_index(const int4& c, int v)
{
  return c.x == v ? 0 : c.y == v ? 1 : c.z == v ? 2 : c.w == v ? 3 : -1;
}

__device__ __forceinline__ void load_cell_xyz(const double* __restrict__ xyz, const int4& c, double x[4][3])
{
  const int v[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    x[r][0] = xyz[3 * (size_t)v[r]]; x[r][1] = xyz[3 * (size_t)v[r] + 1]; x[r][2] = xyz[3 * (size_t)v[r] + 2];
  }
}

} // namespace pnpb200
