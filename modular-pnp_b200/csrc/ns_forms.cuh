// pnp-b200 — PNP-Stokes Newton system: the linearised forms of
// benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl:45-74 on the mixed space P1^3 x RT1 x DG0, assembled ROW-WISE.
//
// Not the generated tabulate_tensor (vector_linear_pnp_ns_forms.h:16355-23898: 17 x 17 cell tensor by a 15-point
// quadrature loop, 34 x 34 macro tensor per interior facet). P1 gradients are constant per cell and the lowest-order
// RT basis is affine, phi_i(x) = gamma_i (x - v_i), gamma_i = s_i / detJ, s = (-1, +1, -1, +1) (the FIAT basis the
// generated tables hold, Piola-mapped) — so grad phi_i = gamma_i I, div phi_i = 3 gamma_i, and every entry of the
// cell tensor is a closed-form combination of the exponential moments of the two ion species,
//     S0 = int E,  S1[a] = int E lambda_a,  S2[a][b] = int E lambda_a lambda_b,   E = exp(cc_k),
// taken with the same quadrature rules as the reference (degree 5 / 15 points for a, degree 4 / 14 points for L —
// exp is not integrated exactly, so the rule is part of the result). One pass computes the 40 moments of a cell
// (ns_cell_moments), a second pass builds every matrix row from the moments of its cells (ns_assemble_row): one thread
// owns one row, adds its contributions in a fixed order and writes each entry once — no atomics, bit-reproducible.
// Interior-facet terms (DG penalty of the RT velocity, ufl :56-58, :71-73) only touch velocity rows; the normal trace of
// phi_i is gamma_i |detJ| / (2 |F|) on its own facet and zero on the other three, the tangential jump products are
// quadratic and integrated with the edge-midpoint rule (exact).
//
// Everything here is PNPB200_HD (host + device) so that tests/host_shims can run the very same row code on the CPU
// against the reference's compiled kernels; the product only ever calls it from kernels (assemble_ns.cu).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PNPB200_HD __host__ __device__ __forceinline__
#else
#define PNPB200_HD inline
#endif

namespace pnpb200 {

constexpr int NS_MOM = 40;        // per cell: species k: 15-point S0, S1[4], S2[10]; 14-point T0, T1[4]  (k = 0, 1)

// read-only view of the mesh + dofmap + state, and the CSR system being written
struct NsView {
  int nv, nc, nf, ndof;
  const double* xyz;              // [3 nv]
  const int* cells;               // [4 nc], vertices ascending (UFC ordering)
  const int* cell_facets;         // [4 nc], facet i is opposite local vertex i
  const int* facet_cells;         // [2 nf], lower cell index first ('+'), -1 = boundary
  const int* vcell_ptr;           // vertex -> cells adjacency (CSR), cells ascending
  const int* vcell;
  const double* cc;               // previous iterate: [3 nv], dof 3 v + k (k: cation, anion, potential)
  const double* uu;               // [nf] facet-normal velocity dofs
  const double* pp;               // [nc]
  double par[8];                  // permittivity, diffusivity0, valency0, diffusivity1, valency1, mu, penalty1, penalty2
  double* mom;                    // [NS_MOM nc]
  const int* ia;                  // mixed CSR: dofs 3 v + k | 3 nv + f | 3 nv + nf + c
  const int* ja;
  double* val;
  double* rhs;
  const unsigned char* bc;        // [ndof] homogeneous Dirichlet flags (may be null)
};

struct NsTet {
  double v[4][3], g[4][3], gam[4], detJ, det;
};

PNPB200_HD void ns_tet(const NsView& S, int c, NsTet& T, int* vid)
{
  for (int a = 0; a < 4; ++a) {
    vid[a] = S.cells[4 * c + a];
    for (int d = 0; d < 3; ++d) T.v[a][d] = S.xyz[3 * (size_t)vid[a] + d];
  }
  double J[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[3 * i + j] = T.v[j + 1][i] - T.v[0][i];
  const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  T.detJ = J[0] * c00 + J[1] * c01 + J[2] * c02;
  T.det = fabs(T.detJ);
  const double r = 1.0 / T.detJ;
  const double K[9] = {c00 * r, (J[2] * J[7] - J[1] * J[8]) * r, (J[1] * J[5] - J[2] * J[4]) * r,
                       c01 * r, (J[0] * J[8] - J[2] * J[6]) * r, (J[2] * J[3] - J[0] * J[5]) * r,
                       c02 * r, (J[1] * J[6] - J[0] * J[7]) * r, (J[0] * J[4] - J[1] * J[3]) * r};
  for (int d = 0; d < 3; ++d) {
    T.g[1][d] = K[d]; T.g[2][d] = K[3 + d]; T.g[3][d] = K[6 + d];
    T.g[0][d] = -(K[d] + K[3 + d] + K[6 + d]);
  }
  for (int i = 0; i < 4; ++i) T.gam[i] = ((i & 1) ? 1.0 : -1.0) * r;
}

PNPB200_HD double ns_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// h = 2 R, R = circumradius from the products of opposite edge lengths (Heron form, the formula of the generated
// code, vector_linear_pnp_ns_forms.h:17263-17281)
PNPB200_HD double ns_circumradius(const NsTet& T)
{
  double l[4][4];
  for (int a = 0; a < 4; ++a) for (int b = a + 1; b < 4; ++b) {
    double s = 0.0;
    for (int d = 0; d < 3; ++d) s += (T.v[a][d] - T.v[b][d]) * (T.v[a][d] - T.v[b][d]);
    l[a][b] = sqrt(s);
  }
  const double la = l[1][2] * l[0][3], lb = l[0][2] * l[1][3], lc = l[0][1] * l[2][3];
  const double s = 0.5 * (la + lb + lc);
  return sqrt(s * (s - la) * (s - lb) * (s - lc)) / T.det;
}

// index of the symmetric pair (a, b) in the 10 stored second moments
PNPB200_HD int ns_sym(int a, int b) { const int lo = a < b ? a : b, hi = a < b ? b : a; return lo * 4 - lo * (lo - 1) / 2 + (hi - lo); }

// point p of an orbit-structured rule: orbit 0 = centroid, 1 = (a, a, a, 1 - 3a) x 4, 2 = (a, a, 1/2 - a, 1/2 - a) x 6
PNPB200_HD void ns_orbit_point(int kind, double a, int p, double* lam)
{
  if (kind == 0) { lam[0] = lam[1] = lam[2] = lam[3] = 0.25; return; }
  if (kind == 1) { for (int i = 0; i < 4; ++i) lam[i] = (i == p) ? 1.0 - 3.0 * a : a; return; }
  int i = 0, j = 1, n = 0;
  for (int ii = 0; ii < 4; ++ii) for (int jj = ii + 1; jj < 4; ++jj) { if (n == p) { i = ii; j = jj; } ++n; }
  for (int k = 0; k < 4; ++k) lam[k] = (k == i || k == j) ? a : 0.5 - a;
}

// pass 1: the 40 exponential moments of cell c (scaled by |detJ|)
PNPB200_HD void ns_cell_moments(const NsView& S, int c)
{
  NsTet T; int vid[4];
  ns_tet(S, c, T, vid);
  double cn[2][4];
  for (int k = 0; k < 2; ++k) for (int a = 0; a < 4; ++a) cn[k][a] = S.cc[3 * (size_t)vid[a] + k];
  double m[NS_MOM];
  for (int i = 0; i < NS_MOM; ++i) m[i] = 0.0;
  // degree 5, 15 points (Keast): centroid, (1/3,1/3,1/3,0), (1/11,1/11,1/11,8/11), 6 points on the edge bisectors
  const int kind15[4] = {0, 1, 1, 2}, cnt15[4] = {1, 4, 4, 6};
  const double a15[4] = {0.25, 1.0 / 3.0, 1.0 / 11.0, 0.06655015357366429824};       // (1 - sqrt(7/13)) / 4
  const double w15[4] = {6544.0 / 36015.0 / 6.0, 81.0 / 2240.0 / 6.0, 161051.0 / 2304960.0 / 6.0, 338.0 / 5145.0 / 6.0};
  for (int o = 0; o < 4; ++o) for (int p = 0; p < cnt15[o]; ++p) {
    double lam[4]; ns_orbit_point(kind15[o], a15[o], p, lam);
    for (int k = 0; k < 2; ++k) {
      const double e = w15[o] * T.det * exp(lam[0] * cn[k][0] + lam[1] * cn[k][1] + lam[2] * cn[k][2] + lam[3] * cn[k][3]);
      double* mk = m + 20 * k;
      mk[0] += e;
      for (int a = 0; a < 4; ++a) mk[1 + a] += e * lam[a];
      int n = 5;
      for (int a = 0; a < 4; ++a) for (int b = a; b < 4; ++b) mk[n++] += e * lam[a] * lam[b];
    }
  }
  // degree 4, 14 points: the 6 edge midpoints and two vertex orbits (constants from the moment equations, tools/tet_rules.py)
  const int kind14[3] = {2, 1, 1}, cnt14[3] = {6, 4, 4};
  const double a14[3] = {0.5, 0.10052676522520447969, 0.31437287349319219275};
  const double w14[3] = {1.0 / 315.0, 0.014764970790496785072, 0.02213979111426511969};
  for (int o = 0; o < 3; ++o) for (int p = 0; p < cnt14[o]; ++p) {
    double lam[4]; ns_orbit_point(kind14[o], a14[o], p, lam);
    for (int k = 0; k < 2; ++k) {
      const double e = w14[o] * T.det * exp(lam[0] * cn[k][0] + lam[1] * cn[k][1] + lam[2] * cn[k][2] + lam[3] * cn[k][3]);
      double* mk = m + 20 * k + 15;
      mk[0] += e;
      for (int a = 0; a < 4; ++a) mk[1 + a] += e * lam[a];
    }
  }
  for (int i = 0; i < NS_MOM; ++i) S.mom[(size_t)NS_MOM * c + i] = m[i];
}

// adds v to the entry (row, col) of the CSR matrix; the row's columns are ascending
PNPB200_HD void ns_add(const NsView& S, int row, int col, double v)
{
  int lo = S.ia[row], hi = S.ia[row + 1] - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (S.ja[mid] < col) lo = mid + 1; else hi = mid; }
  S.val[lo] += v;             // the pattern holds every (row, col) the forms couple; the owner thread is the only writer
}

// The 17 local dofs of cell c in the mixed numbering: 3 v + k (k-major in the local order), 3 nv + facet, 3 nv + nf + c
PNPB200_HD int ns_local_dof(const NsView& S, int c, const int* vid, int l)
{
  if (l < 12) return 3 * vid[l & 3] + (l >> 2);
  if (l < 16) return 3 * S.nv + S.cell_facets[4 * c + (l - 12)];
  return 3 * S.nv + S.nf + c;
}

// Row `li` (local index 0..16) of the cell tensor of cell c and its rhs entry, from the cell's moments.
// row[17] in the local column order (cat 0-3, an 4-7, phi 8-11, RT 12-15, DG0 16).
PNPB200_HD void ns_cell_row(const NsView& S, int c, const NsTet& T, const int* vid, int li, double* row, double* rhs)
{
  const double eps = S.par[0], D[2] = {S.par[1], S.par[3]}, q[2] = {S.par[2], S.par[4]}, mu = S.par[5];
  const double* m = S.mom + (size_t)NS_MOM * c;
  for (int j = 0; j < 17; ++j) row[j] = 0.0;
  double gc[3][3];                                   // gradients of the previous iterates
  for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d)
    gc[k][d] = S.cc[3 * (size_t)vid[0] + k] * T.g[0][d] + S.cc[3 * (size_t)vid[1] + k] * T.g[1][d] +
               S.cc[3 * (size_t)vid[2] + k] * T.g[2][d] + S.cc[3 * (size_t)vid[3] + k] * T.g[3][d];
  // previous velocity uu(x) = gu x - cu
  double gu = 0.0, cu[3] = {0, 0, 0};
  for (int j = 0; j < 4; ++j) {
    const double u = S.uu[S.cell_facets[4 * c + j]] * T.gam[j];
    gu += u;
    for (int d = 0; d < 3; ++d) cu[d] += u * T.v[j][d];
  }
  const double vol = T.det / 6.0;
  if (li < 8) {                                      // ion row: species k, test function lambda_a
    const int k = li >> 2, a = li & 3;
    const double* mk = m + 20 * k;
    double drift[3];
    for (int d = 0; d < 3; ++d) drift[d] = gc[k][d] + q[k] * gc[2][d];
    const double da = ns_dot(drift, T.g[a]);
    double X0[3] = {0, 0, 0};                        // int E x
    for (int b = 0; b < 4; ++b) for (int d = 0; d < 3; ++d) X0[d] += mk[1 + b] * T.v[b][d];
    for (int b = 0; b < 4; ++b) {
      double X1[3] = {0, 0, 0};                      // int E lambda_b x
      for (int e = 0; e < 4; ++e) for (int d = 0; d < 3; ++d) X1[d] += mk[5 + ns_sym(e, b)] * T.v[e][d];
      double conv[3];
      for (int d = 0; d < 3; ++d) conv[d] = gu * X1[d] - mk[1 + b] * cu[d];
      const double gg = ns_dot(T.g[b], T.g[a]);
      row[4 * k + b] = D[k] * (mk[0] * gg + da * mk[1 + b]) - ns_dot(conv, T.g[a]);
      row[8 + b] = q[k] * D[k] * mk[0] * gg;
    }
    for (int j = 0; j < 4; ++j) {
      double f[3];
      for (int d = 0; d < 3; ++d) f[d] = X0[d] - mk[0] * T.v[j][d];
      row[12 + j] = -T.gam[j] * ns_dot(f, T.g[a]);
    }
    // rhs with the 14-point moments
    const double* tk = mk + 15;
    double Y0[3] = {0, 0, 0}, conv[3];
    for (int b = 0; b < 4; ++b) for (int d = 0; d < 3; ++d) Y0[d] += tk[1 + b] * T.v[b][d];
    for (int d = 0; d < 3; ++d) conv[d] = gu * Y0[d] - tk[0] * cu[d];
    *rhs = -D[k] * tk[0] * da + ns_dot(conv, T.g[a]);
  } else if (li < 12) {                              // potential row
    const int a = li - 8;
    double r = -eps * ns_dot(gc[2], T.g[a]) * vol;
    for (int b = 0; b < 4; ++b) {
      row[8 + b] = eps * ns_dot(T.g[b], T.g[a]) * vol;
      for (int k = 0; k < 2; ++k) row[4 * k + b] = -q[k] * m[20 * k + 5 + ns_sym(a, b)];
    }
    for (int k = 0; k < 2; ++k) r += q[k] * m[20 * k + 15 + 1 + a];
    *rhs = r;
  } else if (li < 16) {                              // velocity row: test function phi_i
    const int i = li - 12;
    for (int j = 0; j < 4; ++j) row[12 + j] = 6.0 * mu * T.gam[i] * T.gam[j] * vol;
    row[16] = -3.0 * T.gam[i] * vol;
    double F0[3] = {0, 0, 0}, G0[3] = {0, 0, 0};     // sum_k q_k int E_k (x - v_i): 15-point (matrix), 14-point (rhs)
    for (int k = 0; k < 2; ++k) {
      const double* mk = m + 20 * k;
      for (int d = 0; d < 3; ++d) {
        double x0 = 0.0, y0 = 0.0;
        for (int b = 0; b < 4; ++b) { x0 += mk[1 + b] * T.v[b][d]; y0 += mk[16 + b] * T.v[b][d]; }
        F0[d] += q[k] * (x0 - mk[0] * T.v[i][d]);
        G0[d] += q[k] * (y0 - mk[15] * T.v[i][d]);
      }
      for (int b = 0; b < 4; ++b) {
        double X1[3] = {0, 0, 0};
        for (int e = 0; e < 4; ++e) for (int d = 0; d < 3; ++d) X1[d] += mk[5 + ns_sym(e, b)] * T.v[e][d];
        double f[3];
        for (int d = 0; d < 3; ++d) f[d] = X1[d] - mk[1 + b] * T.v[i][d];
        row[4 * k + b] = q[k] * T.gam[i] * ns_dot(gc[2], f);
      }
    }
    for (int b = 0; b < 4; ++b) row[8 + b] = T.gam[i] * ns_dot(T.g[b], F0);
    *rhs = (-6.0 * mu * gu + 3.0 * S.pp[c]) * T.gam[i] * vol - T.gam[i] * ns_dot(gc[2], G0);
  } else {                                           // pressure row
    for (int j = 0; j < 4; ++j) row[12 + j] = 3.0 * T.gam[j] * vol;
    *rhs = -3.0 * gu * vol;
  }
}

// DG terms of interior facet F for the macro test function (side ts, local facet index ti): row8 = the 8 macro columns
// (side 0's four RT functions, then side 1's), rhs contribution returned.
PNPB200_HD double ns_facet_row(const NsView& S, int F, const NsTet* T, const int* cell, int ts, int ti, double* row8)
{
  const double mu = S.par[5], p1 = S.par[6], p2 = S.par[7];
  int lf[2] = {0, 0};                                // local index of F in each cell
  for (int s = 0; s < 2; ++s) for (int j = 0; j < 4; ++j) if (S.cell_facets[4 * cell[s] + j] == F) lf[s] = j;
  const double* fa[3]; int n = 0;
  for (int j = 0; j < 4; ++j) if (j != lf[0]) fa[n++] = T[0].v[j];
  double nn[3] = {(fa[1][1] - fa[0][1]) * (fa[2][2] - fa[0][2]) - (fa[1][2] - fa[0][2]) * (fa[2][1] - fa[0][1]),
                  (fa[1][2] - fa[0][2]) * (fa[2][0] - fa[0][0]) - (fa[1][0] - fa[0][0]) * (fa[2][2] - fa[0][2]),
                  (fa[1][0] - fa[0][0]) * (fa[2][1] - fa[0][1]) - (fa[1][1] - fa[0][1]) * (fa[2][0] - fa[0][0])};
  const double area = 0.5 * sqrt(ns_dot(nn, nn));
  const double h_avg = ns_circumradius(T[0]) + ns_circumradius(T[1]);      // (2 R+ + 2 R-) / 2
  // normal flux of a macro function through F (with its own cell's outward normal): only the function of F itself
  double gam[8], nflux[8], uu8[8];
  for (int s = 0; s < 2; ++s) for (int j = 0; j < 4; ++j) {
    gam[4 * s + j] = T[s].gam[j];
    nflux[4 * s + j] = (j == lf[s]) ? T[s].gam[j] * T[s].det / (2.0 * area) : 0.0;
    uu8[4 * s + j] = S.uu[S.cell_facets[4 * cell[s] + j]];
  }
  const int mt = 4 * ts + ti;
  const double sv = ts == 0 ? 1.0 : -1.0;
  const double* vt = T[ts].v[ti];
  double r = 0.0, gu = 0.0, nu = 0.0;
  for (int mc = 0; mc < 8; ++mc) {
    const double* vc = T[mc >> 2].v[mc & 3];
    // Q = int_F (x - vc).(x - vt) dS by the edge-midpoint rule
    double Q = 0.0;
    for (int e = 0; e < 3; ++e) {
      const double* pa = fa[e]; const double* pb = fa[(e + 1) % 3];
      double s = 0.0;
      for (int d = 0; d < 3; ++d) { const double x = 0.5 * (pa[d] + pb[d]); s += (x - vc[d]) * (x - vt[d]); }
      Q += s;
    }
    Q *= area / 3.0;
    const double su = mc < 4 ? 1.0 : -1.0;
    const double pen = 2.0 * mu * (p2 / h_avg) * su * sv * gam[mc] * gam[mt] * Q;
    row8[mc] = area * mu * p1 * (gam[mc] * nflux[mt] + nflux[mc] * gam[mt]) + pen;
    gu += 0.5 * uu8[mc] * gam[mc]; nu += uu8[mc] * nflux[mc];
    r += uu8[mc] * pen;
  }
  return -(area * 2.0 * mu * p1 * (gu * nflux[mt] + nu * 0.5 * gam[mt]) + r);
}

// pass 2: row r of the mixed system (matrix values zeroed, summed, Dirichlet rows replaced by identity rows with a
// zero right-hand side like dolfin::DirichletBC::apply on the Newton update, linear_pnp_ns.cpp:287-319)
PNPB200_HD void ns_assemble_row(const NsView& S, int r, bool matrix)
{
  if (matrix) for (int p = S.ia[r]; p < S.ia[r + 1]; ++p) S.val[p] = 0.0;
  if (S.bc && S.bc[r]) {
    if (matrix) ns_add(S, r, r, 1.0);
    S.rhs[r] = 0.0;
    return;
  }
  double b = 0.0, row[17];
  NsTet T[2]; int vid[4], cell[2];
  if (r < 3 * S.nv) {                                // P1 rows: the cells around the vertex
    const int v = r / 3, k = r - 3 * v;
    for (int p = S.vcell_ptr[v]; p < S.vcell_ptr[v + 1]; ++p) {
      const int c = S.vcell[p];
      ns_tet(S, c, T[0], vid);
      int a = 0;
      for (int j = 0; j < 4; ++j) if (vid[j] == v) a = j;
      double rb;
      ns_cell_row(S, c, T[0], vid, 4 * k + a, row, &rb);
      b += rb;
      if (matrix) for (int l = 0; l < 16; ++l) ns_add(S, r, ns_local_dof(S, c, vid, l), row[l]);   // no P1-pressure coupling
    }
  } else if (r < 3 * S.nv + S.nf) {                  // velocity rows: cell terms of the 1-2 cells, then the DG facet terms
    const int f = r - 3 * S.nv;
    for (int s = 0; s < 2; ++s) {
      const int c = S.facet_cells[2 * f + s];
      if (c < 0) continue;
      ns_tet(S, c, T[0], vid);
      int i = 0;
      for (int j = 0; j < 4; ++j) if (S.cell_facets[4 * c + j] == f) i = j;
      double rb;
      ns_cell_row(S, c, T[0], vid, 12 + i, row, &rb);
      b += rb;
      if (matrix) for (int l = 0; l < 17; ++l) ns_add(S, r, ns_local_dof(S, c, vid, l), row[l]);
      for (int j = 0; j < 4; ++j) {                  // macro elements of the interior facets of c that hold dof f
        const int F = S.cell_facets[4 * c + j];
        cell[0] = S.facet_cells[2 * F]; cell[1] = S.facet_cells[2 * F + 1];
        if (cell[1] < 0) continue;                   // boundary facet: no dS term
        if (F == f && c != cell[0]) continue;        // the facet's own macro element is taken from its '+' side only
        int vv[4];
        ns_tet(S, cell[0], T[0], vv); ns_tet(S, cell[1], T[1], vv);
        const int ts = (c == cell[0]) ? 0 : 1;
        double row8[8];
        if (F == f) {                                // f is a function of both sides of its own facet
          for (int t2 = 0; t2 < 2; ++t2) {
            int i2 = 0;
            for (int jj = 0; jj < 4; ++jj) if (S.cell_facets[4 * cell[t2] + jj] == f) i2 = jj;
            b += ns_facet_row(S, F, T, cell, t2, i2, row8);
            if (matrix) for (int mc = 0; mc < 8; ++mc) ns_add(S, r, 3 * S.nv + S.cell_facets[4 * cell[mc >> 2] + (mc & 3)], row8[mc]);
          }
        } else {
          b += ns_facet_row(S, F, T, cell, ts, i, row8);
          if (matrix) for (int mc = 0; mc < 8; ++mc) ns_add(S, r, 3 * S.nv + S.cell_facets[4 * cell[mc >> 2] + (mc & 3)], row8[mc]);
        }
      }
    }
  } else {                                           // pressure row of cell c
    const int c = r - 3 * S.nv - S.nf;
    ns_tet(S, c, T[0], vid);
    double rb;
    ns_cell_row(S, c, T[0], vid, 16, row, &rb);
    b = rb;
    if (matrix) for (int l = 12; l < 17; ++l) ns_add(S, r, ns_local_dof(S, c, vid, l), row[l]);
  }
  S.rhs[r] = b;
}

}  // namespace pnpb200
