// TEST INFRASTRUCTURE ONLY (oracle/) — CPU restatement of the PNP-Stokes element kernels of
// benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.ufl:9-74 (mixed space P1^3 x RT1 x DG0, 17 dofs per cell: 4 cation,
// 4 anion, 4 potential, 4 facet-normal velocity dofs, 1 pressure), written as plain quadrature loops over the weak
// forms. Pinned against the reference's own compiled kernels (oracle/_ref/libpnpref_ns.so, built by oracle/Makefile from
// vector_linear_pnp_ns_forms.h:16355-23898) and against golden vectors generated from them (tests/golden).
//
// Conventions taken from the generated code: local dof order (cat 0-3, an 4-7, phi 8-11, RT 12-15, DG0 16), element
// tensor A[17 * test + trial]; macro element of an interior facet = ('+' cell's 17, '-' cell's 17), A[34 * test + trial].
// RT1 basis of the UFC reference tetrahedron as tabulated in the header (FE1_C3 / _C4 / _C5 at :16537-16620; they are
// the affine functions phi^_i(X) = s_i (X - V_i) with s = (-1, +1, -1, +1), V_i = reference vertex i), mapped by the
// contravariant Piola transform: phi_i(x) = J phi^_i / detJ = s_i (x - v_i) / detJ (signed detJ), so div phi_i =
// 3 s_i / detJ and grad phi_i = (s_i / detJ) I.
// Quadrature: the cell rules FFC chose (a: degree 5, 15 points, :16414; L: degree 4, 14 points, :21776) are restated
// from their closed forms / moment equations (tools/tet_rules.py) because exp(cc) is not integrated exactly; the facet
// terms are quadratic polynomials, any rule exact for degree 2 gives the reference's numbers (:17302 uses 3 points).
#include <cmath>
#include <cstring>
#include "pnp_oracle.h"

namespace {

struct Rule { int n; double w[15]; double l[15][4]; };

void sym4(Rule& r, double a, double w)          // the four points (a, a, a, 1 - 3a)
{
  for (int j = 0; j < 4; ++j) { for (int i = 0; i < 4; ++i) r.l[r.n][i] = (i == j) ? 1.0 - 3.0 * a : a; r.w[r.n++] = w; }
}
void sym6(Rule& r, double a, double w)          // the six points (a, a, 1/2 - a, 1/2 - a)
{
  for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) {
    for (int k = 0; k < 4; ++k) r.l[r.n][k] = (k == i || k == j) ? a : 0.5 - a;
    r.w[r.n++] = w;
  }
}
const Rule& rule15()    // degree 5 (Keast), weights for the reference volume 1/6
{
  static Rule r; static bool init = false;
  if (!init) {
    r.n = 0;
    for (int i = 0; i < 4; ++i) r.l[0][i] = 0.25;
    r.w[r.n++] = 6544.0 / 36015.0 / 6.0;
    sym4(r, 1.0 / 3.0, 81.0 / 2240.0 / 6.0);
    sym4(r, 1.0 / 11.0, 161051.0 / 2304960.0 / 6.0);
    sym6(r, (1.0 - std::sqrt(7.0 / 13.0)) / 4.0, 338.0 / 5145.0 / 6.0);
    init = true;
  }
  return r;
}
const Rule& rule14()    // degree 4: 6 edge midpoints + two vertex orbits (constants: tools/tet_rules.py)
{
  static Rule r; static bool init = false;
  if (!init) {
    r.n = 0;
    sym6(r, 0.5, 1.0 / 315.0);
    sym4(r, 0.10052676522520447969, 0.014764970790496785072);
    sym4(r, 0.31437287349319219275, 0.02213979111426511969);
    init = true;
  }
  return r;
}

struct Tet {
  double v[4][3], g[4][3], detJ, det;
  explicit Tet(const double* x)
  {
    std::memcpy(v, x, sizeof v);
    double J[9], K[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[3 * i + j] = v[j + 1][i] - v[0][i];
    const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    detJ = J[0] * c00 + J[1] * c01 + J[2] * c02;
    det = std::fabs(detJ);
    K[0] = c00 / detJ; K[1] = (J[2] * J[7] - J[1] * J[8]) / detJ; K[2] = (J[1] * J[5] - J[2] * J[4]) / detJ;
    K[3] = c01 / detJ; K[4] = (J[0] * J[8] - J[2] * J[6]) / detJ; K[5] = (J[2] * J[3] - J[0] * J[5]) / detJ;
    K[6] = c02 / detJ; K[7] = (J[1] * J[6] - J[0] * J[7]) / detJ; K[8] = (J[0] * J[4] - J[1] * J[3]) / detJ;
    for (int d = 0; d < 3; ++d) {
      g[1][d] = K[d]; g[2][d] = K[3 + d]; g[3][d] = K[6 + d];
      g[0][d] = -(K[d] + K[3 + d] + K[6 + d]);
    }
  }
  void point(const double* lam, double* x) const
  {
    for (int d = 0; d < 3; ++d) x[d] = lam[0] * v[0][d] + lam[1] * v[1][d] + lam[2] * v[2][d] + lam[3] * v[3][d];
  }
  // RT1 basis function i at the physical point x
  void rt(int i, const double* x, double* phi) const
  {
    const double s = (i & 1) ? 1.0 : -1.0;
    for (int d = 0; d < 3; ++d) phi[d] = s * (x[d] - v[i][d]) / detJ;
  }
  double rt_gamma(int i) const { return ((i & 1) ? 1.0 : -1.0) / detJ; }    // grad phi_i = gamma I
  // h = 2 * circumradius, the formula the generated code uses (:17263-17281)
  double circumradius() const
  {
    auto len = [&](int a, int b) { double s = 0; for (int d = 0; d < 3; ++d) s += (v[a][d] - v[b][d]) * (v[a][d] - v[b][d]); return std::sqrt(s); };
    const double la = len(1, 2) * len(0, 3), lb = len(0, 2) * len(1, 3), lc = len(0, 1) * len(2, 3);
    const double s = 0.5 * (la + lb + lc);
    return std::sqrt(s * (s - la) * (s - lb) * (s - lc)) / (6.0 * (det / 6.0));
  }
};

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

}  // namespace

extern "C" {

// par8 = {permittivity, diffusivity0, valency0, diffusivity1, valency1, mu, penalty1, penalty2}
// bilinear form a, cell part (ufl :45-55 and :59-60)
void orc_ns_cell_a(double* A, const double* cc, const double* uu, const double* par, const double* xyz)
{
  const Tet T(xyz); const Rule& R = rule15();
  const double eps = par[0], D[2] = {par[1], par[3]}, q[2] = {par[2], par[4]}, mu = par[5];
  for (int i = 0; i < 289; ++i) A[i] = 0.0;
  double gc[3][3];                                   // gradients of the previous iterates cc[0..2]
  for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d) {
    gc[k][d] = 0.0;
    for (int a = 0; a < 4; ++a) gc[k][d] += cc[4 * k + a] * T.g[a][d];
  }
  for (int ip = 0; ip < R.n; ++ip) {
    const double* lam = R.l[ip]; const double w = R.w[ip] * T.det;
    double x[3]; T.point(lam, x);
    double phi[4][3], uv[3] = {0, 0, 0};
    for (int j = 0; j < 4; ++j) { T.rt(j, x, phi[j]); for (int d = 0; d < 3; ++d) uv[d] += uu[j] * phi[j][d]; }
    double E[2];
    for (int k = 0; k < 2; ++k) E[k] = std::exp(lam[0] * cc[4 * k] + lam[1] * cc[4 * k + 1] + lam[2] * cc[4 * k + 2] + lam[3] * cc[4 * k + 3]);
    for (int k = 0; k < 2; ++k) {                    // ion rows k (test function lambda_a of species k)
      double drift[3];
      for (int d = 0; d < 3; ++d) drift[d] = gc[k][d] + q[k] * gc[2][d];
      for (int a = 0; a < 4; ++a) {
        const int row = 4 * k + a;
        for (int b = 0; b < 4; ++b) {
          A[17 * row + 4 * k + b] += w * (D[k] * E[k] * (dot3(T.g[b], T.g[a]) + dot3(drift, T.g[a]) * lam[b])
                                          - E[k] * lam[b] * dot3(uv, T.g[a]));
          A[17 * row + 8 + b] += w * q[k] * D[k] * E[k] * dot3(T.g[b], T.g[a]);
        }
        for (int j = 0; j < 4; ++j) A[17 * row + 12 + j] -= w * E[k] * dot3(phi[j], T.g[a]);
      }
    }
    for (int a = 0; a < 4; ++a) {                    // potential rows
      const int row = 8 + a;
      for (int b = 0; b < 4; ++b) {
        A[17 * row + 8 + b] += w * eps * dot3(T.g[b], T.g[a]);
        for (int k = 0; k < 2; ++k) A[17 * row + 4 * k + b] -= w * q[k] * E[k] * lam[b] * lam[a];
      }
    }
    for (int i = 0; i < 4; ++i) {                    // velocity rows (test function phi_i)
      const int row = 12 + i;
      for (int j = 0; j < 4; ++j) A[17 * row + 12 + j] += w * 2.0 * mu * 3.0 * T.rt_gamma(i) * T.rt_gamma(j);
      A[17 * row + 16] -= w * 3.0 * T.rt_gamma(i);
      A[17 * 16 + 12 + i] += w * 3.0 * T.rt_gamma(i);
      for (int b = 0; b < 4; ++b) {
        for (int k = 0; k < 2; ++k) A[17 * row + 4 * k + b] += w * q[k] * lam[b] * E[k] * dot3(gc[2], phi[i]);
        A[17 * row + 8 + b] += w * (q[0] * E[0] + q[1] * E[1]) * dot3(T.g[b], phi[i]);
      }
    }
  }
}

// linear form L, cell part (ufl :63-70 and :74)
void orc_ns_cell_L(double* b17, const double* cc, const double* uu, const double* pp, const double* par, const double* xyz)
{
  const Tet T(xyz); const Rule& R = rule14();
  const double eps = par[0], D[2] = {par[1], par[3]}, q[2] = {par[2], par[4]}, mu = par[5];
  for (int i = 0; i < 17; ++i) b17[i] = 0.0;
  double gc[3][3];
  for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d) {
    gc[k][d] = 0.0;
    for (int a = 0; a < 4; ++a) gc[k][d] += cc[4 * k + a] * T.g[a][d];
  }
  double gam_uu = 0.0;                               // sym grad uu = gam_uu I, div uu = 3 gam_uu
  for (int j = 0; j < 4; ++j) gam_uu += uu[j] * T.rt_gamma(j);
  for (int ip = 0; ip < R.n; ++ip) {
    const double* lam = R.l[ip]; const double w = R.w[ip] * T.det;
    double x[3]; T.point(lam, x);
    double phi[4][3], uv[3] = {0, 0, 0};
    for (int j = 0; j < 4; ++j) { T.rt(j, x, phi[j]); for (int d = 0; d < 3; ++d) uv[d] += uu[j] * phi[j][d]; }
    double E[2];
    for (int k = 0; k < 2; ++k) E[k] = std::exp(lam[0] * cc[4 * k] + lam[1] * cc[4 * k + 1] + lam[2] * cc[4 * k + 2] + lam[3] * cc[4 * k + 3]);
    for (int k = 0; k < 2; ++k) {
      double drift[3];
      for (int d = 0; d < 3; ++d) drift[d] = gc[k][d] + q[k] * gc[2][d];
      for (int a = 0; a < 4; ++a) b17[4 * k + a] += w * (-D[k] * E[k] * dot3(drift, T.g[a]) + E[k] * dot3(uv, T.g[a]));
    }
    for (int a = 0; a < 4; ++a) b17[8 + a] += w * (-eps * dot3(gc[2], T.g[a]) + (q[0] * E[0] + q[1] * E[1]) * lam[a]);
    for (int i = 0; i < 4; ++i)
      b17[12 + i] += w * (-2.0 * mu * 3.0 * gam_uu * T.rt_gamma(i) + pp[0] * 3.0 * T.rt_gamma(i)
                          - (q[0] * E[0] + q[1] * E[1]) * dot3(gc[2], phi[i]));
    b17[16] -= w * 3.0 * gam_uu;
  }
}

// Interior-facet terms of a (ufl :56-58) when trial = 1, of L (ufl :71-73) when trial = 0 (then uu8 is the previous
// velocity on the macro element and out has 34 entries, sign of L included).
static void ns_facet(double* out, int bilinear, const double* uu8, const double* par,
                     const double* xyz0, const double* xyz1, int f0, int f1)
{
  const Tet T[2] = {Tet(xyz0), Tet(xyz1)};
  const double mu = par[5], p1 = par[6], p2 = par[7];
  static const int fv[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  const double *a = T[0].v[fv[f0][0]], *b = T[0].v[fv[f0][1]], *c = T[0].v[fv[f0][2]];
  double n[3] = {(b[1] - a[1]) * (c[2] - a[2]) - (b[2] - a[2]) * (c[1] - a[1]),
                 (b[2] - a[2]) * (c[0] - a[0]) - (b[0] - a[0]) * (c[2] - a[2]),
                 (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])};
  const double twice_area = std::sqrt(dot3(n, n));
  double out_dir = 0.0;                              // n must point out of the '+' cell: away from its vertex f0
  for (int d = 0; d < 3; ++d) out_dir += n[d] * (a[d] - T[0].v[f0][d]);
  const double sgn = out_dir > 0 ? 1.0 : -1.0;
  for (int d = 0; d < 3; ++d) n[d] *= sgn / twice_area;          // n('+'); n('-') = -n
  const double h_avg = 0.5 * (2.0 * T[0].circumradius() + 2.0 * T[1].circumradius());
  (void)f1;
  const int nout = bilinear ? 34 * 34 : 34;
  for (int i = 0; i < nout; ++i) out[i] = 0.0;
  static const double bl[3][3] = {{2.0 / 3, 1.0 / 6, 1.0 / 6}, {1.0 / 6, 2.0 / 3, 1.0 / 6}, {1.0 / 6, 1.0 / 6, 2.0 / 3}};
  for (int ip = 0; ip < 3; ++ip) {
    const double w = 0.5 * twice_area / 3.0;
    double x[3];
    for (int d = 0; d < 3; ++d) x[d] = bl[ip][0] * a[d] + bl[ip][1] * b[d] + bl[ip][2] * c[d];
    // macro basis function m = 4 * side + j: value on its own side, zero on the other
    double phi[8][3], gam[8], nflux[8];              // nflux = phi('+').n('+') + phi('-').n('-') contribution
    for (int s = 0; s < 2; ++s) for (int j = 0; j < 4; ++j) {
      T[s].rt(j, x, phi[4 * s + j]);
      gam[4 * s + j] = T[s].rt_gamma(j);
      nflux[4 * s + j] = (s == 0 ? 1.0 : -1.0) * dot3(phi[4 * s + j], n);
    }
    for (int ti = 0; ti < 8; ++ti) {                 // test function v
      const int row = 17 * (ti / 4) + 12 + (ti % 4);
      const double sv = ti < 4 ? 1.0 : -1.0;         // jump(v) = v('+') - v('-')
      if (bilinear) {
        for (int tj = 0; tj < 8; ++tj) {
          const int col = 17 * (tj / 4) + 12 + (tj % 4);
          const double su = tj < 4 ? 1.0 : -1.0;
          double v = 2.0 * mu * p1 * 0.5 * gam[tj] * nflux[ti];         // avg(sym grad u) : sym(v (x) n) jumps
          v += 2.0 * mu * p1 * nflux[tj] * 0.5 * gam[ti];
          v += 2.0 * mu * (p2 / h_avg) * su * sv * dot3(phi[tj], phi[ti]);
          out[34 * row + col] += w * v;
        }
      } else {
        double gu = 0.0, nu = 0.0, ju[3] = {0, 0, 0};
        for (int tj = 0; tj < 8; ++tj) {
          gu += 0.5 * uu8[tj] * gam[tj]; nu += uu8[tj] * nflux[tj];
          for (int d = 0; d < 3; ++d) ju[d] += (tj < 4 ? 1.0 : -1.0) * uu8[tj] * phi[tj][d];
        }
        out[row] -= w * (2.0 * mu * p1 * gu * nflux[ti] + 2.0 * mu * p1 * nu * 0.5 * gam[ti]
                         + 2.0 * mu * (p2 / h_avg) * sv * dot3(ju, phi[ti]));
      }
    }
  }
}

void orc_ns_facet_a(double* A1156, const double* par, const double* xyz0, const double* xyz1, int f0, int f1)
{ ns_facet(A1156, 1, nullptr, par, xyz0, xyz1, f0, f1); }

void orc_ns_facet_L(double* b34, const double* uu8, const double* par, const double* xyz0, const double* xyz1, int f0, int f1)
{ ns_facet(b34, 0, uu8, par, xyz0, xyz1, f0, f1); }

}  // extern "C"

// ---- global assembly of the mixed system as COO triplets (dolfin::assemble restated: cell loop, then interior-facet
// loop; src/pde.cpp:542-555). Dof numbering 3 v + k | 3 nv + facet | 3 nv + nf + cell. cells = sorted vertex ids,
// cell_facets[4 c + i] = facet opposite vertex i, facet_cells[2 f + {0, 1}] ('+' first, -1 = boundary).
// ref_lib != NULL: drive the REFERENCE's compiled kernels (oracle/_ref/libpnpref_ns.so) instead of the restatement.
// coo arrays hold 289 nc + 1156 n_interior entries (zeros included); returns the number written, -1 if ref_lib fails.
#include <dlfcn.h>
extern "C" long orc_ns_assemble(int nv, const double* xyz, int nc, const int* cells, const int* cell_facets, int nf,
                                const int* facet_cells, const double* cc, const double* uu, const double* pp,
                                const double* par8, const char* ref_lib, int* coo_row, int* coo_col, double* coo_val,
                                double* rhs)
{
  typedef void (*cell_a_fn)(double*, const double*, const double*, const double*, const double*);
  typedef void (*cell_L_fn)(double*, const double*, const double*, const double*, const double*, const double*);
  typedef void (*ref_facet_a_fn)(double*, const double*, const double*, const double*, const double*, const double*, int, int);
  typedef void (*ref_facet_L_fn)(double*, const double*, const double*, const double*, const double*, const double*, const double*, int, int);
  cell_a_fn cell_a = orc_ns_cell_a; cell_L_fn cell_L = orc_ns_cell_L;
  ref_facet_a_fn rfa = nullptr; ref_facet_L_fn rfL = nullptr;
  if (ref_lib) {
    void* h = dlopen(ref_lib, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    cell_a = (cell_a_fn)dlsym(h, "ref_ns_cell_a"); cell_L = (cell_L_fn)dlsym(h, "ref_ns_cell_L");
    rfa = (ref_facet_a_fn)dlsym(h, "ref_ns_facet_a"); rfL = (ref_facet_L_fn)dlsym(h, "ref_ns_facet_L");
    if (!cell_a || !cell_L || !rfa || !rfL) return -1;
  }
  const int ndof = 3 * nv + nf + nc;
  for (int i = 0; i < ndof; ++i) rhs[i] = 0.0;
  long n = 0;
  auto gather = [&](int c, int* dof, double* x12, double* cc12, double* uu4) {
    for (int a = 0; a < 4; ++a) {
      const int v = cells[4 * c + a];
      for (int d = 0; d < 3; ++d) x12[3 * a + d] = xyz[3 * v + d];
      for (int k = 0; k < 3; ++k) { dof[4 * k + a] = 3 * v + k; cc12[4 * k + a] = cc[3 * v + k]; }
      dof[12 + a] = 3 * nv + cell_facets[4 * c + a];
      uu4[a] = uu[cell_facets[4 * c + a]];
    }
    dof[16] = 3 * nv + nf + c;
  };
  for (int c = 0; c < nc; ++c) {
    int dof[17]; double x[12], c12[12], u4[4], A[289], b[17];
    gather(c, dof, x, c12, u4);
    cell_a(A, c12, u4, par8, x);
    cell_L(b, c12, u4, pp + c, par8, x);
    for (int i = 0; i < 17; ++i) {
      rhs[dof[i]] += b[i];
      for (int j = 0; j < 17; ++j) { coo_row[n] = dof[i]; coo_col[n] = dof[j]; coo_val[n++] = A[17 * i + j]; }
    }
  }
  for (int f = 0; f < nf; ++f) {
    const int c0 = facet_cells[2 * f], c1 = facet_cells[2 * f + 1];
    if (c1 < 0) continue;
    int dof[34]; double x0[12], x1[12], c24[24], u8[8], p2[2] = {pp[c0], pp[c1]}, A[1156], b[34];
    gather(c0, dof, x0, c24, u8); gather(c1, dof + 17, x1, c24 + 12, u8 + 4);
    int f0 = 0, f1 = 0;
    for (int j = 0; j < 4; ++j) { if (cell_facets[4 * c0 + j] == f) f0 = j; if (cell_facets[4 * c1 + j] == f) f1 = j; }
    if (rfa) { rfa(A, c24, u8, par8, x0, x1, f0, f1); rfL(b, c24, u8, p2, par8, x0, x1, f0, f1); }
    else { orc_ns_facet_a(A, par8, x0, x1, f0, f1); orc_ns_facet_L(b, u8, par8, x0, x1, f0, f1); }
    for (int i = 0; i < 34; ++i) {
      rhs[dof[i]] += b[i];
      for (int j = 0; j < 34; ++j) { coo_row[n] = dof[i]; coo_col[n] = dof[j]; coo_val[n++] = A[34 * i + j]; }
    }
  }
  return n;
}
