// TEST INFRASTRUCTURE ONLY (oracle/) — see pnp_oracle.h.
// CPU restatement of the reference's three P1-tetrahedron element kernels, written from the
// weak forms (benchmarks/pnp_exact_solution/vector_linear_pnp_forms.ufl:27-43) and from the
// hand-edited EAFE scheme (include/EAFE.h:4851-4909), with the same quadrature rules the
// FFC-generated code tabulates (W24/FE0 forms.h:7941-7968, W15/FE0 forms.h:8409-8428).
// Validated against the reference's own compiled kernels in tests/test_oracle_elements.py.
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <cstdio>
#include "pnp_oracle.h"
#include "oracle_internal.hpp"

namespace orc {

// ---- quadrature tables -----------------------------------------------------------------
// 24-point degree-6 rule on the reference tetrahedron (values as printed by FFC with
// precision 15, forms.h:7941-7943): three 4-point orbits and one 12-point orbit.
static double W24[24], P24[24][4];
// 15-point degree-5 rule (forms.h:8409-8411): centroid, two 4-orbits, one 6-orbit.
static double W15[15], P15[15][4];
static bool tables_ready = false;
// form variant of benchmarks/pnp_diode/vector_linear_pnp_forms.h (:8606-8608, :8721-8724, L :I[9]):
// poisson_scale multiplies the charge terms of the Poisson row; with conc_cutoff the Jacobian's
// diffusion terms use exp(w) only for w > -50 (else 0). Defaults reproduce pnp_exact_solution.
static double g_poisson_scale = 1.0;
static int g_conc_cutoff = 0;
static void (*g_ref_set_scale)(double) = nullptr;

static void set_pt(double P[4], double x, double y, double z)
{
  // P1 basis on the UFC reference tetrahedron: (1-x-y-z, x, y, z)  (FE0 tables)
  P[0] = 1.0 - x - y - z; P[1] = x; P[2] = y; P[3] = z;
}

static void init_tables()
{
  if (tables_ready) return;
  // --- 24-point rule, point order of forms.h:7942 ---
  const double wa = 0.00665379170969465, a0 = 0.214602871259152, a1 = 0.356191386222545;
  const double wb = 0.00167953517588678, b0 = 0.0406739585346113, b1 = 0.877978124396166;
  const double wc = 0.0092261969239424,  c0 = 0.322337890142276, c1 = 0.0329863295731731;
  const double wd = 0.00803571428571428, d0 = 0.0636610018750175, d1 = 0.269672331458316,
               d2 = 0.603005664791649;
  int n = 0;
  const double orb4[3][3] = {{wa, a0, a1}, {wb, b0, b1}, {wc, c0, c1}};
  for (int o = 0; o < 3; ++o) {
    const double w = orb4[o][0], s = orb4[o][1], t = orb4[o][2];
    set_pt(P24[n], t, s, s); W24[n++] = w;
    set_pt(P24[n], s, s, s); W24[n++] = w;
    set_pt(P24[n], s, s, t); W24[n++] = w;
    set_pt(P24[n], s, t, s); W24[n++] = w;
  }
  const double o12[12][3] = {
    {d1, d0, d0}, {d0, d1, d0}, {d0, d0, d1}, {d2, d0, d0}, {d0, d2, d0}, {d0, d0, d2},
    {d0, d1, d2}, {d1, d2, d0}, {d2, d0, d1}, {d0, d2, d1}, {d1, d0, d2}, {d2, d1, d0}};
  for (int k = 0; k < 12; ++k) { set_pt(P24[n], o12[k][0], o12[k][1], o12[k][2]); W24[n++] = wd; }
  // --- 15-point rule, point order of forms.h:8410 ---
  const double w0 = 0.0302836780970892, w1 = 0.00602678571428572, w2 = 0.011645249086029,
               w3 = 0.0109491415613864;
  const double t3 = 0.333333333333333, e0 = 0.0909090909090909, e1 = 0.727272727272727;
  const double f0 = 0.0665501535736643, f1 = 0.433449846426336;
  n = 0;
  set_pt(P15[n], 0.25, 0.25, 0.25); W15[n++] = w0;
  set_pt(P15[n], 0.0, t3, t3); W15[n++] = w1;
  set_pt(P15[n], t3, t3, t3);  W15[n++] = w1;
  set_pt(P15[n], t3, t3, 0.0); W15[n++] = w1;
  set_pt(P15[n], t3, 0.0, t3); W15[n++] = w1;
  set_pt(P15[n], e1, e0, e0);  W15[n++] = w2;
  set_pt(P15[n], e0, e0, e0);  W15[n++] = w2;
  set_pt(P15[n], e0, e0, e1);  W15[n++] = w2;
  set_pt(P15[n], e0, e1, e0);  W15[n++] = w2;
  set_pt(P15[n], f1, f0, f0);  W15[n++] = w3;
  set_pt(P15[n], f0, f1, f0);  W15[n++] = w3;
  set_pt(P15[n], f0, f0, f1);  W15[n++] = w3;
  set_pt(P15[n], f0, f1, f1);  W15[n++] = w3;
  set_pt(P15[n], f1, f0, f1);  W15[n++] = w3;
  set_pt(P15[n], f1, f1, f0);  W15[n++] = w3;
  tables_ready = true;
}

// affine geometry: gradients of the four barycentric functions and |det J|
// (ufc_geometry conventions, see oracle/shim/ufc_geometry.h)
void tet_geometry(const double* x, double g[4][3], double* absdet)
{
  double J[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[3 * i + j] = x[3 * (j + 1) + i] - x[i];
  const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8],
               c02 = J[3] * J[7] - J[4] * J[6], c10 = J[2] * J[7] - J[1] * J[8],
               c11 = J[0] * J[8] - J[2] * J[6], c12 = J[1] * J[6] - J[0] * J[7],
               c20 = J[1] * J[5] - J[2] * J[4], c21 = J[2] * J[3] - J[0] * J[5],
               c22 = J[0] * J[4] - J[1] * J[3];
  const double det = J[0] * c00 + J[3] * c10 + J[6] * c20;
  // K = J^{-1}; grad lambda_{r+1} = row r of K
  g[1][0] = c00 / det; g[1][1] = c10 / det; g[1][2] = c20 / det;
  g[2][0] = c01 / det; g[2][1] = c11 / det; g[2][2] = c21 / det;
  g[3][0] = c02 / det; g[3][1] = c12 / det; g[3][2] = c22 / det;
  for (int d = 0; d < 3; ++d) g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
  *absdet = std::fabs(det);
}

static inline double dot3(const double* a, const double* b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// ---- EAFE element matrix (include/EAFE.h:4825-4909) --------------------------------------
void eafe_element(double* A, const double* eta, const double* alpha, const double* beta,
                  const double* gamma, const double* x)
{
  init_tables();
  double g[4][3], det;
  tet_geometry(x, g, &det);
  // P1 stiffness  S_rs = (det/6) grad l_r . grad l_s   (:4825-4849, factor 0.1666..67 as printed)
  double S[4][4];
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s) S[r][s] = 0.166666666666667 * (det * dot3(g[r], g[s]));
  // edge weights (:4852-4868) — same operation order as the reference for the
  // cancellation-prone part: fermi = eta-beta, d = fermi_s-fermi_r, d/(exp(d)-1)
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s) {
      const double fr = eta[r] - beta[r], fs = eta[s] - beta[s];
      const double d = fs - fr;
      double aE = 0.5 * (alpha[r] + alpha[s]) * std::exp(fs);
      aE *= (std::fabs(d) < 3.0e-16) ? 1.0 - 0.5 * d : d / (std::exp(d) - 1.0);
      const double bern = (r == s) ? 0.0 : aE * std::exp(beta[s]);
      A[4 * r + s] = S[r][s] * bern;           // (:4871-4873)
    }
  // zero column sums (:4876-4879)
  for (int r = 0; r < 4; ++r)
    A[4 * r + r] = -A[0 * 4 + r] - A[1 * 4 + r] - A[2 * 4 + r] - A[3 * 4 + r];
  // mass-lumped reaction (:4886-4909)
  for (int ip = 0; ip < 24; ++ip) {
    double F0 = 0.0, F1 = 0.0;
    for (int r = 0; r < 4; ++r) { F0 += P24[ip][r] * gamma[r]; F1 += P24[ip][r] * eta[r]; }
    const double reaction = F0 * W24[ip] * det * std::exp(F1);
    for (int j = 0; j < 4; ++j) A[4 * j + j] += P24[ip][j] * reaction;
  }
}

// ---- Galerkin Jacobian (forms.ufl:27-34; generated forms.h:7917-8360) --------------------
// local dof order is component-major: phi x4, c1 x4, c2 x4 (forms.h:7706-7726)
void jac_element(double* A, const double* uu, const double* eps, const double* D, const double* q,
                 const double* x)
{
  init_tables();
  double g[4][3], det;
  tet_geometry(x, g, &det);
  double G[4][4];
  for (int r = 0; r < 4; ++r) for (int s = 0; s < 4; ++s) G[r][s] = det * dot3(g[r], g[s]);
  for (int i = 0; i < 144; ++i) A[i] = 0.0;
  // constant gradients of the previous iterate and of the valency field
  double gu[3][3] = {{0}}, gq[3][3] = {{0}};
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 4; ++r)
      for (int d = 0; d < 3; ++d) { gu[k][d] += uu[4 * k + r] * g[r][d]; gq[k][d] += q[4 * k + r] * g[r][d]; }
  for (int ip = 0; ip < 24; ++ip) {
    const double* f = P24[ip];
    const double W = W24[ip];
    double e_ip = 0.0, u0 = 0.0;
    for (int r = 0; r < 4; ++r) { e_ip += f[r] * eps[r]; u0 += f[r] * uu[r]; }
    // (0,0): permittivity * grad u0 . grad v0
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) A[i * 12 + j] += W * e_ip * G[i][j];
    for (int k = 1; k < 3; ++k) {
      double Dk = 0.0, qk = 0.0, uk = 0.0;
      for (int r = 0; r < 4; ++r) { Dk += f[r] * D[4 * k + r]; qk += f[r] * q[4 * k + r]; uk += f[r] * uu[4 * k + r]; }
      const double ex = std::exp(uk);
      const double cx = (g_conc_cutoff && !(uk > -50.0)) ? 0.0 : ex;      // diode conc(w)
      // drift vector  grad(uu_k + q_k uu_0) = grad uu_k + q grad uu_0 + uu_0 grad q
      double bv[3];
      for (int d = 0; d < 3; ++d) bv[d] = gu[k][d] + qk * gu[0][d] + u0 * gq[k][d];
      for (int i = 0; i < 4; ++i) {
        const double bi = det * dot3(bv, g[i]);
        for (int j = 0; j < 4; ++j) {
          // (k,k): D conc(u) (grad u_k + drift * u_k) . grad v_k
          A[(4 * k + i) * 12 + 4 * k + j] += W * Dk * cx * (G[i][j] + bi * f[j]);
          // (k,0): q D conc(u) grad u_0 . grad v_k
          A[(4 * k + i) * 12 + j] += W * qk * Dk * cx * G[i][j];
          // (0,k): -poisson_scale q e^{u} u_k v_0
          A[i * 12 + 4 * k + j] += -(W * det) * g_poisson_scale * qk * ex * f[i] * f[j];
        }
      }
    }
  }
}

// ---- residual (forms.ufl:36-43; generated forms.h:8385-8644) ------------------------------
void res_element(double* b, const double* uu, const double* eps, const double* fc, const double* D,
                 const double* q, const double* reac, const double* x)
{
  init_tables();
  double g[4][3], det;
  tet_geometry(x, g, &det);
  for (int i = 0; i < 12; ++i) b[i] = 0.0;
  double gu[3][3] = {{0}}, gq[3][3] = {{0}};
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 4; ++r)
      for (int d = 0; d < 3; ++d) { gu[k][d] += uu[4 * k + r] * g[r][d]; gq[k][d] += q[4 * k + r] * g[r][d]; }
  for (int ip = 0; ip < 15; ++ip) {
    const double* f = P15[ip];
    const double W = W15[ip];
    double e_ip = 0.0, u0 = 0.0, fcv = 0.0;
    for (int r = 0; r < 4; ++r) { e_ip += f[r] * eps[r]; u0 += f[r] * uu[r]; fcv += f[r] * fc[r]; }
    double src = fcv;
    for (int k = 1; k < 3; ++k) {
      double Dk = 0.0, qk = 0.0, uk = 0.0, rk = 0.0;
      for (int r = 0; r < 4; ++r) {
        Dk += f[r] * D[4 * k + r]; qk += f[r] * q[4 * k + r];
        uk += f[r] * uu[4 * k + r]; rk += f[r] * reac[4 * k + r];
      }
      const double ex = std::exp(uk);
      src += qk * ex;
      double bv[3];
      for (int d = 0; d < 3; ++d) bv[d] = gu[k][d] + qk * gu[0][d] + u0 * gq[k][d];
      for (int i = 0; i < 4; ++i)
        b[4 * k + i] += W * det * rk * f[i] - W * Dk * ex * (det * dot3(bv, g[i]));
    }
    for (int i = 0; i < 4; ++i)
      b[i] += W * det * g_poisson_scale * src * f[i] - W * e_ip * (det * dot3(gu[0], g[i]));
  }
}

// ---- DG0 entropy indicator of the refinement loop (ufl/poisson_cell_marker.ufl:14-15,
// include/poisson_cell_marker.h:5255-5422): L = int D exp(w) |grad(mu) - E|^2 dx with the 24-point
// rule; the reference always binds E = 0 (src/mesh_refiner.cpp:250-251), so the integrand is
// D exp(w) |grad mu|^2 with the cell-constant P1 gradient. Divided by the DG0 mass
// 0.166666666666667 |det J| (:5203-5231) by mass_lumping_solver (src/mesh_refiner.cpp:297-334).
double marker_element(const double* mu, const double* logw, const double* D, const double* x)
{
  init_tables();
  double g[4][3], det;
  tet_geometry(x, g, &det);
  double gr[3] = {0, 0, 0};
  for (int r = 0; r < 4; ++r) for (int d = 0; d < 3; ++d) gr[d] += mu[r] * g[r][d];
  const double g2 = det * (gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
  double A = 0.0;
  for (int ip = 0; ip < 24; ++ip) {
    double F0 = 0.0, F1 = 0.0;
    for (int r = 0; r < 4; ++r) { F0 += P24[ip][r] * D[r]; F1 += P24[ip][r] * logw[r]; }
    A += F0 * W24[ip] * std::exp(F1) * g2;
  }
  return A / (0.166666666666667 * det);
}

// ---- kernel dispatch (restated vs the reference's own compiled kernels) --------------------
ElementKernels kernels = {eafe_element, jac_element, res_element, 0};

} // namespace orc

extern "C" {

void orc_eafe_element(double* A16, const double* eta4, const double* alpha4, const double* beta4,
                      const double* gamma4, const double* xyz12)
{ orc::eafe_element(A16, eta4, alpha4, beta4, gamma4, xyz12); }

void orc_jac_element(double* A144, const double* uu12, const double* eps4, const double* diff12,
                     const double* val12, const double* xyz12)
{ orc::jac_element(A144, uu12, eps4, diff12, val12, xyz12); }

void orc_res_element(double* b12, const double* uu12, const double* eps4, const double* fc4,
                     const double* diff12, const double* val12, const double* reac12,
                     const double* xyz12)
{ orc::res_element(b12, uu12, eps4, fc4, diff12, val12, reac12, xyz12); }

double orc_marker_element(const double* mu4, const double* logw4, const double* diff4, const double* xyz12)
{ return orc::marker_element(mu4, logw4, diff4, xyz12); }

int orc_use_ref_kernels(int on, const char* libpath)
{
  if (!on) {
    orc::kernels = {orc::eafe_element, orc::jac_element, orc::res_element, 0};
    orc::g_ref_set_scale = nullptr;
    return 0;
  }
  void* h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  auto e = (orc::eafe_fn)dlsym(h, "ref_eafe_tabulate");
  auto j = (orc::jac_fn)dlsym(h, "ref_jac_tabulate");
  auto r = (orc::res_fn)dlsym(h, "ref_res_tabulate");
  if (!e || !j || !r) return -1;
  orc::kernels = {e, j, r, 1};
  orc::g_ref_set_scale = (void (*)(double))dlsym(h, "ref_set_poisson_scale");   // diode kernels only
  if (orc::g_ref_set_scale) orc::g_ref_set_scale(orc::g_poisson_scale);
  return 0;
}

void orc_set_form_variant(double poisson_scale, int conc_cutoff)
{
  orc::g_poisson_scale = poisson_scale;
  orc::g_conc_cutoff = conc_cutoff;
  if (orc::g_ref_set_scale) orc::g_ref_set_scale(poisson_scale);
}

}
