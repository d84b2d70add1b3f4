// TEST INFRASTRUCTURE ONLY (oracle/). Block ILU(k) for BSR(3) and its use as the Krylov
// preconditioner: CPU restatement of FASP's fasp_solver_dbsr_krylov_ilu [EXT: FASP is not in
// /root/reference, no version pinned, CMakeLists.txt:40-48] as it is called at
// benchmarks/pnp_diode/linear_pnp.cpp:103-109 with benchmarks/pnp_diode/bsr.dat:34-38
// (ILU_type = 2 "ILUk", ILU_lfil = 2). Restated from the published algorithm (Saad, Iterative
// Methods, Alg. 10.5/10.6, on 3x3 blocks):
//   symbolic : level of fill. lev(a_ij) = 0 on the pattern of A; eliminating row k from row i
//              creates (i,j) with lev = lev(i,k) + lev(k,j) + 1; entries with lev <= lfil are kept.
//   numeric  : IKJ on that pattern, L unit lower triangular, inverses of the U_kk stored.
//   apply    : z = U^{-1} L^{-1} r.
// PARITY: unpinned against real FASP (nothing in the reference tree holds ILU factors or the
// iteration counts they give); the CUDA path (csrc/ilu.cu) is checked against THIS restatement.
#include "pnp_oracle.h"
#include "oracle_internal.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

namespace orc {
int pvfgmres_impl(int nrow, const int* IA, const int* JA, const double* val, const double* b, double* x,
                  const std::function<void(const double*, double*)>* pc, const orc_solver_param* prm, double* hist, int hist_len);
}

struct orc_ilu {
  int n = 0;
  std::vector<int> IA, JA, diag;       // factor pattern (columns ascending), position of the diagonal block
  std::vector<double> val;             // L (strict lower, unit diagonal implied) and U blocks, row-major 3x3
  std::vector<double> dinv;            // inverses of the U_ii
};

namespace {
inline void inv3(const double* a, double* o)
{
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double id = 1.0 / (a[0] * c0 + a[1] * c1 + a[2] * c2);
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}
inline void mul3(const double* a, const double* b, double* c)       // c = a b
{
  for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) c[3 * r + q] = a[3 * r] * b[q] + a[3 * r + 1] * b[3 + q] + a[3 * r + 2] * b[6 + q];
}
inline void submul3(const double* a, const double* b, double* c)    // c -= a b
{
  for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) c[3 * r + q] -= a[3 * r] * b[q] + a[3 * r + 1] * b[3 + q] + a[3 * r + 2] * b[6 + q];
}
}

extern "C" {

// symbolic factorisation: pattern of L + U with level of fill <= lfil (columns ascending per row)
int orc_ilu_symbolic_count(int n, const int* IA, const int* JA, int lfil, int* fIA)
{
  std::vector<std::vector<std::pair<int, int>>> rows(n);   // (column, level), ascending column — U parts reused
  fIA[0] = 0;
  std::vector<int> lev(n, -1), cols;
  for (int i = 0; i < n; ++i) {
    cols.clear();
    for (int p = IA[i]; p < IA[i + 1]; ++p) { lev[JA[p]] = 0; cols.push_back(JA[p]); }
    if (lev[i] < 0) { lev[i] = 0; cols.push_back(i); }
    std::sort(cols.begin(), cols.end());
    for (size_t q = 0; q < cols.size(); ++q) {              // cols grows while we walk it (kept sorted)
      const int k = cols[q];
      if (k >= i) break;
      const int lik = lev[k];
      for (const auto& e : rows[k]) {
        if (e.first <= k) continue;
        const int nl = lik + e.second + 1;
        if (nl > lfil) continue;
        if (lev[e.first] < 0) { lev[e.first] = nl; cols.insert(std::upper_bound(cols.begin(), cols.end(), e.first), e.first); }
        else if (nl < lev[e.first]) lev[e.first] = nl;
      }
    }
    rows[i].reserve(cols.size());
    for (int c : cols) { rows[i].emplace_back(c, lev[c]); lev[c] = -1; }
    fIA[i + 1] = fIA[i] + (int)cols.size();
  }
  return fIA[n];
}

orc_ilu* orc_ilu_setup(int n, const int* IA, const int* JA, const double* val, int lfil)
{
  orc_ilu* F = new orc_ilu;
  F->n = n;
  // symbolic (same walk as the count, keeping the pattern)
  std::vector<std::vector<std::pair<int, int>>> rows(n);
  std::vector<int> lev(n, -1), cols;
  F->IA.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    cols.clear();
    for (int p = IA[i]; p < IA[i + 1]; ++p) { lev[JA[p]] = 0; cols.push_back(JA[p]); }
    if (lev[i] < 0) { lev[i] = 0; cols.push_back(i); }
    std::sort(cols.begin(), cols.end());
    for (size_t q = 0; q < cols.size(); ++q) {
      const int k = cols[q];
      if (k >= i) break;
      const int lik = lev[k];
      for (const auto& e : rows[k]) {
        if (e.first <= k) continue;
        const int nl = lik + e.second + 1;
        if (nl > lfil) continue;
        if (lev[e.first] < 0) { lev[e.first] = nl; cols.insert(std::upper_bound(cols.begin(), cols.end(), e.first), e.first); }
        else if (nl < lev[e.first]) lev[e.first] = nl;
      }
    }
    for (int c : cols) { rows[i].emplace_back(c, lev[c]); lev[c] = -1; F->JA.push_back(c); }
    F->IA[i + 1] = (int)F->JA.size();
  }
  // numeric IKJ
  const int nnz = F->IA[n];
  F->val.assign(9 * (size_t)nnz, 0.0); F->dinv.assign(9 * (size_t)n, 0.0); F->diag.assign(n, -1);
  std::vector<int> where(n, -1);
  for (int i = 0; i < n; ++i) {
    for (int p = F->IA[i]; p < F->IA[i + 1]; ++p) { where[F->JA[p]] = p; if (F->JA[p] == i) F->diag[i] = p; }
    for (int p = IA[i]; p < IA[i + 1]; ++p) std::memcpy(&F->val[9 * (size_t)where[JA[p]]], &val[9 * (size_t)p], 9 * sizeof(double));
    for (int p = F->IA[i]; p < F->IA[i + 1]; ++p) {
      const int k = F->JA[p];
      if (k >= i) break;
      double lik[9];
      mul3(&F->val[9 * (size_t)p], &F->dinv[9 * (size_t)k], lik);                 // L_ik = A_ik U_kk^{-1}
      std::memcpy(&F->val[9 * (size_t)p], lik, sizeof lik);
      for (int q = F->diag[k] + 1; q < F->IA[k + 1]; ++q) {
        const int w = where[F->JA[q]];
        if (w >= 0) submul3(lik, &F->val[9 * (size_t)q], &F->val[9 * (size_t)w]);  // A_ij -= L_ik U_kj
      }
    }
    inv3(&F->val[9 * (size_t)F->diag[i]], &F->dinv[9 * (size_t)i]);
    for (int p = F->IA[i]; p < F->IA[i + 1]; ++p) where[F->JA[p]] = -1;
  }
  return F;
}

int  orc_ilu_nnz(const orc_ilu* F) { return F->IA[F->n]; }
void orc_ilu_factors(const orc_ilu* F, int* IA, int* JA, double* val, double* dinv)
{
  std::copy(F->IA.begin(), F->IA.end(), IA); std::copy(F->JA.begin(), F->JA.end(), JA);
  std::copy(F->val.begin(), F->val.end(), val); std::copy(F->dinv.begin(), F->dinv.end(), dinv);
}
void orc_ilu_free(orc_ilu* F) { delete F; }

// z = U^{-1} L^{-1} r
void orc_ilu_apply(const orc_ilu* F, const double* r, double* z)
{
  const int n = F->n;
  for (int i = 0; i < n; ++i) {
    double t[3] = {r[3 * i], r[3 * i + 1], r[3 * i + 2]};
    for (int p = F->IA[i]; p < F->diag[i]; ++p) {
      const double* a = &F->val[9 * (size_t)p]; const double* y = &z[3 * F->JA[p]];
      for (int c = 0; c < 3; ++c) t[c] -= a[3 * c] * y[0] + a[3 * c + 1] * y[1] + a[3 * c + 2] * y[2];
    }
    z[3 * i] = t[0]; z[3 * i + 1] = t[1]; z[3 * i + 2] = t[2];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t[3] = {z[3 * i], z[3 * i + 1], z[3 * i + 2]};
    for (int p = F->diag[i] + 1; p < F->IA[i + 1]; ++p) {
      const double* a = &F->val[9 * (size_t)p]; const double* y = &z[3 * F->JA[p]];
      for (int c = 0; c < 3; ++c) t[c] -= a[3 * c] * y[0] + a[3 * c + 1] * y[1] + a[3 * c + 2] * y[2];
    }
    const double* d = &F->dinv[9 * (size_t)i];
    for (int c = 0; c < 3; ++c) z[3 * i + c] = d[3 * c] * t[0] + d[3 * c + 1] * t[1] + d[3 * c + 2] * t[2];
  }
}

// fasp_solver_dbsr_krylov_ilu restated: ILU(lfil) setup + pvfgmres preconditioned by it
int orc_solver_dbsr_krylov_ilu(int nrow, const int* IA, const int* JA, const double* val, const double* b, double* x,
                               const orc_solver_param* p, int lfil)
{
  orc_ilu* F = orc_ilu_setup(nrow, IA, JA, val, lfil);
  std::function<void(const double*, double*)> pc = [&](const double* r, double* z) { orc_ilu_apply(F, r, z); };
  const int its = orc::pvfgmres_impl(nrow, IA, JA, val, b, x, &pc, p, nullptr, 0);
  orc_ilu_free(F);
  return its;
}

}
