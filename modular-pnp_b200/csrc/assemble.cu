// pnp-b200 — device assembly of the linearised PNP system:
//   PDE::setup_linear_algebra            src/pde.cpp:542-555   (Galerkin a, L; Dirichlet rows)
//   Linear_PNP::apply_eafe               benchmarks/pnp_exact_solution/linear_pnp.cpp:195-285
//   BC re-apply after EAFE               linear_pnp.cpp:78-80
//   zero-row fix                         src/pde.cpp:596-621
// Two passes, both atomic-free and deterministic (a sorted-key scatter: the key of a contribution is the
// slot it is written to, fixed per mesh in mesh_setup.cu: build_assembly_tables):
//   (A) cell kernels evaluate the quadrature sums of the generated forms
//       (forms.h:7917-8360 bilinear, 24-point rule; :8385-8644 linear, 15-point rule) once per
//       cell (geometry is constant per P1 cell, so the 12x12 tensor factorises into 32 numbers) and
//       write, per local vertex pair, one 32-byte record (Q1 G, Q2 G, det M1, det M2) into the run of
//       that vertex pair — runs are contiguous and ordered by ascending cell id;
//   (B) block kernels own one 3x3 block (i,j) each, sum the run of {i,j} (one 256-bit load per
//       cell, the same run serves (i,j) and (j,i)), evaluate the EAFE edge formula
//       (include/EAFE.h:4851-4879) for the (k,k) entries and write the finished block once, in
//       BSR-T layout.
// The residual works the same way with 24-byte records per (cell, vertex). Round 1 wrote one
// 256-byte record per cell and gathered 24- and 16-byte pieces of it through L2 (40 ms per
// Jacobian, 8 ms per residual at 256^3, 3-4 sectors fetched per piece used).
#include "internal.hpp"
#include "fem.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace pnpb200 {

namespace {

#ifndef PNPB200_ASM_CTAS
#define PNPB200_ASM_CTAS 4      // resident CTAs per SM the pair-run block kernel is compiled for
#endif
constexpr int REC_EAFE = 32;   // 10 x (G, det*M1, det*M2) pair-major, then Q[2]
constexpr int REC_FULL = 68;   // + S[2] T1[16] T2[16] (+2 pad)

// ---- (A) Jacobian cell records -------------------------------------------------------------
__device__ __forceinline__ void store_rec4(double* p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void load_rec4(const double* p, double& a, double& b, double& c, double& d)
{
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// FULL = false (EAFE on): pair records through cpslot. FULL = true (Galerkin-only Jacobian, no_eafe()):
// the 68-number cell record of round 1 (the (k,k) terms are not symmetric in the vertex pair).
template <bool FULL, int MINB>
__global__ void __launch_bounds__(128, MINB)
cell_jacobian_kernel(int nc, const int4* __restrict__ cells, const double* __restrict__ xyz,
                     const double* __restrict__ uu, const double* __restrict__ diff,
                     const double* __restrict__ val, double pscale, int cutoff, const int* __restrict__ cpslot,
                     double* __restrict__ rec)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int4 cv = cells[c];
  const int v[4] = {cv.x, cv.y, cv.z, cv.w};
  double x[4][3], g[4][3], det;
  load_cell_xyz(xyz, cv, x);
  tet_geom(x, g, det);
  double u0[4], u[2][4], D[2][4], q[2][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const size_t d = 3 * (size_t)v[r];
    u0[r] = uu[d]; u[0][r] = uu[d + 1]; u[1][r] = uu[d + 2];
    D[0][r] = diff[d + 1]; D[1][r] = diff[d + 2];
    q[0][r] = val[d + 1];  q[1][r] = val[d + 2];
  }
  double Q[2] = {0.0, 0.0}, M[2][10];
  double S[2] = {0.0, 0.0}, c0[2][4], c1[2][4], c2[2][4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int t = 0; t < 10; ++t) M[k][t] = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) { c0[k][t] = 0.0; c1[k][t] = 0.0; c2[k][t] = 0.0; }
  }
#pragma unroll 1
  for (int ip = 0; ip < 24; ++ip) {
    const double f[4] = {c_P24[ip][0], c_P24[ip][1], c_P24[ip][2], c_P24[ip][3]};
    const double W = c_W24[ip];
    const double u0ip = f[0] * u0[0] + f[1] * u0[1] + f[2] * u0[2] + f[3] * u0[3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double uk = f[0] * u[k][0] + f[1] * u[k][1] + f[2] * u[k][2] + f[3] * u[k][3];
      const double Dk = f[0] * D[k][0] + f[1] * D[k][1] + f[2] * D[k][2] + f[3] * D[k][3];
      const double qk = f[0] * q[k][0] + f[1] * q[k][1] + f[2] * q[k][2] + f[3] * q[k][3];
      const double ex = exp(uk);
      const double cx = (cutoff && !(uk > -50.0)) ? 0.0 : ex;     // diode conc(w), pnp_diode forms.h:8606-8608
      Q[k] += W * qk * Dk * cx;
      const double t = W * qk * ex * pscale;                      // poisson_scale, pnp_diode forms.h:8721-8724
      int m = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = r; s < 4; ++s) M[k][m++] += t * f[r] * f[s];
      if (FULL) {
        const double wd = W * Dk * cx;
        S[k] += wd;
#pragma unroll
        for (int j = 0; j < 4; ++j) { c0[k][j] += wd * f[j]; c1[k][j] += wd * qk * f[j]; c2[k][j] += wd * u0ip * f[j]; }
      }
    }
  }
  if (!FULL) {
    // one 32-byte record per local vertex pair, into the run of that pair
    int m = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s) {
        const double G = det * dot3(g[r], g[s]);
        store_rec4(rec + 4 * (size_t)cpslot[10 * (size_t)c + m], Q[0] * G, Q[1] * G, det * M[0][m], det * M[1][m]);
        ++m;
      }
    return;
  }
  double* o = rec + (size_t)c * REC_FULL;
  // pair-major: the three numbers a block (i,j) needs from this cell are contiguous
  int m = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = r; s < 4; ++s) { o[3 * m] = det * dot3(g[r], g[s]); o[3 * m + 1] = det * M[0][m]; o[3 * m + 2] = det * M[1][m]; ++m; }
  o[30] = Q[0]; o[31] = Q[1];
  if (FULL) {
    o[32] = S[0]; o[33] = S[1];
    double gu0[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int d = 0; d < 3; ++d) gu0[d] += u0[r] * g[r][d];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double guk[3] = {0, 0, 0}, gqk[3] = {0, 0, 0};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) { guk[d] += u[k][r] * g[r][d]; gqk[d] += q[k][r] * g[r][d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double a = det * dot3(guk, g[i]), b = det * dot3(gu0, g[i]), cc = det * dot3(gqk, g[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[34 + 16 * k + 4 * i + j] = a * c0[k][j] + b * c1[k][j] + cc * c2[k][j];
      }
    }
  }
}

// EAFE edge weight of the ordered pair (row r, column s) for ion k, exactly the expression of
// include/EAFE.h:4857-4867 (no expm1, same 3e-16 branch): returns bern[r][s]
__device__ __forceinline__ double eafe_weight(const double* __restrict__ uu, const double* __restrict__ diff,
                                              int r, int s, int k, double qk)
{
  const double eta_r = uu[3 * (size_t)r + k], eta_s = uu[3 * (size_t)s + k];
  double beta_r = eta_r, beta_s = eta_s;
  if (qk != 0.0) { beta_r = eta_r + uu[3 * (size_t)r] * qk; beta_s = eta_s + uu[3 * (size_t)s] * qk; }
  const double fr = eta_r - beta_r, fs = eta_s - beta_s;
  const double d = fs - fr;
  double aE = 0.5 * (diff[3 * (size_t)r + k] + diff[3 * (size_t)s + k]) * exp(fs);
  aE *= (fabs(d) < 3.0e-16) ? 1.0 - 0.5 * d : d / (exp(d) - 1.0);
  return aE * exp(beta_s);
}

// Per-vertex operands of the EAFE edge weights, one 32-byte entry per (vertex, ion): (f, exp(f), exp(beta), D)
// with beta = eta + q phi, f = eta - beta — exactly the quantities eafe_weight() forms at an edge end, so
// the weights built from the table are bit-identical to evaluating include/EAFE.h:4857-4867 per edge,
// with one exp per directed edge and ion (the B(d) denominator) instead of three.
__global__ void eafe_vertex_table_kernel(int nv, const double* __restrict__ uu, const double* __restrict__ diff,
                                         double q1, double q2, double* __restrict__ tab)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const double phi = uu[3 * (size_t)v];
#pragma unroll
  for (int k = 1; k <= 2; ++k) {
    const double qk = k == 1 ? q1 : q2;
    const double eta = uu[3 * (size_t)v + k];
    double beta = eta;
    if (qk != 0.0) beta = eta + phi * qk;
    const double f = eta - beta;
    store_rec4(tab + 8 * (size_t)v + 4 * (k - 1), f, exp(f), exp(beta), diff[3 * (size_t)v + k]);
  }
}

// bern[r][s] of EAFE.h:4857-4867 from the two table entries (r = row end, s = column end)
__device__ __forceinline__ double eafe_weight_tab(double fr, double Dr, double fs, double efs, double ebs, double Ds)
{
  const double d = fs - fr;
  double aE = 0.5 * (Dr + Ds) * efs;
  aE *= (fabs(d) < 3.0e-16) ? 1.0 - 0.5 * d : d / (exp(d) - 1.0);
  return aE * ebs;
}

// ---- (B), EAFE on: a half-warp per block row, lane k = block k of the row -------------------------
// Lane k sums the record run of the vertex pair {i, j_k} (ascending cell id; the same run, hence the
// same rounding, serves (i,j) and (j,i)) and evaluates the two EAFE edge weights of its block and their
// transposes; the EAFE diagonal (zero column sums, EAFE.h:4876-4879) is a half-warp reduction of the
// transposed weights. The run of the DIAGONAL pair {i,i} (every cell of the vertex, 24 on a Kuhn mesh
// against 4-6 for an edge) is summed by the whole half-warp, two records per lane and a fixed shuffle
// tree, so that no lane loops four times longer than its neighbours. Rows are visited in natural
// order, the finished blocks go to the row's physical, colour-major position.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
assemble_pairs_kernel(int n, const int4* __restrict__ rd, const int* __restrict__ ja, const int2* __restrict__ pblk,
                      const double* __restrict__ rec, const double* __restrict__ k00, const double* __restrict__ omega,
                      const double* __restrict__ etab, const unsigned char* __restrict__ bcflag, double* __restrict__ val)
{
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);      // my half-warp
  int s = 0, len = 0, nl = 0;
  bool bc0 = false, bc1 = false, bc2 = false;
  double fi1 = 0, efi1 = 0, ebi1 = 0, Di1 = 0, fi2 = 0, efi2 = 0, ebi2 = 0, Di2 = 0;
  double dg1 = 0.0, dg2 = 0.0, dm1 = 0.0, dm2 = 0.0;           // sums of the diagonal pair's run
  if (i < n) {
    const int4 d = rd[i];
    s = d.x; len = d.y; nl = d.z;
    bc0 = bcflag[3 * (size_t)i] != 0; bc1 = bcflag[3 * (size_t)i + 1] != 0; bc2 = bcflag[3 * (size_t)i + 2] != 0;
    load_rec4(etab + 8 * (size_t)i, fi1, efi1, ebi1, Di1);
    load_rec4(etab + 8 * (size_t)i + 4, fi2, efi2, ebi2, Di2);
    const int2 run = pblk[s + nl];                              // the diagonal block sits behind the nl lower blocks
    const double* r = rec + 4 * (size_t)run.x;
    for (int q = lane; q < run.y; q += 16) {
      double g1, g2, m1, m2;
      load_rec4(r + 4 * (size_t)q, g1, g2, m1, m2);
      dg1 += g1; dg2 += g2; dm1 += m1; dm2 += m2;
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    dg1 += __shfl_xor_sync(hmask, dg1, o); dg2 += __shfl_xor_sync(hmask, dg2, o);
    dm1 += __shfl_xor_sync(hmask, dm1, o); dm2 += __shfl_xor_sync(hmask, dm2, o);
  }
  const int npass = (len + 15) >> 4;                          // uniform over the half-warp
  double d1 = 0.0, d2 = 0.0;                                  // -(column sums) of the EAFE off-diagonals
  int diag_k = -1; double diag_a[9];
  for (int ps = 0; ps < npass; ++ps) {
    const int k = ps * 16 + lane;
    const bool act = k < len;
    const int b = s + (act ? k : 0), j = act ? ja[b] : -1;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (act) a[0] = k00[b];
    if (act && i == j) {                          // keep the diagonal block until the column sums are known
      a[3] = dg1; a[6] = dg2; a[1] = -dm1; a[2] = -dm2;
      diag_k = k;
#pragma unroll
      for (int t = 0; t < 9; ++t) diag_a[t] = a[t];
      continue;
    }
    if (act) {
      const int2 run = pblk[b];
      const double* r = rec + 4 * (size_t)run.x;
#pragma unroll 2
      for (int q = 0; q < run.y; ++q) {
        double g1, g2, m1, m2;
        load_rec4(r + 4 * (size_t)q, g1, g2, m1, m2);
        a[3] += g1;                 // (1,0): q1 D1 e^{u1} grad u0 . grad v1
        a[6] += g2;                 // (2,0)
        a[1] -= m1;                 // (0,1): -q1 e^{u1} u1 v0
        a[2] -= m2;                 // (0,2)
      }
      const double w = omega[b];                 // omega is symmetric: the same weight serves A_ij and A_ji
      double fj, efj, ebj, Dj;
      load_rec4(etab + 8 * (size_t)j, fj, efj, ebj, Dj);
      a[4] = w * eafe_weight_tab(fi1, Di1, fj, efj, ebj, Dj);
      d1 -= w * eafe_weight_tab(fj, Dj, fi1, efi1, ebi1, Di1);
      load_rec4(etab + 8 * (size_t)j + 4, fj, efj, ebj, Dj);
      a[8] = w * eafe_weight_tab(fi2, Di2, fj, efj, ebj, Dj);
      d2 -= w * eafe_weight_tab(fj, Dj, fi2, efi2, ebi2, Di2);
      // DirichletBC::apply: row -> identity row (pattern kept)
      if (bc0) { a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; }
      if (bc1) { a[3] = 0.0; a[4] = 0.0; a[5] = 0.0; }
      if (bc2) { a[6] = 0.0; a[7] = 0.0; a[8] = 0.0; }
      int st;
      double* o = val + bsrt_off(s, len, nl, k, st);
#pragma unroll
      for (int t = 0; t < 9; ++t) o[(size_t)t * st] = a[t];
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { d1 += __shfl_xor_sync(hmask, d1, o); d2 += __shfl_xor_sync(hmask, d2, o); }
  if (diag_k >= 0) {
    diag_a[4] = d1; diag_a[8] = d2;
    if (bc0) { diag_a[0] = 1.0; diag_a[1] = 0.0; diag_a[2] = 0.0; }
    if (bc1) { diag_a[3] = 0.0; diag_a[4] = 1.0; diag_a[5] = 0.0; }
    if (bc2) { diag_a[6] = 0.0; diag_a[7] = 0.0; diag_a[8] = 1.0; }
    int st;
    double* o = val + bsrt_off(s, len, nl, diag_k, st);
#pragma unroll
    for (int t = 0; t < 9; ++t) o[(size_t)t * st] = diag_a[t];
  }
}

// ---- (B), Galerkin-only Jacobian (no_eafe()): the round-1 gather from 68-number cell records -------
// ---- (B) a half-warp per block row, lane k = block k of the row ------------------------------
// Rows are visited in natural order so that the cell records stream through L2 once; the finished
// blocks go to the row's physical, colour-major position. The row's incident cells are loaded
// cooperatively (one cell per lane) and broadcast with shuffles in ascending cell order, so every
// lane accumulates its block in the same deterministic order without re-reading the cell list.
// The EAFE diagonal (zero column sums, EAFE.h:4876-4879) is a half-warp reduction of the transposed
// edge weights each off-diagonal lane evaluates next to its own.
template <bool EAFE, int MINB>
__global__ void __launch_bounds__(256, MINB)
assemble_blocks_kernel(int n, const int4* __restrict__ rd, const int* __restrict__ ja,
                       const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                       const int4* __restrict__ cells, const double* __restrict__ rec,
                       const double* __restrict__ k00, const double* __restrict__ omega,
                       const double* __restrict__ uu, const double* __restrict__ diff, double q1, double q2,
                       const unsigned char* __restrict__ bcflag, double* __restrict__ val)
{
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);      // my half-warp
  constexpr int RS = EAFE ? REC_EAFE : REC_FULL;
  int s = 0, len = 0, nl = 0, c0 = 0, c1 = 0;
  bool bc0 = false, bc1 = false, bc2 = false;
  if (i < n) {
    const int4 d = rd[i];
    s = d.x; len = d.y; nl = d.z; c0 = v2c_ptr[i]; c1 = v2c_ptr[i + 1];
    bc0 = bcflag[3 * (size_t)i] != 0; bc1 = bcflag[3 * (size_t)i + 1] != 0; bc2 = bcflag[3 * (size_t)i + 2] != 0;
  }
  const int npass = (len + 15) >> 4;                          // uniform over the half-warp
  double d1 = 0.0, d2 = 0.0;                                  // -(column sums) of the EAFE off-diagonals
  int diag_k = -1; double diag_a[9];
  for (int ps = 0; ps < npass; ++ps) {
    const int k = ps * 16 + lane;
    const bool act = k < len;
    const int b = s + (act ? k : 0), j = act ? ja[b] : -1;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (act) a[0] = k00[b];
    // cells of row i, 16 at a time: lane q loads cell c0 + base + q, everybody reads it by shuffle
    for (int base = c0; base < c1; base += 16) {
      const int mine = base + lane < c1 ? v2c_idx[base + lane] : -1;
      int4 mycv = make_int4(-1, -1, -1, -1);
      if (mine >= 0) mycv = cells[mine];
      const int cnt = c1 - base < 16 ? c1 - base : 16;
      for (int q = 0; q < cnt; ++q) {
        const int src = (threadIdx.x & 16) + q;
        int4 cv;
        cv.x = __shfl_sync(hmask, mycv.x, src); cv.y = __shfl_sync(hmask, mycv.y, src);
        cv.z = __shfl_sync(hmask, mycv.z, src); cv.w = __shfl_sync(hmask, mycv.w, src);
        const int c = __shfl_sync(hmask, mine, src);
        const int lj = act ? local_index(cv, j) : -1;
        if (lj < 0) continue;
        const int li = local_index(cv, i);
        const double* r = rec + (size_t)c * RS;
        const int sidx = sym4(li, lj);
        const double G = r[3 * sidx];
        a[3] += r[30] * G;          // (1,0): q1 D1 e^{u1} grad u0 . grad v1
        a[6] += r[31] * G;          // (2,0)
        a[1] -= r[3 * sidx + 1];    // (0,1): -q1 e^{u1} u1 v0
        a[2] -= r[3 * sidx + 2];    // (0,2)
        if (!EAFE) {
          a[4] += r[32] * G + r[34 + 4 * li + lj];
          a[8] += r[33] * G + r[50 + 4 * li + lj];
        }
      }
    }
    if (EAFE && act && i != j) {
      const double w = omega[b];                 // omega is symmetric: the same weight serves A_ij and A_ji
      a[4] = w * eafe_weight(uu, diff, i, j, 1, q1);
      a[8] = w * eafe_weight(uu, diff, i, j, 2, q2);
      d1 -= w * eafe_weight(uu, diff, j, i, 1, q1);
      d2 -= w * eafe_weight(uu, diff, j, i, 2, q2);
    }
    if (act && i == j) {                          // keep the diagonal block until the column sums are known
      diag_k = k;
#pragma unroll
      for (int t = 0; t < 9; ++t) diag_a[t] = a[t];
      continue;
    }
    if (act) {
      // DirichletBC::apply: row -> identity row (pattern kept)
      if (bc0) { a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; }
      if (bc1) { a[3] = 0.0; a[4] = 0.0; a[5] = 0.0; }
      if (bc2) { a[6] = 0.0; a[7] = 0.0; a[8] = 0.0; }
      int st;
      double* o = val + bsrt_off(s, len, nl, k, st);
#pragma unroll
      for (int t = 0; t < 9; ++t) o[(size_t)t * st] = a[t];
    }
  }
  if (EAFE) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { d1 += __shfl_xor_sync(hmask, d1, o); d2 += __shfl_xor_sync(hmask, d2, o); }
  }
  if (diag_k >= 0) {
    if (EAFE) { diag_a[4] = d1; diag_a[8] = d2; }
    if (bc0) { diag_a[0] = 1.0; diag_a[1] = 0.0; diag_a[2] = 0.0; }
    if (bc1) { diag_a[3] = 0.0; diag_a[4] = 1.0; diag_a[5] = 0.0; }
    if (bc2) { diag_a[6] = 0.0; diag_a[7] = 0.0; diag_a[8] = 1.0; }
    int st;
    double* o = val + bsrt_off(s, len, nl, diag_k, st);
#pragma unroll
    for (int t = 0; t < 9; ++t) o[(size_t)t * st] = diag_a[t];
  }
}

// zero-row fix (src/pde.cpp:596-621): a scalar row with no non-zero value gets diag = 1.
// Such a row has a zero diagonal, so only rows whose diagonal entry is 0 are scanned.
__global__ void zero_row_fix_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                    const int* __restrict__ dslot, double* __restrict__ val)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = rs[i], len = ia[i + 1] - ia[i], d = dslot[i];
  if (d < 0) return;
  const int nl = d - s;
  int std_;
  double* dg = val + bsrt_off(s, len, nl, nl, std_);
  for (int k = 0; k < 3; ++k) {
    if (dg[(size_t)(3 * k + k) * std_] != 0.0) continue;
    bool nz = false;
    for (int p = 0; p < len && !nz; ++p) {
      int st;
      const double* v = val + bsrt_off(s, len, nl, p, st);
      for (int c = 0; c < 3; ++c)
        if (v[(size_t)(3 * k + c) * st] != 0.0) { nz = true; break; }
    }
    if (!nz) dg[(size_t)(3 * k + k) * std_] = 1.0;
  }
}

// ---- residual: (A) per-cell 12-vector, (B) per-vertex gather + BC + norm partials ------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
cell_residual_kernel(int nc, const int4* __restrict__ cells, const double* __restrict__ xyz,
                     const double* __restrict__ uu, const double* __restrict__ eps,
                     const double* __restrict__ fixed, const double* __restrict__ diff,
                     const double* __restrict__ val, const double* __restrict__ reac,
                     double pscale, const int* __restrict__ cvslot, double* __restrict__ rec)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int4 cv = cells[c];
  const int v[4] = {cv.x, cv.y, cv.z, cv.w};
  double x[4][3], g[4][3], det;
  load_cell_xyz(xyz, cv, x);
  tet_geom(x, g, det);
  double u0[4], e4[4], fc[4], u[2][4], D[2][4], q[2][4], rc[2][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const size_t d = 3 * (size_t)v[r];
    u0[r] = uu[d]; u[0][r] = uu[d + 1]; u[1][r] = uu[d + 2];
    D[0][r] = diff[d + 1]; D[1][r] = diff[d + 2];
    q[0][r] = val[d + 1];  q[1][r] = val[d + 2];
    rc[0][r] = reac[d + 1]; rc[1][r] = reac[d + 2];
    e4[r] = eps[v[r]]; fc[r] = fixed[v[r]];
  }
  double F[4] = {0, 0, 0, 0}, E = 0.0, R[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  double s0[2] = {0, 0}, s1[2] = {0, 0}, s2[2] = {0, 0};
#pragma unroll 1
  for (int ip = 0; ip < 15; ++ip) {
    const double f[4] = {c_P15[ip][0], c_P15[ip][1], c_P15[ip][2], c_P15[ip][3]};
    const double W = c_W15[ip];
    const double u0ip = f[0] * u0[0] + f[1] * u0[1] + f[2] * u0[2] + f[3] * u0[3];
    const double eip = f[0] * e4[0] + f[1] * e4[1] + f[2] * e4[2] + f[3] * e4[3];
    double src = f[0] * fc[0] + f[1] * fc[1] + f[2] * fc[2] + f[3] * fc[3];
    E += W * eip;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double uk = f[0] * u[k][0] + f[1] * u[k][1] + f[2] * u[k][2] + f[3] * u[k][3];
      const double Dk = f[0] * D[k][0] + f[1] * D[k][1] + f[2] * D[k][2] + f[3] * D[k][3];
      const double qk = f[0] * q[k][0] + f[1] * q[k][1] + f[2] * q[k][2] + f[3] * q[k][3];
      const double rk = f[0] * rc[k][0] + f[1] * rc[k][1] + f[2] * rc[k][2] + f[3] * rc[k][3];
      const double ex = exp(uk);
      src += qk * ex;
      const double wd = W * Dk * ex;
      s0[k] += wd; s1[k] += wd * qk; s2[k] += wd * u0ip;
#pragma unroll
      for (int i = 0; i < 4; ++i) R[k][i] += W * rk * f[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) F[i] += W * src * f[i];
  }
  double gu0[3] = {0, 0, 0};
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int d = 0; d < 3; ++d) gu0[d] += u0[r] * g[r][d];
  // the three numbers for local vertex i go to that vertex' slot of this cell (3 contiguous doubles)
  const int4 sl = reinterpret_cast<const int4*>(cvslot)[c];
  double* o[4] = {rec + 3 * (size_t)sl.x, rec + 3 * (size_t)sl.y, rec + 3 * (size_t)sl.z, rec + 3 * (size_t)sl.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i][0] = det * pscale * F[i] - E * (det * dot3(gu0, g[i]));
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double guk[3] = {0, 0, 0}, gqk[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int d = 0; d < 3; ++d) { guk[d] += u[k][r] * g[r][d]; gqk[d] += q[k][r] * g[r][d]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i][k + 1] = det * R[k][i]
                    - det * (s0[k] * dot3(guk, g[i]) + s1[k] * dot3(gu0, g[i]) + s2[k] * dot3(gqk, g[i]));
  }
}

__global__ void __launch_bounds__(256)
residual_gather_kernel(int nv, const int* __restrict__ v2c_ptr, const double* __restrict__ rec,
                       const unsigned char* __restrict__ bcflag, double* __restrict__ out,
                       double* __restrict__ partial /* [2*gridDim] sumsq, max */)
{
  __shared__ double sm[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ss = 0.0, mx = 0.0;
  if (i < nv) {
    double b[3] = {0, 0, 0};
    // the records of the vertex' cells: contiguous, ascending cell id
    for (int p = v2c_ptr[i]; p < v2c_ptr[i + 1]; ++p) {
      const double* r = rec + 3 * (size_t)p;
      b[0] += r[0]; b[1] += r[1]; b[2] += r[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (bcflag[3 * (size_t)i + k]) b[k] = 0.0;
      out[3 * (size_t)i + k] = b[k];
      ss += b[k] * b[k];
      mx = fmax(mx, fabs(b[k]));
    }
  }
  ss = block_sum(ss, sm);
  mx = block_max(mx, sm);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = ss; partial[2 * blockIdx.x + 1] = mx; }
}

__global__ void __launch_bounds__(1024) finalize_norm_kernel(int nblk, const double* __restrict__ partial,
                                                            double* __restrict__ out2)
{
  __shared__ double sm[32];
  double ss = 0.0, mx = 0.0;
  for (int t = threadIdx.x; t < nblk; t += blockDim.x) { ss += partial[2 * t]; mx = fmax(mx, partial[2 * t + 1]); }
  ss = block_sum(ss, sm);
  mx = block_max(mx, sm);
  if (threadIdx.x == 0) { out2[0] = ss; out2[1] = mx; }
}

// ---- post-processing functionals ------------------------------------------------
// Error::compute_l2_error / compute_semi_h1_error (src/error.cpp:58-103 with include/L2Error.h,
// include/SemiH1error.h): sum over components of int e_k^2 and int |grad e_k|^2, e = u_h - u_exact
// nodal (P1): exact with the element mass matrix det/120 (1 + delta_rs) and the constant gradient.
__global__ void __launch_bounds__(256)
error_cells_kernel(int nc, const int4* __restrict__ cells, const double* __restrict__ xyz,
                   const double* __restrict__ uu, const double* __restrict__ exact,
                   const int* __restrict__ vowner, int me, double* __restrict__ partial)
{
  __shared__ double sm[32];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double l2 = 0.0, h1 = 0.0;
  bool mine = c < nc;
  if (mine && vowner) {          // multi-GPU: a ghost-layer cell is counted by the lowest owner rank of its vertices
    const int4 cv = cells[c];
    mine = min(min(vowner[cv.x], vowner[cv.y]), min(vowner[cv.z], vowner[cv.w])) == me;
  }
  if (mine) {
    const int4 cv = cells[c];
    const int v[4] = {cv.x, cv.y, cv.z, cv.w};
    double x[4][3], g[4][3], det;
    load_cell_xyz(xyz, cv, x);
    tet_geom(x, g, det);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double e[4], sum = 0.0, sq = 0.0, gr[3] = {0, 0, 0};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        e[r] = uu[3 * (size_t)v[r] + k] - exact[3 * (size_t)v[r] + k];
        sum += e[r]; sq += e[r] * e[r];
#pragma unroll
        for (int d = 0; d < 3; ++d) gr[d] += e[r] * g[r][d];
      }
      l2 += det * (sum * sum + sq) / 120.0;
      h1 += det * dot3(gr, gr) / 6.0;
    }
  }
  l2 = block_sum(l2, sm);
  h1 = block_sum(h1, sm);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = l2; partial[2 * blockIdx.x + 1] = h1; }
}

__global__ void __launch_bounds__(1024) finalize_sum2_kernel(int nblk, const double* __restrict__ partial, double* __restrict__ out2)
{
  __shared__ double sm[32];
  double a = 0.0, b = 0.0;
  for (int t = threadIdx.x; t < nblk; t += blockDim.x) { a += partial[2 * t]; b += partial[2 * t + 1]; }
  a = block_sum(a, sm);
  b = block_sum(b, sm);
  if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
}

// Linear_PNP::get_total_charge (linear_pnp.cpp:322-351): fixed_charge + sum_k q_k exp(u_k), nodal
__global__ void total_charge_kernel(int nv, const double* __restrict__ uu, const double* __restrict__ fixed,
                                    const double* __restrict__ val, double* __restrict__ out)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  out[v] = fixed[v] + val[1] * exp(uu[3 * (size_t)v + 1]) + val[2] * exp(uu[3 * (size_t)v + 2]);
}

} // namespace

// ---- DG0 entropy indicator of the adaptive-refinement loop -----------------------------------
// Mesh_Refiner::compute_entropy_error_vector (src/mesh_refiner.cpp:212-271): per ion species the form
// L = int D exp(w) |grad(mu) - E|^2 v dx of ufl/poisson_cell_marker.ufl:14-15 with E bound to zero
// (:250-251), 24-point rule (include/poisson_cell_marker.h:5255-5422), divided by the DG0 mass
// 0.166666666666667 |det J| (:5203-5231; mass_lumping_solver src/mesh_refiner.cpp:297-334), summed
// over the species. Bindings of benchmarks/pnp_refine_exact/main.cpp:260-266,318-338:
// mu = u_k + q_k phi, w = u_k, D = D_k. One thread per cell.
__global__ void __launch_bounds__(128)
entropy_indicator_kernel(int nc, const int4* __restrict__ cells, const double* __restrict__ xyz,
                         const double* __restrict__ uu, const double* __restrict__ diff,
                         const double* __restrict__ val, double* __restrict__ eta)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int4 cv = cells[c];
  const int v[4] = {cv.x, cv.y, cv.z, cv.w};
  double x[4][3], g[4][3], det;
  load_cell_xyz(xyz, cv, x);
  tet_geom(x, g, det);
  double phi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) phi[r] = uu[3 * (size_t)v[r]];
  double e = 0.0;
  for (int k = 1; k < 3; ++k) {
    double lw[4], D[4], gr[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const size_t o = 3 * (size_t)v[r] + k;
      lw[r] = uu[o]; D[r] = diff[o];
      const double mu = uu[o] + val[o] * phi[r];
      gr[0] += mu * g[r][0]; gr[1] += mu * g[r][1]; gr[2] += mu * g[r][2];
    }
    const double g2 = det * (gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
    double A = 0.0;
#pragma unroll 4
    for (int ip = 0; ip < 24; ++ip) {
      const double F0 = c_P24[ip][0] * D[0] + c_P24[ip][1] * D[1] + c_P24[ip][2] * D[2] + c_P24[ip][3] * D[3];
      const double F1 = c_P24[ip][0] * lw[0] + c_P24[ip][1] * lw[1] + c_P24[ip][2] * lw[2] + c_P24[ip][3] * lw[3];
      A += F0 * c_W24[ip] * exp(F1) * g2;
    }
    e += A / (0.166666666666667 * det);
  }
  eta[c] = e;
}

// Mesh_Refiner::mark_for_refinement (src/mesh_refiner.cpp:187-209): marker = eta > tolerance
__global__ void mark_cells_kernel(int nc, const double* __restrict__ eta, double tol,
                                  unsigned char* __restrict__ marker, unsigned long long* __restrict__ count)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool m = c < nc && eta[c] > tol;
  if (c < nc) marker[c] = m ? 1 : 0;
  const unsigned b = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

void entropy_indicator(Handle& h, double* d_eta)
{
  if (!h.have_coef) throw Error(ERROR_INPUT_PAR, "coefficients not set");
  entropy_indicator_kernel<<<cdiv(h.nc, 128), 128, 0, h.stream>>>(h.nc, reinterpret_cast<const int4*>(h.cells.p), h.xyz.p,
                                                                h.uu.p, h.diff.p, h.val.p, d_eta);
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}

// marker[c] = eta[c] > tol; with target_size > 0 the tolerance of mark_for_refinement_with_target_size
// (src/mesh_refiner.cpp:155-185): max(sorted_eta[index], tol), index = n - min(n, permissible cells)
// (the reference reads sorted_eta[n] when more cells are permissible than exist — clamped to n - 1 here)
long mark_cells(Handle& h, const double* d_eta, double tol, long target_size, unsigned char* d_marker, double* tol_used)
{
  cudaStream_t s = h.stream;
  const int nc = h.nc;
  if (target_size > 0) {
    DevBuf<double> sorted; sorted.alloc(nc, s);
    size_t bytes = 0;
    PNP_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_eta, sorted.p, nc, 0, 64, s));
    DevBuf<unsigned char> tmp; tmp.alloc(bytes, s);
    PNP_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_eta, sorted.p, nc, 0, 64, s));
    PNP_COUNT_LAUNCH();
    const long permissible = (long)nc > target_size ? 1 : target_size - (long)nc;
    long index = permissible > (long)nc ? (long)nc : (long)nc - permissible;
    if (index > (long)nc - 1) index = (long)nc - 1;
    double v = 0.0;
    PNP_CUDA(cudaMemcpyAsync(&v, sorted.p + index, sizeof(double), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (v > tol) tol = v;
  }
  DevBuf<unsigned long long> cnt; cnt.alloc(1, s); cnt.zero();
  mark_cells_kernel<<<cdiv(nc, 256), 256, 0, s>>>(nc, d_eta, tol, d_marker, cnt.p);
  PNP_COUNT_LAUNCH();
  unsigned long long hc = 0;
  PNP_CUDA(cudaMemcpyAsync(&hc, cnt.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if (tol_used) *tol_used = tol;
  return (long)hc;
}

// ---- damped Newton of the diode benchmark (benchmarks/pnp_diode/pnp_newton_solver.h:220-268) ----
// ranges of the previous solution, of the update and of their sum (growth limiting), and the trial
// iterate base + alpha * du of the backtracking loop, without leaving the device
__global__ void __launch_bounds__(256)
ranges_kernel(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ partial)
{
  __shared__ double sm[32];
  double v[6] = {1e300, -1e300, 1e300, -1e300, 1e300, -1e300};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double x = a[i], y = b[i], z = x + y;
    v[0] = fmin(v[0], x); v[1] = fmax(v[1], x); v[2] = fmin(v[2], y); v[3] = fmax(v[3], y); v[4] = fmin(v[4], z); v[5] = fmax(v[5], z);
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double r = (q & 1) ? v[q] : -v[q];                       // min as max of the negation (values of any sign)
    r = warp_max(r);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = sm[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sm[w]);
      partial[(size_t)blockIdx.x * 6 + q] = m;
    }
  }
}
__global__ void ranges_final_kernel(int nblk, const double* __restrict__ partial, double* __restrict__ out)
{
  const int q = threadIdx.x;
  if (q >= 6) return;
  double m = partial[q];
  for (int b = 1; b < nblk; ++b) m = fmax(m, partial[(size_t)b * 6 + q]);
  out[q] = m;                                                // entries 0, 2, 4 hold -min
}
__global__ void trial_kernel(size_t n, const double* __restrict__ base, const double* __restrict__ du, double alpha, double* __restrict__ uu)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) uu[i] = base[i] + alpha * du[i];
}

void update_ranges(Handle& h, const double* a, const double* b, size_t n, double* out6)
{
  cudaStream_t s = h.stream;
  const int nblk = (int)std::min<size_t>(592, (n + 255) / 256 ? (n + 255) / 256 : 1);
  h.red.alloc((size_t)nblk * 6 + 8, s);
  ranges_kernel<<<nblk, 256, 0, s>>>(n, a, b, h.red.p + 8); PNP_COUNT_LAUNCH();
  ranges_final_kernel<<<1, 32, 0, s>>>(nblk, h.red.p + 8, h.red.p); PNP_COUNT_LAUNCH();
  comm_allreduce(h, h.red.p, 6, true);
  PNP_CUDA(cudaMemcpyAsync(h.h_red, h.red.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  for (int q = 0; q < 6; ++q) out6[q] = (q & 1) ? h.h_red[q] : -h.h_red[q];
}

void set_trial(Handle& h, double alpha)
{
  const size_t n = owned_dofs(h);
  trial_kernel<<<592, 256, 0, h.stream>>>(n, h.uu_base.p, h.du.p, alpha, h.uu.p); PNP_COUNT_LAUNCH();
  h.rhs_current = false;
}

void error_norms(Handle& h, const double* d_exact, double* l2, double* h1)
{
  cudaStream_t s = h.stream;
  const int nblk = cdiv(h.nc, 256);
  h.red.alloc(2 * (size_t)nblk + 2, s);
  const bool dist = h.comm.active() && h.vowner.p;
  if (dist) halo_exchange(h, h.hier.lv[0]->halo, h.uu.p);
  error_cells_kernel<<<nblk, 256, 0, s>>>(h.nc, reinterpret_cast<const int4*>(h.cells.p), h.xyz.p, h.uu.p, d_exact,
                                          dist ? h.vowner.p : nullptr, h.comm.rank, h.red.p + 2);
  PNP_COUNT_LAUNCH();
  finalize_sum2_kernel<<<1, 1024, 0, s>>>(nblk, h.red.p + 2, h.red.p); PNP_COUNT_LAUNCH();
  comm_allreduce(h, h.red.p, 2, false);          // all ranks return the global norms
  PNP_CUDA(cudaMemcpyAsync(h.h_red, h.red.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  *l2 = sqrt(h.h_red[0]);
  *h1 = sqrt(h.h_red[0] + h.h_red[1]);       // Error::compute_h1_error: sqrt(l2^2 + semi_h1^2)
}

void total_charge(Handle& h, double* d_out)
{
  total_charge_kernel<<<cdiv(h.nv, 256), 256, 0, h.stream>>>(h.nv, h.uu.p, h.fixed.p, h.val.p, d_out);
  PNP_COUNT_LAUNCH();
}

void assemble_residual(Handle& h, DevBuf<double>& out, double* l2, double* linf)
{
  cudaStream_t s = h.stream;
  if (!h.have_coef) throw Error(ERROR_INPUT_PAR, "coefficients not set");
  const int4* cells = reinterpret_cast<const int4*>(h.cells.p);
  const size_t need = (size_t)h.nc * 12;
  if (h.cellrec.n < need) h.cellrec.alloc(need, s);
  out.alloc(3 * (size_t)h.nv, s);
  // 4 CTAs/SM (128 registers): measured 26.7 ms per residual at 256^3 vs 32.0 at 2 CTAs/SM
  cell_residual_kernel<4><<<cdiv(h.nc, 128), 128, 0, s>>>(h.nc, cells, h.xyz.p, h.uu.p, h.eps.p, h.fixed.p, h.diff.p,
                                                         h.val.p, h.reac.p, h.poisson_scale, h.cvslot.p, h.cellrec.p);
  PNP_COUNT_LAUNCH();
  const int nblk = cdiv(h.nv, 256);
  h.red.alloc(2 * (size_t)nblk + 2, s);
  residual_gather_kernel<<<nblk, 256, 0, s>>>(h.nv, h.v2c_ptr.p, h.cellrec.p, h.bcflag.p, out.p, h.red.p + 2);
  PNP_COUNT_LAUNCH();
  finalize_norm_kernel<<<1, 1024, 0, s>>>(nblk, h.red.p + 2, h.red.p);
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
  h.rhs_current = (&out == &h.rhs);
  if (l2 || linf) {
    comm_allreduce(h, h.red.p, 1, false);        // sum of squares over all ranks (ghost rows are zero)
    comm_allreduce(h, h.red.p + 1, 1, true);     // max
    PNP_CUDA(cudaMemcpyAsync(h.h_red, h.red.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (l2) *l2 = sqrt(h.h_red[0]);
    if (linf) *linf = h.h_red[1];
  }
}

void assemble_system(Handle& h)
{
  cudaStream_t s = h.stream;
  if (!h.have_coef) throw Error(ERROR_INPUT_PAR, "coefficients not set");
  const int4* cells = reinterpret_cast<const int4*>(h.cells.p);
  BsrT& A = h.hier.lv[0]->A;
  A.packed7 = false;                    // the values are about to change: the 7-value copy of the last solve is stale
  static const bool timing = getenv("PNPB200_ASM_TIMING") != nullptr;
  auto mark = [&](int k) { if (timing) cudaEventRecord(h.ev[4 + k], s); };
  mark(0);
  // rhs first (re-uses the record buffer); the residual the previous Newton step ended with IS this
  // step's rhs (the reference assembles it again, src/pde.cpp:546), so it is only rebuilt if stale
  if (!h.rhs_current) assemble_residual(h, h.rhs, nullptr, nullptr);
  mark(1);
  const size_t need = h.use_eafe ? 4 * h.n_pair_records : (size_t)h.nc * REC_FULL;
  if (h.share_scratch < 0) {
    // can the device hold the records AND the flexible Krylov basis (31 packed + 30 padded vectors, bsr.dat's restart)?
    const char* e = getenv("PNPB200_SHARE_SCRATCH");
    if (e) h.share_scratch = atoi(e) ? 1 : 0;
    else {
      size_t fr = 0, tot = 0;
      PNP_CUDA(cudaMemGetInfo(&fr, &tot));
      const double held = h.cellrec.n * 8.0;                                  // already ours (pool memory does not count as free)
      const double want = need * 8.0 * 1.0625 + (31.0 * 3 + 30.0 * VS) * 8.0 * (double)h.nv * 1.0625 + 45.0 * (double)A.nnzb;   // + hierarchy and setup workspace
      h.share_scratch = (double)fr + held < want + 6e9 ? 1 : 0;
    }
  }
  if (h.cellrec.n < need) h.cellrec.alloc(need, s);
  if (h.use_eafe) {
    // exp-bound: 3 CTAs/SM (136 registers) is the measured optimum (23 ms; 25 at 4, 35 at 6)
    cell_jacobian_kernel<false, 3><<<cdiv(h.nc, 128), 128, 0, s>>>(h.nc, cells, h.xyz.p, h.uu.p, h.diff.p, h.val.p, h.poisson_scale, h.conc_cutoff, h.cpslot.p, h.cellrec.p);
    PNP_COUNT_LAUNCH();
    mark(2);
    h.eafe_tab.alloc(8 * (size_t)h.nv, s);
    eafe_vertex_table_kernel<<<cdiv(h.nv, 256), 256, 0, s>>>(h.nv, h.uu.p, h.diff.p, h.q_eafe[1], h.q_eafe[2], h.eafe_tab.p);
    PNP_COUNT_LAUNCH();
    assemble_pairs_kernel<PNPB200_ASM_CTAS><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(
        A.n, A.rd.p, A.ja.p, h.pblk.p, h.cellrec.p, h.k00.p, h.omega.p, h.eafe_tab.p, h.bcflag.p, A.val.p);
    PNP_COUNT_LAUNCH();
  } else {
    cell_jacobian_kernel<true, 2><<<cdiv(h.nc, 128), 128, 0, s>>>(h.nc, cells, h.xyz.p, h.uu.p, h.diff.p, h.val.p, h.poisson_scale, h.conc_cutoff, nullptr, h.cellrec.p);
    PNP_COUNT_LAUNCH();
    mark(2);
    assemble_blocks_kernel<false, 4><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(
        A.n, A.rd.p, A.ja.p, h.v2c_ptr.p, h.v2c_idx.p, cells, h.cellrec.p, h.k00.p, h.omega.p,
        h.uu.p, h.diff.p, h.q_eafe[1], h.q_eafe[2], h.bcflag.p, A.val.p);
    PNP_COUNT_LAUNCH();
  }
  mark(3);
  zero_row_fix_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, A.rs.p, A.dslot.p, A.val.p);
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
  if (timing) {
    cudaEventRecord(h.ev[3], s); cudaEventSynchronize(h.ev[3]);
    float a, b, c, d;
    cudaEventElapsedTime(&a, h.ev[4], h.ev[5]); cudaEventElapsedTime(&b, h.ev[5], h.ev[6]);
    cudaEventElapsedTime(&c, h.ev[6], h.ev[7]); cudaEventElapsedTime(&d, h.ev[7], h.ev[3]);
    fprintf(stderr, "[pnpb200 asm] residual %.2f ms | jacobian cells %.2f | blocks %.2f | zero-row %.2f\n", a, b, c, d);
  }
}

} // namespace pnpb200
