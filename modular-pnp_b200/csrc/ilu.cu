// pnp-b200 — block ILU(k) for BSR(3) on the device: stands in for FASP's
// fasp_solver_dbsr_krylov_ilu [EXT] as benchmarks/pnp_diode/linear_pnp.cpp:103-109 calls it with
// benchmarks/pnp_diode/bsr.dat:34-38 (ILU_type 2 = ILU(k), ILU_lfil 2).
//   symbolic : level-of-fill pattern on the host, once per sparsity pattern (the diode meshes have
//              10^3 - 10^5 vertices; the pattern only changes when the mesh does);
//   numeric  : IKJ block factorisation on the device, every Newton step (L unit lower, inverses of
//              the diagonal blocks of U stored);
//   apply    : z = U^{-1} L^{-1} r, forward and backward substitution.
// All three device phases are SYNCHRONISATION-FREE sparse triangular sweeps (Liu et al.): a warp
// takes the next row from an atomic ticket counter, spins on the "done" flags of the rows it
// depends on and publishes its own. Tickets are handed out in dependency order (ascending rows for
// the factorisation and the forward solve, descending for the backward solve), so every row a warp
// waits for is held by a warp that is already running: no level analysis, no kernel per level, no
// deadlock. Rows are in vertex (row-id) order like the reference's; vectors of the solver's sweep
// space are read / written through the position map.
// Checked against oracle/ilu_restated.cpp: factors 1e-12, apply 1e-12, Krylov count.
#include "internal.hpp"
#include <algorithm>

namespace pnpb200 {

namespace {

constexpr int ILU_WARPS = 4;        // warps per CTA

__device__ __forceinline__ int ld_flag(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// position of factor entry (i, j) in A's physical storage, -1 for fill
__global__ void ilu_src_kernel(int n, const int* __restrict__ fia, const int* __restrict__ fja, const int4* __restrict__ rd,
                               const int* __restrict__ ja, int* __restrict__ src)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 d = rd[i];
  for (int p = fia[i]; p < fia[i + 1]; ++p) {
    const int j = fja[p];
    int b = -1;
    for (int k = 0; k < d.y; ++k) if (ja[d.x + k] == j) { b = d.x + k; break; }
    src[p] = b;
  }
}

__device__ __forceinline__ void inv3_dev(const double* a, double* o)
{
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double id = 1.0 / (a[0] * c0 + a[1] * c1 + a[2] * c2);
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// numeric factorisation. Shared memory per warp: the row's blocks w[maxlen][9] and its columns.
__global__ void __launch_bounds__(32 * ILU_WARPS)
ilu_numeric_kernel(int n, int maxlen, int epoch, const int* __restrict__ fia, const int* __restrict__ fja,
                   const int* __restrict__ fdiag, const int* __restrict__ src, const int4* __restrict__ rd,
                   const double* __restrict__ aval, double* fval, double* dinv, int* flag, int* ticket)
{
  extern __shared__ double ilu_smem[];
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* w = ilu_smem + (size_t)wp * maxlen * 9;
  int* cols = reinterpret_cast<int*>(ilu_smem + (size_t)ILU_WARPS * maxlen * 9) + wp * maxlen;
  while (true) {
    int i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) return;
    const int p0 = fia[i], len = fia[i + 1] - p0, dg = fdiag[i] - p0;
    const int4 d = rd[i];
    for (int q = lane; q < len; q += 32) {
      cols[q] = fja[p0 + q];
      const int b = src[p0 + q];
      if (b >= 0) {
        int st;
        const size_t o = bsrt_off(d.x, d.y, d.z, b - d.x, st);
#pragma unroll
        for (int t = 0; t < 9; ++t) w[q * 9 + t] = aval[o + (size_t)t * st];
      } else {
#pragma unroll
        for (int t = 0; t < 9; ++t) w[q * 9 + t] = 0.0;
      }
    }
    __syncwarp();
    for (int q = 0; q < dg; ++q) {                         // eliminate with row k = cols[q], ascending
      const int k = cols[q];
      if (lane == 0) while (ld_flag(flag + k) != epoch) { }
      __syncwarp();
      __threadfence();
      // L_ik = w_q * U_kk^{-1}: lane (r, c) of the first 9 forms one entry
      double lik = 0.0;
      if (lane < 9) {
        const int r = lane / 3, c = lane % 3;
        // rows of other warps (possibly other SMs) written during this launch: read around L1 (ld.cg)
        const double* dk = dinv + 9 * (size_t)k;
        lik = w[q * 9 + 3 * r] * __ldcg(dk + c) + w[q * 9 + 3 * r + 1] * __ldcg(dk + 3 + c) + w[q * 9 + 3 * r + 2] * __ldcg(dk + 6 + c);
      }
      __syncwarp();
      if (lane < 9) w[q * 9 + lane] = lik;
      __syncwarp();
      // w_j -= L_ik U_kj for the entries j > k of row k that row i holds; one lane per entry of row k
      const int u0 = fdiag[k] + 1, u1 = fia[k + 1];
      for (int p = u0 + lane; p < u1; p += 32) {
        const int j = fja[p];
        int lo = q + 1, hi = len - 1, pos = -1;              // columns ascending: binary search behind q
        while (lo <= hi) { const int mid = (lo + hi) >> 1; const int c = cols[mid]; if (c == j) { pos = mid; break; } if (c < j) lo = mid + 1; else hi = mid - 1; }
        if (pos < 0) continue;
        double u[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) u[e] = __ldcg(fval + 9 * (size_t)p + e);
        double* t = w + pos * 9;
        const double* l = w + q * 9;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) t[3 * r + c] -= l[3 * r] * u[c] + l[3 * r + 1] * u[3 + c] + l[3 * r + 2] * u[6 + c];
      }
      __syncwarp();
    }
    if (lane == 0) inv3_dev(w + dg * 9, dinv + 9 * (size_t)i);
    for (int e = lane; e < 9 * len; e += 32) fval[9 * (size_t)p0 + e] = w[e];
    __threadfence();
    __syncwarp();
    if (lane == 0) { *reinterpret_cast<volatile int*>(flag + i) = epoch; }
    __syncwarp();
  }
}

// forward substitution y = L^{-1} r (unit diagonal); r in the caller's vector layout: r[rs * pos[i] + c]
// (pos == nullptr: row-id order)
__global__ void __launch_bounds__(32 * ILU_WARPS)
ilu_forward_kernel(int n, int epoch, const int* __restrict__ fia, const int* __restrict__ fja, const int* __restrict__ fdiag,
                   const double* __restrict__ fval, const double* __restrict__ r, int rs, const int* __restrict__ pos,
                   double* y, int* flag, int* ticket)
{
  const int lane = threadIdx.x & 31;
  while (true) {
    int i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) return;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int p = fia[i] + lane; p < fdiag[i]; p += 32) {
      const int k = fja[p];
      const double* a = fval + 9 * (size_t)p;
      const double a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6], a7 = a[7], a8 = a[8];
      while (ld_flag(flag + k) != epoch) { }
      __threadfence();
      const volatile double* yk = y + 3 * (size_t)k;
      const double y0 = yk[0], y1 = yk[1], y2 = yk[2];
      t0 += a0 * y0 + a1 * y1 + a2 * y2; t1 += a3 * y0 + a4 * y1 + a5 * y2; t2 += a6 * y0 + a7 * y1 + a8 * y2;
    }
    t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (lane == 0) {
      const size_t o = (size_t)rs * (size_t)(pos ? pos[i] : i);
      y[3 * (size_t)i] = r[o] - t0; y[3 * (size_t)i + 1] = r[o + 1] - t1; y[3 * (size_t)i + 2] = r[o + 2] - t2;
      __threadfence();
      *reinterpret_cast<volatile int*>(flag + i) = epoch;
    }
    __syncwarp();
  }
}

// backward substitution x = U^{-1} y, rows descending; z[zs * pos[i] + c] = x_i (pad entry zeroed when zs == 4)
__global__ void __launch_bounds__(32 * ILU_WARPS)
ilu_backward_kernel(int n, int epoch, const int* __restrict__ fia, const int* __restrict__ fja, const int* __restrict__ fdiag,
                    const double* __restrict__ fval, const double* __restrict__ dinv, const double* __restrict__ y,
                    double* xw, double* z, int zs, const int* __restrict__ pos, int* flag, int* ticket)
{
  const int lane = threadIdx.x & 31;
  while (true) {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, 1);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= n) return;
    const int i = n - 1 - tk;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int p = fdiag[i] + 1 + lane; p < fia[i + 1]; p += 32) {
      const int j = fja[p];
      const double* a = fval + 9 * (size_t)p;
      const double a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6], a7 = a[7], a8 = a[8];
      while (ld_flag(flag + j) != epoch) { }
      __threadfence();
      const volatile double* xj = xw + 3 * (size_t)j;
      const double x0 = xj[0], x1 = xj[1], x2 = xj[2];
      t0 += a0 * x0 + a1 * x1 + a2 * x2; t1 += a3 * x0 + a4 * x1 + a5 * x2; t2 += a6 * x0 + a7 * x1 + a8 * x2;
    }
    t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (lane == 0) {
      const double s0 = y[3 * (size_t)i] - t0, s1 = y[3 * (size_t)i + 1] - t1, s2 = y[3 * (size_t)i + 2] - t2;
      const double* d = dinv + 9 * (size_t)i;
      const double x0 = d[0] * s0 + d[1] * s1 + d[2] * s2, x1 = d[3] * s0 + d[4] * s1 + d[5] * s2, x2 = d[6] * s0 + d[7] * s1 + d[8] * s2;
      xw[3 * (size_t)i] = x0; xw[3 * (size_t)i + 1] = x1; xw[3 * (size_t)i + 2] = x2;
      const size_t o = (size_t)zs * (size_t)(pos ? pos[i] : i);
      z[o] = x0; z[o + 1] = x1; z[o + 2] = x2;
      if (zs == 4) z[o + 3] = 0.0;
      __threadfence();
      *reinterpret_cast<volatile int*>(flag + i) = epoch;
    }
    __syncwarp();
  }
}

int ilu_grid(int n)
{
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int want = (n + ILU_WARPS - 1) / ILU_WARPS;
  const int cap = sms * 4;                                   // every CTA of the grid must be resident (spin waits)
  return want < cap ? (want > 0 ? want : 1) : cap;
}

} // namespace

// symbolic factorisation on the host (level of fill <= lfil) + the map into A's storage; kept until the pattern changes
void ilu_symbolic(Handle& h, int lfil)
{
  cudaStream_t s = h.stream;
  IluFactor& F = h.ilu;
  const BsrT& A = h.hier.lv[0]->A;
  const int n = A.n;
  std::vector<int> IA = A.ia.to_host(n + 1);
  DevBuf<int> ja_nat; ja_nat.alloc(A.nnzb, s);
  bsrt_export_ja(A, ja_nat.p, s);
  std::vector<int> JA = ja_nat.to_host(A.nnzb);
  std::vector<int> fIA(n + 1, 0), fJA, fdiag(n, -1), lev(n, -1), cols;
  std::vector<std::vector<std::pair<int, int>>> rows(n);
  size_t maxlen = 0;
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
    for (int c : cols) { if (c == i) fdiag[i] = (int)fJA.size(); rows[i].emplace_back(c, lev[c]); lev[c] = -1; fJA.push_back(c); }
    if (fJA.size() > 0x7fffffffull / 9) throw Error(ERROR_ALLOC_MEM, "ILU(k): factor pattern exceeds 2^31 / 9 blocks");
    fIA[i + 1] = (int)fJA.size();
    maxlen = std::max(maxlen, cols.size());
  }
  F.n = n; F.nnz = fIA[n]; F.maxlen = (int)maxlen; F.lfil = lfil;
  F.ia.alloc(n + 1, s); F.ja.alloc(F.nnz, s); F.diag.alloc(n, s); F.src.alloc(F.nnz, s);
  F.ia.upload(fIA.data(), n + 1); F.ja.upload(fJA.data(), F.nnz); F.diag.upload(fdiag.data(), n);
  F.val.alloc(9 * (size_t)F.nnz, s); F.dinv.alloc(9 * (size_t)n, s); F.y.alloc(3 * (size_t)n, s); F.x.alloc(3 * (size_t)n, s);
  F.flag.alloc(3 * (size_t)n, s); F.flag.zero(); F.ticket.alloc(4, s);
  F.epoch = 0;
  ilu_src_kernel<<<cdiv(n, 128), 128, 0, s>>>(n, F.ia.p, F.ja.p, A.rd.p, A.ja.p, F.src.p); PNP_COUNT_LAUNCH();
  PNP_CUDA(cudaStreamSynchronize(s));       // host staging vectors go out of scope
  F.symbolic_valid = true;
}

void ilu_setup(Handle& h, int lfil)
{
  cudaStream_t s = h.stream;
  IluFactor& F = h.ilu;
  const BsrT& A = h.hier.lv[0]->A;
  if (lfil < 0) lfil = 0;
  if (!F.symbolic_valid || F.n != A.n || F.lfil != lfil || F.pattern_nnzb != A.nnzb) { ilu_symbolic(h, lfil); F.pattern_nnzb = A.nnzb; }
  const size_t smem = (size_t)ILU_WARPS * F.maxlen * (9 * sizeof(double) + sizeof(int));
  if (smem > 200 * 1024) throw Error(ERROR_ALLOC_MEM, "ILU(k): a factor row is too long for the shared-memory row buffer (lower ILU_lfil)");
  static bool configured[64] = {false};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    PNP_CUDA(cudaFuncSetAttribute(ilu_numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  ++F.epoch;
  PNP_CUDA(cudaMemsetAsync(F.ticket.p, 0, 4 * sizeof(int), s));
  // resident CTAs are bounded by the shared-memory row buffers
  int grid = ilu_grid(F.n);
  const int per_sm = (int)std::max<size_t>(1, (200 * 1024) / std::max<size_t>(smem, 1));
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  grid = std::min(grid, sms * std::min(per_sm, 4));
  ilu_numeric_kernel<<<grid, 32 * ILU_WARPS, smem, s>>>(F.n, F.maxlen, F.epoch, F.ia.p, F.ja.p, F.diag.p, F.src.p, A.rd.p, A.val.p,
                                                      F.val.p, F.dinv.p, F.flag.p, F.ticket.p);
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}

// z = U^{-1} L^{-1} r. sweep_space: r / z are solver vectors (rs / zs doubles per row, indexed by sweep position)
void ilu_apply(Handle& h, const double* r, int rs, double* z, int zs, bool sweep_space)
{
  cudaStream_t s = h.stream;
  IluFactor& F = h.ilu;
  const BsrT& A = h.hier.lv[0]->A;
  if (!F.symbolic_valid) throw Error(ERROR_INPUT_PAR, "ILU factors not set up");
  const int* pos = sweep_space ? A.pos.p : nullptr;
  ++F.epoch;
  PNP_CUDA(cudaMemsetAsync(F.ticket.p, 0, 4 * sizeof(int), s));
  const int grid = ilu_grid(F.n);
  ilu_forward_kernel<<<grid, 32 * ILU_WARPS, 0, s>>>(F.n, F.epoch, F.ia.p, F.ja.p, F.diag.p, F.val.p, r, rs, pos, F.y.p,
                                                    F.flag.p + F.n, F.ticket.p + 1);
  PNP_COUNT_LAUNCH();
  ilu_backward_kernel<<<grid, 32 * ILU_WARPS, 0, s>>>(F.n, F.epoch, F.ia.p, F.ja.p, F.diag.p, F.val.p, F.dinv.p, F.y.p, F.x.p, z, zs, pos,
                                                     F.flag.p + 2 * (size_t)F.n, F.ticket.p + 2);
  PNP_COUNT_LAUNCH();
}

} // namespace pnpb200
