// pnp-b200 — BSR(3) kernels on the BSR-T layout:
//   fasp_blas_dbsr_mxv / aAxpy   [EXT FASP]      -> bsrt_spmv / bsrt_residual
//   fasp_dbsr_getdiaginv         [EXT FASP]      -> bsrt_diaginv
//   fasp_smoother_dbsr_gs_*      [EXT FASP]      -> multicolour block Gauss-Seidel
// A half-warp (16 lanes) owns one block row; lane k owns block k of the row, so each of the 9
// component loads is unit-stride across lanes. Rows are STORED colour by colour (rs[] = physical
// start of a row, rowid[] = row at a physical position, pia[] = CSR offsets in physical order): one
// Gauss-Seidel colour is one contiguous piece of ja/val; SpMV walks the rows in natural order. All
// kernels here are HBM-bound streaming kernels.
#include "internal.hpp"
#include <cub/cub.cuh>

namespace pnpb200 {

namespace {

// One block row per half-warp, lane k = block k of the row.
//   MODE 0: y = A x          MODE 1: y = b - A x      (rows in natural order; idx = packed
//           (physical start, length) per row, one 8-byte load)
//   MODE 2: x_row = D^{-1}(b - sum_{j != row} A_row,j x_j), y == x: one Gauss-Seidel colour.
//           rows[] = rowid + p0 and idx = pia + p0 are indexed by physical position, so row id,
//           start and length come from one level of loads.
// The kernels are latency/occupancy bound (ncu: long-scoreboard stalls, 66-68 % achieved warps):
// the colour kernel is compiled for 8 resident CTAs per SM (32 registers, no spills; measured
// 4.7 ms vs 5.3 ms per fine-level sweep at 256^3), the natural-order passes for 6 (they spill at 8).
// Tried and measured slower: ld.global.cs / L2 evict_first+evict_last hints on the matrix loads
// (6.3-7.2 ms per sweep), requesting b and D^{-1} before the row's blocks (6.2 vs 5.5 ms: the extra
// registers cost more occupancy than the overlap gains).
template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB)
bsrt_row_kernel(int nrows, const int* __restrict__ rows, const int* __restrict__ idx,
                const int* __restrict__ ja, const double* __restrict__ val,
                const double* __restrict__ dinv, const double* __restrict__ b, const double* x, double* y)
{
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  int row = -1, s = 0, len = 0;
  if (t < nrows) {
    if (MODE == 2) { row = rows[t]; s = idx[t]; len = idx[t + 1] - s; }
    else { row = t; const int2 sl = reinterpret_cast<const int2*>(idx)[t]; s = sl.x; len = sl.y; }
  }
  const double* v = val + 9 * (size_t)s;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int k = lane; k < len; k += 16) {
    const int jc = ja[s + k];
    if (MODE == 2 && jc == row) continue;
    const size_t j = 3 * (size_t)jc;
    const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2];
    a0 += v[k] * x0 + v[len + k] * x1 + v[2 * len + k] * x2;
    a1 += v[3 * len + k] * x0 + v[4 * len + k] * x1 + v[5 * len + k] * x2;
    a2 += v[6 * len + k] * x0 + v[7 * len + k] * x1 + v[8 * len + k] * x2;
  }
  a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
  if (row >= 0 && lane < 3) {
    const size_t o = 3 * (size_t)row;
    if (MODE == 0) y[o + lane] = lane == 0 ? a0 : lane == 1 ? a1 : a2;
    else if (MODE == 1) y[o + lane] = b[o + lane] - (lane == 0 ? a0 : lane == 1 ? a1 : a2);
    else {
      const double r0 = b[o] - a0, r1 = b[o + 1] - a1, r2 = b[o + 2] - a2;
      const double* d = dinv + 9 * (size_t)row + 3 * lane;
      y[o + lane] = d[0] * r0 + d[1] * r1 + d[2] * r2;
    }
  }
}

// Small levels (<= SMALL_LEVEL_ROWS rows): the whole multicolour sweep in ONE launch of one CTA,
// colours separated by __syncthreads() instead of kernel boundaries (the coarsest levels of the
// K-cycle are visited 8-16 times per Krylov iteration and are pure launch latency). Measured at
// 256^3: threshold 600 -> -1.3 % Krylov time, threshold 6144 -> +2.5 % (one SM is too slow for a
// 3.5 k-row level), so only the last one or two levels take this path.
constexpr int SMALL_LEVEL_ROWS = 512;
__global__ void __launch_bounds__(1024, 1)
mcgs_small_kernel(int ngroups, int reverse, const int* __restrict__ gptr, const int* __restrict__ rowid,
                  const int* __restrict__ pia, const int* __restrict__ ja, const double* __restrict__ val,
                  const double* __restrict__ dinv, const double* __restrict__ b, double* x)
{
  const int hw = threadIdx.x >> 4, nhw = blockDim.x >> 4, lane = threadIdx.x & 15;
  for (int gg = 0; gg < ngroups; ++gg) {
    const int g = reverse ? ngroups - 1 - gg : gg;
    const int p0 = gptr[g], p1 = gptr[g + 1];
    for (int base = p0; base < p1; base += nhw) {        // uniform trip count: shuffles stay converged
      const int p = base + hw;
      int row = -1, s = 0, len = 0;
      if (p < p1) { row = rowid[p]; s = pia[p]; len = pia[p + 1] - s; }
      const double* v = val + 9 * (size_t)s;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int k = lane; k < len; k += 16) {
        const int jc = ja[s + k];
        if (jc == row) continue;
        const size_t j = 3 * (size_t)jc;
        const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2];
        a0 += v[k] * x0 + v[len + k] * x1 + v[2 * len + k] * x2;
        a1 += v[3 * len + k] * x0 + v[4 * len + k] * x1 + v[5 * len + k] * x2;
        a2 += v[6 * len + k] * x0 + v[7 * len + k] * x1 + v[8 * len + k] * x2;
      }
      a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
      if (row >= 0 && lane < 3) {
        const size_t o = 3 * (size_t)row;
        const double r0 = b[o] - a0, r1 = b[o + 1] - a1, r2 = b[o + 2] - a2;
        const double* d = dinv + 9 * (size_t)row + 3 * lane;
        x[o + lane] = d[0] * r0 + d[1] * r1 + d[2] * r2;
      }
    }
    __syncthreads();
  }
}

__global__ void diaginv_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                               const int* __restrict__ dslot, const double* __restrict__ val,
                               double* __restrict__ dinv)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = rs[i], len = ia[i + 1] - ia[i], d = dslot[i] - s;
  double a[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) a[t] = val[9 * (size_t)s + (size_t)t * len + d];
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  const double id = 1.0 / det;
  double* o = dinv + 9 * (size_t)i;
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// FASP layout (natural block order, row-major 3x3 per block) <-> BSR-T (physical)
template <bool TO_T>
__global__ void relayout_kernel(int nnzb, const int* __restrict__ ia, const int* __restrict__ rs,
                                const int* __restrict__ brow, const double* __restrict__ in, double* __restrict__ out)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * (size_t)nnzb) return;
  const int b = (int)(e / 9), t = (int)(e % 9);          // b = physical block
  const int i = brow[b], s = rs[i], len = ia[i + 1] - ia[i], k = b - s;
  const size_t et = 9 * (size_t)s + (size_t)t * len + k;
  const size_t ef = 9 * ((size_t)ia[i] + k) + t;
  if (TO_T) out[et] = in[ef]; else out[ef] = in[et];
}

// natural <-> physical column arrays (row granularity)
template <bool TO_PHYS>
__global__ void permute_ja_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                  const int* __restrict__ in, int* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = ia[i], len = ia[i + 1] - a, s = rs[i];
  for (int k = 0; k < len; ++k) { if (TO_PHYS) out[s + k] = in[a + k]; else out[a + k] = in[s + k]; }
}

__global__ void row_tables_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                  const int* __restrict__ ja, int* __restrict__ brow, int* __restrict__ dslot)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = rs[i], len = ia[i + 1] - ia[i];
  int d = -1;
  for (int p = s; p < s + len; ++p) { brow[p] = i; if (ja[p] == i) d = p; }
  dslot[i] = d;
}

__global__ void tslot_kernel(int nnzb, const int* __restrict__ ia, const int* __restrict__ rs,
                             const int* __restrict__ ja, const int* __restrict__ brow, int* __restrict__ tslot)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int i = brow[b], j = ja[b];
  int lo = rs[j], hi = rs[j] + (ia[j + 1] - ia[j]) - 1, res = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, c = ja[mid];
    if (c == i) { res = mid; break; }
    if (c < i) lo = mid + 1; else hi = mid - 1;
  }
  tslot[b] = res;
}

// physical row starts: rs[rowid[p]] = exclusive prefix over p of len(rowid[p])
__global__ void rowlen_by_position_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rowid,
                                          int* __restrict__ len)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { const int i = rowid[p]; len[p] = ia[i + 1] - ia[i]; }
  if (p == n) len[n] = 0;
}
__global__ void scatter_rs_kernel(int n, const int* __restrict__ rowid, const int* __restrict__ start,
                                  int* __restrict__ rs, int2* __restrict__ rsl)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { const int i = rowid[p]; rs[i] = start[p]; rsl[i] = make_int2(start[p], start[p + 1] - start[p]); }
}

// ---- Jones-Plassmann colouring with hashed priorities (two kernels per round so that the
// result does not depend on thread timing); natural CSR graph ------------------------------------
__global__ void color_select_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                    const int* __restrict__ color, unsigned char* __restrict__ cand,
                                    int* __restrict__ remaining)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cand[i] = 0;
  if (color[i] >= 0) return;
  const unsigned pi = hash32((unsigned)i);
  bool is_max = true;
  for (int p = ia[i]; p < ia[i + 1]; ++p) {
    const int j = ja[p];
    if (j == i || color[j] >= 0) continue;
    const unsigned pj = hash32((unsigned)j);
    if (pj > pi || (pj == pi && j > i)) { is_max = false; break; }
  }
  if (is_max) cand[i] = 1; else atomicAdd(remaining, 1);
}

__global__ void color_assign_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                    int* __restrict__ color, const unsigned char* __restrict__ cand,
                                    int* __restrict__ overflow)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !cand[i]) return;
  unsigned long long used = 0ull;
  for (int p = ia[i]; p < ia[i + 1]; ++p) {
    const int j = ja[p];
    if (j == i) continue;
    const int cj = color[j];            // neighbours of a candidate are never candidates
    if (cj >= 0) used |= 1ull << cj;
  }
  const int c = __ffsll(~used) - 1;     // smallest free colour
  if (c < 0 || c >= 63) { *overflow = 1; color[i] = 62; } else color[i] = c;
}

__global__ void max_color_kernel(int n, const int* __restrict__ color, int* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = i < n ? color[i] : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, c, o); c = t > c ? t : c; }
  if ((threadIdx.x & 31) == 0 && c >= 0) atomicMax(out, c);
}

// Culberson's iterated greedy: rebuild the colouring class by class (members of a class are
// independent, so a class can be recoloured in parallel); never uses more colours than before
__global__ void recolor_class_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                     const int* __restrict__ oldc, int cls, int* __restrict__ newc)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || oldc[i] != cls) return;
  unsigned long long used = 0ull;
  for (int p = ia[i]; p < ia[i + 1]; ++p) {
    const int j = ja[p];
    if (j == i) continue;
    const int cj = newc[j];
    if (cj >= 0) used |= 1ull << cj;
  }
  newc[i] = __ffsll(~used) - 1;
}

// sort key = (tile of the row, colour): a sweep visits tile after tile and, inside a tile, colour
// after colour, so the x-window a tile gathers from stays in L2 across its colours
__global__ void group_key_kernel(int n, int n_own, int ghost_key, int tile_rows, int* __restrict__ color, int* __restrict__ hist)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int key = i >= n_own ? ghost_key : (i / tile_rows) * 64 + color[i];   // ghost rows: last group, never swept
  color[i] = key;
  atomicAdd(&hist[key], 1);
}

__global__ void iota_kernel(int n, int* __restrict__ a)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

} // namespace

void bsrt_spmv(const BsrT& A, const double* x, double* y, cudaStream_t s)
{
  bsrt_row_kernel<0, 6><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(A.n, nullptr, reinterpret_cast<const int*>(A.rsl.p), A.ja.p, A.val.p,
                                                                   nullptr, nullptr, x, y);
  PNP_COUNT_LAUNCH();
}

void bsrt_residual(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s)
{
  bsrt_row_kernel<1, 6><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(A.n, nullptr, reinterpret_cast<const int*>(A.rsl.p), A.ja.p, A.val.p,
                                                                   nullptr, b, x, r);
  PNP_COUNT_LAUNCH();
}

void bsrt_diaginv(const BsrT& A, double* dinv, cudaStream_t s)
{
  diaginv_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, A.rs.p, A.dslot.p, A.val.p, dinv);
  PNP_COUNT_LAUNCH();
}

void bsrt_from_fasp(const double* d_val_fasp, BsrT& A, cudaStream_t s)
{
  relayout_kernel<true><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.brow.p, d_val_fasp, A.val.p);
  PNP_COUNT_LAUNCH();
}

void bsrt_to_fasp(const BsrT& A, double* d_val_fasp, cudaStream_t s)
{
  relayout_kernel<false><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.brow.p, A.val.p, d_val_fasp);
  PNP_COUNT_LAUNCH();
}

void bsrt_export_ja(const BsrT& A, int* d_ja_nat, cudaStream_t s)
{
  permute_ja_kernel<false><<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, A.rs.p, A.ja.p, d_ja_nat);
  PNP_COUNT_LAUNCH();
}

void mcgs_sweep(const Level& L, const double* b, double* x, bool reverse, cudaStream_t s)
{
  const BsrT& A = L.A;
  if (A.n <= SMALL_LEVEL_ROWS && L.ngroups > 0) {
    mcgs_small_kernel<<<1, 1024, 0, s>>>(L.ngroups, reverse ? 1 : 0, L.d_color_ptr.p, A.rowid.p, A.pia.p, A.ja.p, A.val.p,
                                        L.dinv.p, b, x);
    PNP_COUNT_LAUNCH();
    return;
  }
  for (int cc = 0; cc < L.ngroups; ++cc) {
    const int c = reverse ? L.ngroups - 1 - cc : cc;
    const int cnt = L.color_ptr[c + 1] - L.color_ptr[c];
    if (cnt == 0) continue;
    bsrt_row_kernel<2, 8><<<cdiv((size_t)cnt * 16, 256), 256, 0, s>>>(cnt, A.rowid.p + L.color_ptr[c], A.pia.p + L.color_ptr[c],
                                                                     A.ja.p, A.val.p, L.dinv.p, b, x, x);
    PNP_COUNT_LAUNCH();
  }
}

// Colour the natural graph (A.ia, ja_nat), then lay the matrix out colour by colour.
// `ws` (optional) supplies the temporaries during AMG setup; at create time they come from the pool.
void color_and_layout(Level& L, const int* ja_nat, Workspace* ws, cudaStream_t s)
{
  BsrT& A = L.A;
  const int n = A.n;
  DevBuf<int> t_color, t_rem, t_rows, t_cout, t_len, t_start, t_hist, t_newc; DevBuf<unsigned char> t_cand, t_cub;
  auto geti = [&](DevBuf<int>& own, size_t cnt) -> int* { if (ws) return ws->get<int>(cnt); own.alloc(cnt, s); return own.p; };
  int* color = geti(t_color, n);
  int* rem = geti(t_rem, 2);                       // [0] remaining, [1] overflow
  unsigned char* cand;
  if (ws) cand = ws->get<unsigned char>(n); else { t_cand.alloc(n, s); cand = t_cand.p; }
  PNP_CUDA(cudaMemsetAsync(color, 0xff, (size_t)n * sizeof(int), s));
  PNP_CUDA(cudaMemsetAsync(rem, 0, 2 * sizeof(int), s));
  int hrem = n, rounds = 0;
  while (hrem > 0) {
    PNP_CUDA(cudaMemsetAsync(rem, 0, sizeof(int), s));
    color_select_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, ja_nat, color, cand, rem); PNP_COUNT_LAUNCH();
    color_assign_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, ja_nat, color, cand, rem + 1); PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaMemcpyAsync(&hrem, rem, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (++rounds > 100000) throw Error(ERROR_UNKNOWN, "colouring did not terminate");
  }
  int hh2[2];
  PNP_CUDA(cudaMemcpyAsync(hh2, rem, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if (hh2[1]) throw Error(ERROR_DATA_STRUCTURE, "matrix graph needs more than 62 colours");
  // iterated greedy passes (reverse class order, then forward): fewer, fuller colour classes
  {
    int passes = 3;
    if (const char* e = getenv("PNPB200_RECOLOR_PASSES")) passes = atoi(e);
    int* newc = geti(t_newc, n);
    for (int ps = 0; ps < passes; ++ps) {
      int ncol_now = 0;
      PNP_CUDA(cudaMemsetAsync(rem, 0, sizeof(int), s));
      max_color_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, color, rem); PNP_COUNT_LAUNCH();
      PNP_CUDA(cudaMemcpyAsync(&ncol_now, rem, sizeof(int), cudaMemcpyDeviceToHost, s));
      PNP_CUDA(cudaStreamSynchronize(s));
      ++ncol_now;
      PNP_CUDA(cudaMemsetAsync(newc, 0xff, (size_t)n * sizeof(int), s));
      for (int q = 0; q < ncol_now; ++q) {
        const int cls = (ps % 2 == 0) ? ncol_now - 1 - q : q;
        recolor_class_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, ja_nat, color, cls, newc); PNP_COUNT_LAUNCH();
      }
      PNP_CUDA(cudaMemcpyAsync(color, newc, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
  }
  // optional row tiles (sweep order tile -> colour); measured on B200 at 256^3: no gain over one
  // tile (6.8 ms vs 5.6 ms per sweep: 26x more launches, L2 does not hold the window), so default 1
  int ntiles = 1;
  if (const char* e = getenv("PNPB200_GS_TILES")) { ntiles = atoi(e); if (ntiles < 1) ntiles = 1; if (ntiles > 64) ntiles = 64; }
  const int tile_rows = (n + ntiles - 1) / ntiles;
  const int ngroups_max = ntiles * 64;
  int* hist = geti(t_hist, ngroups_max);
  PNP_CUDA(cudaMemsetAsync(hist, 0, (size_t)ngroups_max * sizeof(int), s));
  const int n_own = owned_rows(L);
  group_key_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, n_own, ngroups_max - 1, tile_rows, color, hist); PNP_COUNT_LAUNCH();
  std::vector<int> hh(ngroups_max);
  PNP_CUDA(cudaMemcpyAsync(hh.data(), hist, (size_t)ngroups_max * sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  // non-empty (tile, colour) groups in sweep order
  L.color_ptr.assign(1, 0);
  int maxc = 0;
  for (int g = 0; g < ngroups_max; ++g)
    if (hh[g] > 0) { L.color_ptr.push_back(L.color_ptr.back() + hh[g]); if (g % 64 + 1 > maxc) maxc = g % 64 + 1; }
  L.ngroups = (int)L.color_ptr.size() - 1;
  L.d_color_ptr.alloc(L.color_ptr.size(), s);
  L.d_color_ptr.upload(L.color_ptr.data(), L.color_ptr.size());
  if (n_own < n && hh[ngroups_max - 1] > 0) { --L.ngroups; maxc = 0; for (int g = 0; g < ngroups_max - 1; ++g) if (hh[g] > 0 && g % 64 + 1 > maxc) maxc = g % 64 + 1; }
  L.ncolors = maxc;
  int key_bits = 6; while ((1 << key_bits) < ngroups_max) ++key_bits;
  // rowid = rows sorted by (tile, colour), ascending inside a group (stable radix sort)
  int* rows_in = geti(t_rows, n);
  int* color_out = geti(t_cout, n);
  A.rowid.alloc(n, s);
  iota_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, rows_in); PNP_COUNT_LAUNCH();
  size_t bytes = 0;
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, color, color_out, rows_in, A.rowid.p, n, 0, key_bits, s));
  unsigned char* tmp;
  if (ws) tmp = ws->get<unsigned char>(bytes); else { t_cub.alloc(bytes, s); tmp = t_cub.p; }
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, color, color_out, rows_in, A.rowid.p, n, 0, key_bits, s));
  PNP_COUNT_LAUNCH();
  // physical row starts and the physical column array
  int* len = geti(t_len, n + 1);
  A.pia.alloc(n + 1, s);
  int* start = A.pia.p;
  rowlen_by_position_kernel<<<cdiv(n + 1, 256), 256, 0, s>>>(n, A.ia.p, A.rowid.p, len); PNP_COUNT_LAUNCH();
  exclusive_scan_int(len, start, n + 1, s);
  A.rs.alloc(n, s); A.rsl.alloc(n, s);
  scatter_rs_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.rowid.p, start, A.rs.p, A.rsl.p); PNP_COUNT_LAUNCH();
  A.ja.alloc(A.nnzb, s);
  permute_ja_kernel<true><<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, ja_nat, A.ja.p); PNP_COUNT_LAUNCH();
  A.brow.alloc(A.nnzb, s); A.dslot.alloc(n, s); A.tslot.alloc(A.nnzb, s);
  row_tables_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, A.brow.p, A.dslot.p); PNP_COUNT_LAUNCH();
  tslot_kernel<<<cdiv(A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.ja.p, A.brow.p, A.tslot.p); PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}

} // namespace pnpb200
