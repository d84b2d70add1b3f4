// pnp-b200 — BSR(3) kernels on the BSR-T layout:
//   fasp_blas_dbsr_mxv / aAxpy   [EXT FASP]      -> bsrt_spmv / bsrt_residual
//   fasp_dbsr_getdiaginv         [EXT FASP]      -> bsrt_diaginv
//   fasp_smoother_dbsr_gs_*      [EXT FASP]      -> multicolour block Gauss-Seidel
// A half-warp (16 lanes) owns one block row; lane k owns block k of the row, so each of the 9
// component loads is unit-stride across lanes. All kernels are HBM-bound streaming kernels.
#include "internal.hpp"
#include <cub/cub.cuh>

namespace pnpb200 {

namespace {

// y = A x   (MODE 0)     y = b - A x   (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(256)
bsrt_spmv_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                 const double* __restrict__ val, const double* __restrict__ x,
                 const double* __restrict__ b, double* __restrict__ y)
{
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  int s = 0, len = 0;
  if (row < n) { s = ia[row]; len = ia[row + 1] - s; }
  const double* v = val + 9 * (size_t)s;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int k = lane; k < len; k += 16) {
    const size_t j = 3 * (size_t)ja[s + k];
    const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2];
    a0 += v[k] * x0 + v[len + k] * x1 + v[2 * len + k] * x2;
    a1 += v[3 * len + k] * x0 + v[4 * len + k] * x1 + v[5 * len + k] * x2;
    a2 += v[6 * len + k] * x0 + v[7 * len + k] * x1 + v[8 * len + k] * x2;
  }
  a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
  if (row < n && lane < 3) {
    const double r = lane == 0 ? a0 : lane == 1 ? a1 : a2;
    const size_t o = 3 * (size_t)row + lane;
    y[o] = MODE == 0 ? r : b[o] - r;
  }
}

__global__ void diaginv_kernel(int n, const int* __restrict__ ia, const int* __restrict__ dslot,
                               const double* __restrict__ val, double* __restrict__ dinv)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = ia[i], len = ia[i + 1] - s, d = dslot[i] - s;
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

// FASP layout (block-major, row-major 3x3) <-> BSR-T
template <bool TO_T>
__global__ void relayout_kernel(int nnzb, const int* __restrict__ ia, const int* __restrict__ brow,
                                const double* __restrict__ in, double* __restrict__ out)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * (size_t)nnzb) return;
  const int b = (int)(e / 9), t = (int)(e % 9);
  const int i = brow[b], s = ia[i], len = ia[i + 1] - s;
  const size_t et = 9 * (size_t)s + (size_t)t * len + (b - s);
  if (TO_T) out[et] = in[e]; else out[e] = in[et];
}

// one colour of block Gauss-Seidel: x_i = D_i^{-1} (b_i - sum_{j != i} A_ij x_j)
__global__ void __launch_bounds__(256)
mcgs_kernel(int nrows_color, const int* __restrict__ rows, const int* __restrict__ ia,
            const int* __restrict__ ja, const double* __restrict__ val,
            const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x)
{
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  int row = -1, s = 0, len = 0;
  if (t < nrows_color) { row = rows[t]; s = ia[row]; len = ia[row + 1] - s; }
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

// ---- Jones-Plassmann colouring with hashed priorities (two kernels per round so that the
// result does not depend on thread timing) ---------------------------------------------------
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

__global__ void color_hist_kernel(int n, const int* __restrict__ color, int* __restrict__ hist)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[color[i]], 1);
}

__global__ void iota_kernel(int n, int* __restrict__ a)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

} // namespace

void bsrt_spmv(const BsrT& A, const double* x, double* y, cudaStream_t s)
{
  bsrt_spmv_kernel<0><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(A.n, A.ia.p, A.ja.p, A.val.p, x, nullptr, y);
  PNP_COUNT_LAUNCH();
}

void bsrt_residual(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s)
{
  bsrt_spmv_kernel<1><<<cdiv((size_t)A.n * 16, 256), 256, 0, s>>>(A.n, A.ia.p, A.ja.p, A.val.p, x, b, r);
  PNP_COUNT_LAUNCH();
}

void bsrt_diaginv(const BsrT& A, double* dinv, cudaStream_t s)
{
  diaginv_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, A.dslot.p, A.val.p, dinv);
  PNP_COUNT_LAUNCH();
}

void bsrt_from_fasp(const int* d_ia, const int* d_ja, const double* d_val_fasp, BsrT& A, cudaStream_t s)
{
  (void)d_ia; (void)d_ja;
  relayout_kernel<true><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.brow.p, d_val_fasp, A.val.p);
  PNP_COUNT_LAUNCH();
}

void bsrt_to_fasp(const BsrT& A, double* d_val_fasp, cudaStream_t s)
{
  relayout_kernel<false><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.brow.p, A.val.p, d_val_fasp);
  PNP_COUNT_LAUNCH();
}

void mcgs_sweep(const Level& L, const double* b, double* x, bool reverse, cudaStream_t s)
{
  const BsrT& A = L.A;
  for (int cc = 0; cc < L.ncolors; ++cc) {
    const int c = reverse ? L.ncolors - 1 - cc : cc;
    const int cnt = L.color_ptr[c + 1] - L.color_ptr[c];
    if (cnt == 0) continue;
    mcgs_kernel<<<cdiv((size_t)cnt * 16, 256), 256, 0, s>>>(cnt, L.color_rows.p + L.color_ptr[c], A.ia.p, A.ja.p,
                                                           A.val.p, L.dinv.p, b, x);
    PNP_COUNT_LAUNCH();
  }
}

void color_level(Level& L, cudaStream_t s)
{
  BsrT& A = L.A;
  const int n = A.n;
  DevBuf<int> color; color.alloc(n, s);
  PNP_CUDA(cudaMemsetAsync(color.p, 0xff, n * sizeof(int), s));
  DevBuf<unsigned char> cand; cand.alloc(n, s);
  DevBuf<int> rem; rem.alloc(66, s); rem.zero();     // [0] remaining, [1] overflow, [2..65] histogram
  int hrem = n, rounds = 0;
  while (hrem > 0) {
    PNP_CUDA(cudaMemsetAsync(rem.p, 0, sizeof(int), s));
    color_select_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.ja.p, color.p, cand.p, rem.p);
    PNP_COUNT_LAUNCH();
    color_assign_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.ja.p, color.p, cand.p, rem.p + 1);
    PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaMemcpyAsync(&hrem, rem.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (++rounds > 100000) throw Error(ERROR_UNKNOWN, "colouring did not terminate");
  }
  color_hist_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, color.p, rem.p + 2); PNP_COUNT_LAUNCH();
  std::vector<int> hh = rem.to_host(66);
  if (hh[1]) throw Error(ERROR_DATA_STRUCTURE, "matrix graph needs more than 62 colours");
  int nc = 0;
  for (int c = 0; c < 64; ++c) if (hh[2 + c] > 0) nc = c + 1;
  L.ncolors = nc;
  L.color_ptr.assign(nc + 1, 0);
  for (int c = 0; c < nc; ++c) L.color_ptr[c + 1] = L.color_ptr[c] + hh[2 + c];
  // rows sorted by colour, ascending inside a colour (stable radix sort on 6 bits)
  DevBuf<int> rows_in, color_out; rows_in.alloc(n, s); color_out.alloc(n, s);
  L.color_rows.alloc(n, s);
  iota_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, rows_in.p); PNP_COUNT_LAUNCH();
  size_t bytes = 0;
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, color.p, color_out.p, rows_in.p, L.color_rows.p, n, 0, 6, s));
  DevBuf<unsigned char> tmp; tmp.alloc(bytes, s);
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, color.p, color_out.p, rows_in.p, L.color_rows.p, n, 0, 6, s));
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}

} // namespace pnpb200
