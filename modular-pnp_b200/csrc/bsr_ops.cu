// pnp-b200 — BSR(3) kernels on the BSR-T/LDU layout:
//   fasp_blas_dbsr_mxv / aAxpy   [EXT FASP]      -> bsrt_spmv / bsrt_residual (+ _rowid forms for the ABI)
//   fasp_dbsr_getdiaginv         [EXT FASP]      -> bsrt_diaginv
//   fasp_smoother_dbsr_gs_*      [EXT FASP]      -> multicolour block Gauss-Seidel
// A half-warp (16 lanes) owns one block row; lane k owns block k of the row, so each of the 9
// component loads is unit-stride across lanes. Rows are STORED colour by colour (rs[] = physical
// start of a row, rowid[] = row at a physical position), each row as [L | D | U] (internal.hpp): one
// Gauss-Seidel colour is one contiguous piece of ja/val. The solver's vectors live in sweep space
// (BsrT::vrowid); the kernels take the index arrays of either space. All kernels here are HBM-bound
// streaming kernels (see profiles/README.md for what bounds each of them).
#include "internal.hpp"
#include <cub/cub.cuh>
#include <algorithm>

namespace pnpb200 {

namespace {

// One block row per half-warp, lane k = block k of the row (BSR-T/LDU layout, internal.hpp).
// rd[] holds one 16-byte descriptor (start, length, nl, vector index of the row) per row: the
// passes over all rows get the descriptors in vector order (vrd, or rd for row-id space), the
// colour kernels prd + p0 (descriptors in sweep order), so vector index, start, length and the L/U
// split come from ONE load. ja[] holds the columns as indices into the vectors of the same space.
//   M_SPMV   y = A x                     M_RES  y = b - A x             (vector order, full rows)
//   M_GS     x_row = D^{-1}(b - sum_{j != row} A_row,j x_j), y == x: one Gauss-Seidel colour
//   M_GS0    the same from a ZERO initial guess, forward sweep: only the lower part contributes
//            (the upper neighbours are still zero), so only L is read
//   M_RESU   r = b - A x right after an M_GS0 sweep: the row equation holds with the lower part,
//            so r_row = -sum_{upper} A_row,j x_j and only U is read (rows >= n_swept, the ghost
//            rows of a multi-GPU level, take the full row)
//   M_GSD    M_GS that also writes delta_row = x_new - x_old
//   M_LDELTA y = A x_new after a BACKWARD M_GSD sweep = b + sum_{lower} A_row,j delta_j: row i was
//            updated with the new upper and the old lower neighbours, so A x_new differs from b
//            only by L (x_new - x_old); only L is read (x = delta)
//   M_LDELTAG the same on a multi-GPU level (hybrid Gauss-Seidel: ghost columns keep their
//            pre-sweep value during the sweep): y = b + sum_{lower or ghost} A_row,j delta_j with
//            delta_ghost = halo exchange of the owners' delta. Ghost columns (>= n_swept) are the
//            tail of the upper part; their column indices are read for the whole row, their
//            values only where needed. Ghost rows get y = b.
// One preconditioner application + Krylov matvec therefore reads the fine matrix 2.5 times
// (L, U, L+D+U, L) instead of 4 times. The colour kernels are compiled for 8 resident CTAs per SM
// (32 registers), the passes over all rows run in the pipelined form below.
// Tried and measured slower: ld.global.cs / L2 evict_first+evict_last hints on the matrix loads
// (6.3-7.2 ms per sweep), requesting b and D^{-1} before the row's blocks (6.2 vs 5.5 ms: the extra
// registers cost more occupancy than the overlap gains); gathering the 24 bytes of a vertex with one
// 16-byte + one 8-byte load instead of three 8-byte loads (SpMV 3.78 vs 3.33 ms).
#ifndef PNPB200_SUM3
#define PNPB200_SUM3 1
#endif
enum { M_SPMV = 0, M_RES = 1, M_GS = 2, M_GS0 = 3, M_RESU = 4, M_GSD = 5, M_LDELTA = 6, M_LDELTAG = 7 };

template <int MODE, int MINB, int XS, int NV>
__global__ void __launch_bounds__(256, MINB)
bsrt_row_kernel(int nrows, int n_swept, const int4* __restrict__ rd, const int* __restrict__ ja,
                const double* __restrict__ val, const double* __restrict__ dinv, const double* __restrict__ b,
                const double* x, double* y, double* dl, int bs, int ys)
{
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  int row = -1, s = 0, len = 0, nl = 0;
  if (t < nrows) { const int4 d = rd[t]; s = d.x; len = d.y; nl = d.z; row = d.w; }
  int k0 = 0, k1 = len;
  if (MODE == M_GS0 || MODE == M_LDELTA) k1 = nl;
  if (MODE == M_LDELTAG && row >= n_swept) k1 = 0;
  if (MODE == M_RESU && row < n_swept) k0 = nl + 1;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  auto block = [&](int k) {
    if ((MODE == M_GS || MODE == M_GSD) && k == nl) return;
    // lower part: component stride nl; D + U part (k >= nl) starts NV*nl further: stride len - nl
    const bool lo = (MODE == M_GS0 || MODE == M_LDELTA) ? true : k < nl;
    const double* v = val + (NV * (size_t)s + (size_t)(lo ? k : (NV - 1) * nl + k));
    const int st = lo ? nl : len - nl;
    const int jcol = ja[s + k];
    if (MODE == M_LDELTAG && k >= nl && (k == nl || jcol < n_swept)) return;
    double x0, x1, x2;
    load_row3<XS>(x, jcol, x0, x1, x2);
    block_mac<NV>(v, st, x0, x1, x2, a0, a1, a2);
  };
  // the first 16 blocks of the row are peeled out of the loop so that the compiler issues their 13
  // loads back to back (a rolled loop under the register cap interleaves loads and FMAs and pays
  // several serial DRAM round trips per row: measured 0.66 vs 0.46 ms for the SpMV at 128^3)
  int k = k0 + lane;
  if (k < k1) block(k);
#pragma unroll 1
  for (k += 16; k < k1; k += 16) block(k);
#if PNPB200_SUM3
  // lane (c, i) = (lane >> 2, lane & 3) of the half-warp holds the row sum of component c
  const double a = half_sum3(a0, a1, a2, lane);
  const int c = lane >> 2, i = lane & 3;
  if (MODE == M_SPMV || MODE == M_RES || MODE == M_RESU || MODE == M_LDELTA || MODE == M_LDELTAG) {
    // b and y carry bs / ys doubles per row (3 = packed: the Krylov basis, 4 = padded: pad entry kept at zero)
    if (row >= 0 && i == 0 && c < ys) {        // c == 3: a == 0
      const size_t o = (size_t)ys * (size_t)row + c, ob = (size_t)bs * (size_t)row + c;
      if (MODE == M_SPMV || c == 3) y[o] = a;
      else if (MODE == M_RES) y[o] = b[ob] - a;
      else if (MODE == M_RESU) y[o] = row < n_swept ? -a : b[ob] - a;
      else y[o] = b[ob] + a;
    }
  } else {
    // x_new[i] = sum_c Dinv[i][c] (b_c - a_c): lane (c, i) forms one product, two shuffles add the three c
    double part = 0.0;
    if (row >= 0 && c < 3 && i < 3) part = dinv[9 * (size_t)t + 3 * i + c] * (b[(size_t)bs * (size_t)row + c] - a);   // dinv: sweep order
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    part += __shfl_xor_sync(0xffffffffu, part, 8);
    if (row >= 0 && lane < XS) {               // lane 3 (padded vectors): part == 0
      const size_t o = (size_t)XS * (size_t)row + lane;
      if (MODE == M_GSD) dl[o] = lane < 3 ? part - x[o] : 0.0;
      y[o] = part;
    }
  }
#else
  static_assert(XS == 3, "the PNPB200_SUM3=0 epilogue only knows packed vectors");
  a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
  if (row >= 0 && lane < 3) {
    const size_t o = 3 * (size_t)row;
    const double a = lane == 0 ? a0 : lane == 1 ? a1 : a2;
    if (MODE == M_SPMV) y[o + lane] = a;
    else if (MODE == M_RES) y[o + lane] = b[o + lane] - a;
    else if (MODE == M_RESU) y[o + lane] = row < n_swept ? -a : b[o + lane] - a;
    else if (MODE == M_LDELTA || MODE == M_LDELTAG) y[o + lane] = b[o + lane] + a;
    else {
      const double r0 = b[o] - a0, r1 = b[o + 1] - a1, r2 = b[o + 2] - a2;
      const double* d = dinv + 9 * (size_t)t + 3 * lane;
      const double xn = d[0] * r0 + d[1] * r1 + d[2] * r2;
      if (MODE == M_GSD) dl[o + lane] = xn - x[o + lane];
      y[o + lane] = xn;
    }
  }
#endif
}

// Software-pipelined, persistent form of bsrt_row_kernel (same modes, same arithmetic). In the
// one-row-per-half-warp kernel every row pays a chain of three dependent memory round trips
// (descriptor -> column indices -> gathered x), so the passes that read only part of a row (L or U)
// gain little: they are latency bound, not bandwidth bound. Here a half-warp walks rows t, t+H,
// t+2H, ... (H = half-warps in the grid) and carries the descriptor of the next two rows and the
// column indices of the next row in registers, so that the matrix values AND the x gather of a row
// are requested in the same instant and a row costs ONE exposed round trip.
template <int MODE, int MINB, int XS, int NV>
__global__ void __launch_bounds__(256, MINB)
bsrt_row_pipe_kernel(int nrows, int n_swept, const int4* __restrict__ rd, const int* __restrict__ ja,
                     const double* __restrict__ val, const double* __restrict__ dinv, const double* __restrict__ b,
                     const double* x, double* y, double* dl, int bs, int ys)
{
  const int lane = threadIdx.x & 15;
  const int H = (gridDim.x * blockDim.x) >> 4;
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int4 none = make_int4(0, 0, 0, -1);
  // first block index of a row's range in this mode, and one past its last
  auto range = [&](const int4& d, int& k0, int& k1) {
    k0 = 0; k1 = d.y;
    if (MODE == M_GS0 || MODE == M_LDELTA) k1 = d.z;
    if (MODE == M_LDELTAG && d.w >= n_swept) k1 = 0;
    if (MODE == M_RESU && d.w < n_swept) k0 = d.z + 1;
  };
  auto first_ja = [&](const int4& d) -> int {
    int k0, k1; range(d, k0, k1);
    const int k = k0 + lane;
    return (d.w >= 0 && k < k1) ? ja[d.x + k] : -1;
  };
  int4 d0 = t < nrows ? rd[t] : none;
  int4 d1 = t + H < nrows ? rd[t + H] : none;
  int jc0 = first_ja(d0);
  // the loop bound is uniform over the warp (its two half-warps own rows t and t + 1)
  for (int tw = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 1; tw < nrows; tw += H) {
    // requests for the rows after this one
    const int4 d2 = t + 2 * H < nrows ? rd[t + 2 * H] : none;
    const int jc1 = first_ja(d1);
    // this row
    const int s = d0.x, len = d0.y, nl = d0.z, row = d0.w;
    int k0, k1; range(d0, k0, k1);
    const int c = lane >> 2, i = lane & 3;
    double bq = 0.0, dq = 0.0;          // epilogue operands, requested up front as well
    if (row >= 0 && c < 3) {
      if (MODE == M_RES || MODE == M_LDELTA || MODE == M_LDELTAG || (MODE == M_RESU && row >= n_swept)) { if (i == 0) bq = b[(size_t)bs * (size_t)row + c]; }
      if (MODE == M_GS || MODE == M_GS0 || MODE == M_GSD) { if (i < 3) { bq = b[(size_t)bs * (size_t)row + c]; dq = dinv[9 * (size_t)t + 3 * i + c]; } }
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    auto block = [&](int k, int jc) {
      if ((MODE == M_GS || MODE == M_GSD) && k == nl) return;
      if (MODE == M_LDELTAG && k >= nl && (k == nl || jc < n_swept)) return;
      const bool lo = (MODE == M_GS0 || MODE == M_LDELTA) ? true : k < nl;
      const double* v = val + (NV * (size_t)s + (size_t)(lo ? k : (NV - 1) * nl + k));
      const int st = lo ? nl : len - nl;
      double x0, x1, x2;
      load_row3<XS>(x, jc, x0, x1, x2);
      block_mac<NV>(v, st, x0, x1, x2, a0, a1, a2);
    };
    int k = k0 + lane;
    if (jc0 >= 0) block(k, jc0);
#pragma unroll 1
    for (k += 16; k < k1; k += 16) block(k, ja[s + k]);      // rows with more than 16 blocks in range (rare)
    const double a = half_sum3(a0, a1, a2, lane);
    if (MODE == M_SPMV || MODE == M_RES || MODE == M_RESU || MODE == M_LDELTA || MODE == M_LDELTAG) {
      if (row >= 0 && i == 0 && c < ys) {      // c == 3 (padded y): a == 0 and bq == 0, the pad entry stays zero
        const size_t o = (size_t)ys * (size_t)row + c;
        if (MODE == M_SPMV) y[o] = a;
        else if (MODE == M_RES) y[o] = bq - a;
        else if (MODE == M_RESU) y[o] = row < n_swept ? -a : bq - a;
        else y[o] = bq + a;
      }
    } else {
      double part = dq * (bq - a);        // lanes outside (c < 3, i < 3): dq = 0
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 8);
      if (row >= 0 && lane < XS) {             // lane 3 (padded vectors): part == 0
        const size_t o = (size_t)XS * (size_t)row + lane;
        if (MODE == M_GSD) dl[o] = lane < 3 ? part - x[o] : 0.0;
        y[o] = part;
      }
    }
    d0 = d1; d1 = d2; jc0 = jc1; t += H;
  }
}

// launch one pass over nrows rows described by rd[]. Measured at 256^3 (fine level, ms per pass;
// "plain" = bsrt_row_kernel, "pipe" = bsrt_row_pipe_kernel, n = resident CTAs per SM):
//   pass            plain(6/8)  pipe(4)  pipe(6, spills)
//   y = A x           3.39       3.30      3.31
//   r = b - A x       3.71       3.39      3.39
//   r = -U x          2.93       2.47      2.57
//   y = b + L delta   3.18       2.48      2.48
//   GS sweep          4.53       5.20      6.93        (colour launches: plain kernel, 8 CTAs/SM)
//   GS sweep, L only  3.73       3.74      4.95
// so the natural-order passes run pipelined with 64 registers, the colour launches plain with 32.
#ifndef PNPB200_PIPE
#define PNPB200_PIPE 1
#endif
#ifndef PNPB200_NAT_CTAS
#define PNPB200_NAT_CTAS 4        // resident CTAs per SM the natural-order passes are compiled for
#endif
#ifndef PNPB200_GS_CTAS
#define PNPB200_GS_CTAS 8         // ... and the colour sweeps
#endif
int g_sm_count = 0;
template <int MODE, int MINB, bool PIPE, int XS, int NV>
void launch_rows_nv(int nrows, int n_swept, const int4* rd, const int* ja, const double* val, const double* dinv, const double* b,
                    const double* x, double* y, double* dl, cudaStream_t s, int bs, int ys)
{
  const int full = cdiv((size_t)nrows * 16, 256);
  if constexpr (PIPE && PNPB200_PIPE) {
  // persistent grid = exactly the CTAs that are resident at once (a second wave would walk the
  // gathered x window a second time: measured 6.5 vs 4.5 ms per Gauss-Seidel sweep)
  if (g_sm_count == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev); if (g_sm_count <= 0) g_sm_count = 148; }
  static int per_sm = 0;             // one static per (MODE, MINB, XS, NV) instantiation
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bsrt_row_pipe_kernel<MODE, MINB, XS, NV>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = MINB;
  }
  const int cap = g_sm_count * per_sm;
  bsrt_row_pipe_kernel<MODE, MINB, XS, NV><<<full < cap ? full : cap, 256, 0, s>>>(nrows, n_swept, rd, ja, val, dinv, b, x, y, dl, bs, ys);
  } else
  bsrt_row_kernel<MODE, MINB, XS, NV><<<full, 256, 0, s>>>(nrows, n_swept, rd, ja, val, dinv, b, x, y, dl, bs, ys);
  PNP_COUNT_LAUNCH();
}
// the level's values: 7 per block when the level was packed after the set-up (bsrt_pack7), else FASP's 9
template <int MODE, int MINB, bool PIPE, int XS = VS>
void launch_rows(int nrows, int n_swept, const int4* rd, const int* ja, const BsrT& A, const double* dinv, const double* b,
                 const double* x, double* y, double* dl, cudaStream_t s, int bs = XS, int ys = XS)
{
  if (nrows <= 0) return;
  if (A.packed7) launch_rows_nv<MODE, MINB, PIPE, XS, 7>(nrows, n_swept, rd, ja, A.val7.p, dinv, b, x, y, dl, s, bs, ys);
  else launch_rows_nv<MODE, MINB, PIPE, XS, 9>(nrows, n_swept, rd, ja, A.val.p, dinv, b, x, y, dl, s, bs, ys);
}

// Small levels (<= SMALL_LEVEL_ROWS rows): the whole multicolour sweep in ONE launch of one CTA,
// colours separated by __syncthreads() instead of kernel boundaries (the coarsest levels of the
// K-cycle are visited 8-16 times per Krylov iteration and are pure launch latency). Measured at
// 256^3: threshold 600 -> -1.3 % Krylov time, threshold 6144 -> +2.5 % (one SM is too slow for a
// 3.5 k-row level), so only the last one or two levels take this path.
// mode 0: plain sweep, 1: forward sweep from a zero initial guess (lower parts only),
// 2: sweep that also writes delta = x_new - x_old.
constexpr int SMALL_LEVEL_ROWS = 512;
__global__ void __launch_bounds__(1024, 1)
mcgs_small_kernel(int ngroups, int reverse, int mode, const int* __restrict__ gptr, const int4* __restrict__ prd,
                  const int* __restrict__ ja, const double* __restrict__ val,
                  const double* __restrict__ dinv, const double* __restrict__ b, double* x, double* dl, int bs)
{
  const int hw = threadIdx.x >> 4, nhw = blockDim.x >> 4, lane = threadIdx.x & 15;
  for (int gg = 0; gg < ngroups; ++gg) {
    const int g = reverse ? ngroups - 1 - gg : gg;
    const int p0 = gptr[g], p1 = gptr[g + 1];
    for (int base = p0; base < p1; base += nhw) {        // uniform trip count: shuffles stay converged
      const int p = base + hw;
      int row = -1, s = 0, len = 0, nl = 0;
      if (p < p1) { const int4 d = prd[p]; s = d.x; len = d.y; nl = d.z; row = d.w; }
      const int k1 = mode == 1 ? nl : len;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int k = lane; k < k1; k += 16) {
        if (k == nl) continue;
        int st;
        const double* v = val + bsrt_off(s, len, nl, k, st);
        double x0, x1, x2;
        load_row3<VS>(x, ja[s + k], x0, x1, x2);
        a0 += v[0] * x0 + v[st] * x1 + v[2 * st] * x2;
        a1 += v[3 * st] * x0 + v[4 * st] * x1 + v[5 * st] * x2;
        a2 += v[6 * st] * x0 + v[7 * st] * x1 + v[8 * st] * x2;
      }
      a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
      if (row >= 0 && lane < 3) {
        const size_t o = (size_t)VS * (size_t)row, ob = (size_t)bs * (size_t)row;
        const double r0 = b[ob] - a0, r1 = b[ob + 1] - a1, r2 = b[ob + 2] - a2;
        const double* d = dinv + 9 * (size_t)p + 3 * lane;
        const double xn = d[0] * r0 + d[1] * r1 + d[2] * r2;
        if (mode == 2) dl[o + lane] = xn - x[o + lane];
        x[o + lane] = xn;
      }
      if (VS == 4 && row >= 0 && lane == 3) { x[4 * (size_t)row + 3] = 0.0; if (mode == 2) dl[4 * (size_t)row + 3] = 0.0; }
    }
    __syncthreads();
  }
}

// The same sweep for a level that fits in the shared memory of one SM (values, columns, descriptors,
// block inverses, b and x: <= SMALL_SMEM_BYTES): everything is staged with one round of coalesced
// loads, the colours then run out of shared memory. The global-memory version pays three dependent
// L2 round trips per colour (36 us per sweep of a 150-row level, 16 such sweeps per Krylov
// iteration of the K-cycle); this one 14-17 us. x is kept in shared memory and written through.
// Tried for the mid-size levels (3.5 k - 80 k rows, 11 colour launches of 4-8 us each): the whole
// sweep as ONE cooperative launch with a grid barrier between colours and the next colour's matrix
// operands requested ahead of the barrier: 43 instead of 62 us per sweep of the 3.5 k-row level,
// slower on the 80 k-row level, 0.5 % of the Krylov time in total — not kept.
constexpr size_t SMALL_SMEM_BYTES = 200 * 1024;
inline size_t small_level_smem(int n, int nnzb) { return (size_t)nnzb * 76 + (size_t)n * (72 + 16 * VS + 16) + 64; }
__global__ void __launch_bounds__(1024, 1)
mcgs_smem_kernel(int n, int nnzb, int ngroups, int reverse, int mode, const int* __restrict__ gptr,
                 const int4* __restrict__ prd, const int* __restrict__ ja, const double* __restrict__ val,
                 const double* __restrict__ dinv, const double* __restrict__ b, double* x, double* dl, int bs)
{
  extern __shared__ double sm_d[];
  double* s_val = sm_d;                         // 9*nnzb
  double* s_dinv = s_val + 9 * (size_t)nnzb;    // 9*n   (sweep order)
  double* s_b = s_dinv + 9 * (size_t)n;         // VS*n
  double* s_x = s_b + VS * (size_t)n;           // VS*n
  int4* s_prd = reinterpret_cast<int4*>(s_x + VS * (size_t)n + ((9 * nnzb + (9 + 2 * VS) * n) & 1));     // 16-byte aligned: n entries
  int* s_ja = reinterpret_cast<int*>(s_prd + n);                            // nnzb
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < 9 * nnzb; e += nt) s_val[e] = val[e];
  for (int e = tid; e < 9 * n; e += nt) s_dinv[e] = dinv[e];
  for (int e = tid; e < VS * n; e += nt) { const int k = e % VS; s_b[e] = k < 3 ? b[(size_t)bs * (e / VS) + k] : 0.0; s_x[e] = mode == 1 ? 0.0 : x[e]; }
  for (int e = tid; e < n; e += nt) s_prd[e] = prd[e];
  for (int e = tid; e < nnzb; e += nt) s_ja[e] = ja[e];
  __syncthreads();
  const int hw = tid >> 4, nhw = nt >> 4, lane = tid & 15;
  for (int gg = 0; gg < ngroups; ++gg) {
    const int g = reverse ? ngroups - 1 - gg : gg;
    const int p0 = gptr[g], p1 = gptr[g + 1];
    for (int base = p0; base < p1; base += nhw) {        // uniform trip count: shuffles stay converged
      const int p = base + hw;
      int row = -1, s = 0, len = 0, nl = 0;
      if (p < p1) { const int4 d = s_prd[p]; s = d.x; len = d.y; nl = d.z; row = d.w; }
      const int k1 = mode == 1 ? nl : len;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int k = lane; k < k1; k += 16) {
        if (k == nl) continue;
        int st;
        const double* v = s_val + bsrt_off(s, len, nl, k, st);
        const int j = VS * s_ja[s + k];
        const double x0 = s_x[j], x1 = s_x[j + 1], x2 = s_x[j + 2];
        a0 += v[0] * x0 + v[st] * x1 + v[2 * st] * x2;
        a1 += v[3 * st] * x0 + v[4 * st] * x1 + v[5 * st] * x2;
        a2 += v[6 * st] * x0 + v[7 * st] * x1 + v[8 * st] * x2;
      }
      a0 = half_sum(a0); a1 = half_sum(a1); a2 = half_sum(a2);
      if (row >= 0 && lane < 3) {
        const int o = VS * row;
        const double r0 = s_b[o] - a0, r1 = s_b[o + 1] - a1, r2 = s_b[o + 2] - a2;
        const double* d = s_dinv + 9 * p + 3 * lane;
        const double xn = d[0] * r0 + d[1] * r1 + d[2] * r2;
        if (mode == 2) dl[o + lane] = xn - s_x[o + lane];
        s_x[o + lane] = xn;
        x[o + lane] = xn;
      }
      if (VS == 4 && row >= 0 && lane == 3) { x[4 * row + 3] = 0.0; if (mode == 2) dl[4 * row + 3] = 0.0; }
    }
    __syncthreads();
  }
}

// inverse of the diagonal blocks, stored in SWEEP order (dinv[9*p..] belongs to row rowid[p]): a
// colour launch of the smoother reads one contiguous piece of it
__global__ void diaginv_kernel(int n, const int4* __restrict__ prd, const double* __restrict__ val,
                               double* __restrict__ dinv)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int4 d = prd[p];
  int st;
  const double* v = val + bsrt_off(d.x, d.y, d.z, d.z, st);
  double a[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) a[t] = v[(size_t)t * st];
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  const double id = 1.0 / det;
  double* o = dinv + 9 * (size_t)p;
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// position of physical block b of row i in the row's NATURAL (column-ascending) order
__device__ __forceinline__ int natural_pos(const int* __restrict__ ja, int s, int len, int b)
{
  const int c = ja[b];
  int kn = 0;
  for (int p = s; p < s + len; ++p) kn += ja[p] < c ? 1 : 0;
  return kn;
}

// FASP layout (natural block order, row-major 3x3 per block) <-> BSR-T/LDU (physical)
template <bool TO_T>
__global__ void relayout_kernel(int nnzb, const int* __restrict__ ia, const int* __restrict__ rs,
                                const int* __restrict__ dslot, const int* __restrict__ ja,
                                const int* __restrict__ brow, const double* __restrict__ in, double* __restrict__ out)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * (size_t)nnzb) return;
  const int b = (int)(e / 9), t = (int)(e % 9);          // b = physical block
  const int i = brow[b], s = rs[i], len = ia[i + 1] - ia[i], k = b - s;
  int st;
  size_t et = bsrt_off(s, len, dslot[i] - s, k, st);
  et += (size_t)t * st;
  const size_t ef = 9 * ((size_t)ia[i] + natural_pos(ja, s, len, b)) + t;
  if (TO_T) out[et] = in[ef]; else out[ef] = in[et];
}

// physical -> natural column array (row granularity)
__global__ void export_ja_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                 const int* __restrict__ in, int* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = ia[i], len = ia[i + 1] - a, s = rs[i];
  for (int k = 0; k < len; ++k) out[a + natural_pos(in, s, len, s + k)] = in[s + k];
}

// natural -> physical: row i's blocks in the order [L | D | U] given the sweep key of every row
// (L = columns with a smaller key, i.e. swept before row i in a forward sweep); also the row of
// each block, the diagonal slot and the two row descriptors
__global__ void layout_row_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs, const int* __restrict__ key,
                                  const int* __restrict__ ja_nat, int* __restrict__ ja, int* __restrict__ brow,
                                  int* __restrict__ dslot, int4* __restrict__ rd, int* __restrict__ err)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = ia[i], len = ia[i + 1] - a, s = rs[i], ki = key[i];
  int w = s;
  bool hasd = false;
  for (int k = 0; k < len; ++k) { const int j = ja_nat[a + k]; if (j == i) hasd = true; else if (key[j] < ki) ja[w++] = j; }
  const int nl = w - s;
  if (hasd) ja[w++] = i; else *err = 1;
  for (int k = 0; k < len; ++k) { const int j = ja_nat[a + k]; if (j != i && key[j] >= ki) ja[w++] = j; }
  for (int p = s; p < s + len; ++p) brow[p] = i;
  dslot[i] = hasd ? s + nl : -1;
  rd[i] = make_int4(s, len, nl, i);
}

__global__ void physical_descriptor_kernel(int n, const int* __restrict__ rowid, const int* __restrict__ pos,
                                           const int4* __restrict__ rd, int4* __restrict__ prd)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { const int i = rowid[p]; int4 d = rd[i]; d.w = pos[i]; prd[p] = d; }      // .w: position in the solver vectors
}

__device__ __forceinline__ int find_sorted(const int* __restrict__ ja, int lo, int hi, int c)   // [lo, hi]
{
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, v = ja[mid];
    if (v == c) return mid;
    if (v < c) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

__global__ void tslot_kernel(int nnzb, const int* __restrict__ ia, const int* __restrict__ rs,
                             const int* __restrict__ dslot, const int* __restrict__ ja,
                             const int* __restrict__ brow, int* __restrict__ tslot)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int i = brow[b], j = ja[b];
  const int s = rs[j], e = s + (ia[j + 1] - ia[j]) - 1, d = dslot[j];
  int res;
  if (i == j) res = d;
  else {
    res = find_sorted(ja, s, d - 1, i);              // lower part of row j
    if (res < 0) res = find_sorted(ja, d + 1, e, i); // upper part
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
                                  int* __restrict__ rs)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rs[rowid[p]] = start[p];
}
// vector order key: (tile of consecutive row ids, colour); ghost rows last
__global__ void vector_key_kernel(int n, int n_own, int tile_rows, int ghost_key, const int* __restrict__ sweep_key, int* __restrict__ vkey)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vkey[i] = i >= n_own ? ghost_key : (i / tile_rows) * 64 + (sweep_key[i] & 63);
}
__global__ void invert_perm_kernel(int n, const int* __restrict__ vrowid, int* __restrict__ pos)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) pos[vrowid[q]] = q;
}
__global__ void vector_descriptor_kernel(int n, const int* __restrict__ vrowid, const int4* __restrict__ rd, int4* __restrict__ vrd)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) { int4 d = rd[vrowid[q]]; d.w = q; vrd[q] = d; }
}
__global__ void gather_pos_kernel(int n, const int* __restrict__ pos, const int* __restrict__ idx, int* __restrict__ out)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = pos[idx[e]];
}

// ---- speculative first-fit colouring (Gebremedhin-Manne), deterministic form -------------------
// Round: (1) every uncoloured row proposes the smallest colour no COLOURED neighbour has, (2) of two
// uncoloured neighbours that proposed the same colour the one with the higher hashed priority keeps
// it, (3) the winners' colours become final. Each kernel reads only what the previous kernel wrote,
// so the result does not depend on thread timing. 5-7 rounds on the AMG levels, where Jones-
// Plassmann (one independent set per round) needed 25-40: measured 8.2 -> 1.5 ms of colouring
// kernels per Newton step at 256^3. Natural CSR graph.
__global__ void color_propose_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                     const int* __restrict__ color, int* __restrict__ tent, int* __restrict__ overflow)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || color[i] >= 0) return;
  unsigned long long used = 0ull;
  for (int p = ia[i]; p < ia[i + 1]; ++p) {
    const int j = ja[p];
    if (j == i) continue;
    const int cj = color[j];
    if (cj >= 0) used |= 1ull << cj;
  }
  int c = __ffsll(~used) - 1;     // smallest free colour
  if (c < 0 || c >= 63) { *overflow = 1; c = 62; }
  tent[i] = c;
}

__global__ void color_resolve_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                     const int* __restrict__ color, const int* __restrict__ tent,
                                     unsigned char* __restrict__ win, int* __restrict__ remaining)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  win[i] = 0;
  if (color[i] >= 0) return;
  const int ci = tent[i];
  const unsigned pi = hash32((unsigned)i);
  bool keep = true;
  for (int p = ia[i]; p < ia[i + 1]; ++p) {
    const int j = ja[p];
    if (j == i || color[j] >= 0 || tent[j] != ci) continue;
    const unsigned pj = hash32((unsigned)j);
    if (pj > pi || (pj == pi && j > i)) { keep = false; break; }
  }
  if (keep) win[i] = 1; else atomicAdd(remaining, 1);
}

__global__ void color_commit_kernel(int n, const int* __restrict__ tent, const unsigned char* __restrict__ win, int* __restrict__ color)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && win[i]) color[i] = tent[i];
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

void bsrt_spmv(const BsrT& A, const double* x, double* y, cudaStream_t s, int ys)
{
  if (gs_stream_pass(A, M_SPMV, A.n, nullptr, x, y, VS, ys, s)) return;       // gs_stream.cu
  launch_rows<M_SPMV, PNPB200_NAT_CTAS, true>(A.n, A.n, A.vrd.p, A.jpos.p, A, nullptr, nullptr, x, y, nullptr, s, VS, ys);
}

void bsrt_residual(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s, int bs, int ys)
{
  if (gs_stream_pass(A, M_RES, A.n, b, x, r, bs, ys, s)) return;
  launch_rows<M_RES, PNPB200_NAT_CTAS, true>(A.n, A.n, A.vrd.p, A.jpos.p, A, nullptr, b, x, r, nullptr, s, bs, ys);
}

void bsrt_spmv_rowid(const BsrT& A, const double* x, double* y, cudaStream_t s)
{
  launch_rows<M_SPMV, PNPB200_NAT_CTAS, true, 3>(A.n, A.n, A.rd.p, A.ja.p, A, nullptr, nullptr, x, y, nullptr, s);
}

void bsrt_residual_rowid(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s)
{
  launch_rows<M_RES, PNPB200_NAT_CTAS, true, 3>(A.n, A.n, A.rd.p, A.ja.p, A, nullptr, b, x, r, nullptr, s);
}

void bsrt_residual_after_sweep(const Level& L, const double* b, const double* x, double* r, cudaStream_t s, int bs)
{
  const BsrT& A = L.A;
  const int own = owned_rows(L);
  if (gs_stream_pass(A, M_RESU, own, b, x, r, bs, VS, s)) {
    // the swept (owned) rows are the first `own` rows in matrix order as well; ghost rows of a multi-GPU level: full rows
    if (own < A.n) launch_rows<M_RESU, PNPB200_GS_CTAS, false>(A.n - own, own, A.prd.p + own, A.jpos.p, A, nullptr, b, x, r, nullptr, s, bs, VS);
    return;
  }
  launch_rows<M_RESU, PNPB200_NAT_CTAS, true>(A.n, own, A.vrd.p, A.jpos.p, A, nullptr, b, x, r, nullptr, s, bs, VS);
}

void bsrt_apply_after_sweep(const Level& L, const double* b, const double* delta, double* y, cudaStream_t s, int bs, int ys)
{
  const BsrT& A = L.A;
  if (owned_rows(L) != A.n) launch_rows<M_LDELTAG, PNPB200_NAT_CTAS, true>(A.n, owned_rows(L), A.vrd.p, A.jpos.p, A, nullptr, b, delta, y, nullptr, s, bs, ys);
  else if (gs_stream_pass(A, M_LDELTA, A.n, b, delta, y, bs, ys, s)) return;
  else launch_rows<M_LDELTA, PNPB200_NAT_CTAS, true>(A.n, A.n, A.vrd.p, A.jpos.p, A, nullptr, b, delta, y, nullptr, s, bs, ys);
}

namespace {
// row-id vectors are packed (3 per row), sweep-space vectors hold VS doubles per row (pad = 0)
__global__ void permute3_kernel(int n, const int* __restrict__ rowid, const double* __restrict__ in, double* __restrict__ out, int to_sweep, int vs)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)vs * (size_t)n) return;
  const size_t p = e / vs; const int k = (int)(e % vs);
  if (k == 3) { if (to_sweep) out[e] = 0.0; return; }
  const size_t r = 3 * (size_t)rowid[p] + k;
  if (to_sweep) out[e] = in[r]; else out[r] = in[e];
}
} // namespace

void vec_to_sweep(const BsrT& A, const double* in_rowid, double* out_sweep, cudaStream_t s, int vs)
{
  permute3_kernel<<<cdiv((size_t)vs * (size_t)A.n, 256), 256, 0, s>>>(A.n, A.vrowid.p, in_rowid, out_sweep, 1, vs); PNP_COUNT_LAUNCH();
}
void vec_from_sweep(const BsrT& A, const double* in_sweep, double* out_rowid, cudaStream_t s, int vs)
{
  permute3_kernel<<<cdiv((size_t)vs * (size_t)A.n, 256), 256, 0, s>>>(A.n, A.vrowid.p, in_sweep, out_rowid, 0, vs); PNP_COUNT_LAUNCH();
}

void bsrt_diaginv(const BsrT& A, double* dinv, cudaStream_t s)
{
  diaginv_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.prd.p, A.val.p, dinv);
  PNP_COUNT_LAUNCH();
}

namespace {
// one thread per stored block: copy its 7 structural entries into the 7-value layout (same [L | D | U] component-major
// parts, 7 components each); raise the flag if entry (1,2) or (2,1) is not zero (a general BSR(3) matrix)
__global__ void pack7_kernel(int nnzb, const int4* __restrict__ rd, const int* __restrict__ brow, const double* __restrict__ val,
                             double* __restrict__ val7, int* __restrict__ general)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int4 d = rd[brow[b]];
  const int s = d.x, len = d.y, nl = d.z, k = b - s;
  const bool lo = k < nl;
  const int st = lo ? nl : len - nl;
  const double* v = val + (9 * (size_t)s + (size_t)(lo ? k : 8 * nl + k));
  double* w = val7 + (7 * (size_t)s + (size_t)(lo ? k : 6 * nl + k));
  if (v[5 * (size_t)st] != 0.0 || v[7 * (size_t)st] != 0.0) *general = 1;
  w[0] = v[0]; w[st] = v[st]; w[2 * (size_t)st] = v[2 * (size_t)st];
  w[3 * (size_t)st] = v[3 * (size_t)st]; w[4 * (size_t)st] = v[4 * (size_t)st];
  w[5 * (size_t)st] = v[6 * (size_t)st]; w[6 * (size_t)st] = v[8 * (size_t)st];
}
}

// policy: PNPB200_PACK7=0 keeps 9 values everywhere; not in low-memory mode (Handle::share_scratch: 320^3 has no room
// for a second copy of the values); the one-CTA kernels of the smallest levels keep FASP's 9
bool bsrt_pack7_wanted(const Handle& h, const BsrT& A)
{
  static const bool on = !(getenv("PNPB200_PACK7") && atoi(getenv("PNPB200_PACK7")) == 0);
  return on && h.share_scratch != 1 && VS == 4 && A.n > SMALL_LEVEL_ROWS && A.nnzb >= 1 && A.rd.p && A.brow.p;
}

// End of the set-up: 7-value copies of the levels that the row kernels sweep. A level that was coarsened has its copy
// already (written by the strength pass, which reads every value anyway); the others are packed here.
void bsrt_pack7(Handle& h, Hierarchy& H)
{
  cudaStream_t s = h.stream;
  const int nl = (int)H.lv.size();
  std::vector<int> tried;
  for (int l = 0; l < nl; ++l) {
    BsrT& A = H.lv[l]->A;
    A.packed7 = false;
    if (!bsrt_pack7_wanted(h, A)) continue;
    if (l == nl - 1 && (H.coarse_dense || H.coarse_rep)) continue;         // never swept
    if (!A.pack7_fresh) {
      A.val7.alloc(7 * (size_t)A.nnzb, s);
      A.pack_flag.alloc(1, s); A.pack_flag.zero();
      pack7_kernel<<<cdiv(A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.rd.p, A.brow.p, A.val.p, A.val7.p, A.pack_flag.p); PNP_COUNT_LAUNCH();
    }
    tried.push_back(l);
  }
  for (int l : tried) {
    int general = 1;
    H.lv[l]->A.pack_flag.download(&general, 1);
    H.lv[l]->A.packed7 = general == 0;
    H.lv[l]->A.pack7_fresh = false;
  }
}

void bsrt_from_fasp(const double* d_val_fasp, BsrT& A, cudaStream_t s)
{
  A.packed7 = false;
  relayout_kernel<true><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.dslot.p, A.ja.p, A.brow.p, d_val_fasp, A.val.p);
  PNP_COUNT_LAUNCH();
}

void bsrt_to_fasp(const BsrT& A, double* d_val_fasp, cudaStream_t s)
{
  relayout_kernel<false><<<cdiv(9 * (size_t)A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.dslot.p, A.ja.p, A.brow.p, A.val.p, d_val_fasp);
  PNP_COUNT_LAUNCH();
}

void bsrt_export_ja(const BsrT& A, int* d_ja_nat, cudaStream_t s)
{
  export_ja_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, A.rs.p, A.ja.p, d_ja_nat);
  PNP_COUNT_LAUNCH();
}

namespace {
// mode 0: plain, 1: forward from a zero initial guess, 2: with delta output
void sweep_impl(const Level& L, const double* b, double* x, double* dl, bool reverse, int mode, cudaStream_t s, int bs)
{
  const BsrT& A = L.A;
  if (A.n <= SMALL_LEVEL_ROWS && L.ngroups > 0 && small_level_smem(A.n, A.nnzb) <= SMALL_SMEM_BYTES) {
    static bool configured[64] = {false};        // the attribute is per device
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      PNP_CUDA(cudaFuncSetAttribute(mcgs_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_BYTES));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    mcgs_smem_kernel<<<1, 1024, small_level_smem(A.n, A.nnzb), s>>>(A.n, A.nnzb, L.ngroups, reverse ? 1 : 0, mode, L.d_color_ptr.p,
                                                                  A.prd.p, A.jpos.p, A.val.p, L.dinv.p, b, x, dl, bs);
    PNP_COUNT_LAUNCH();
    return;
  }
  if (A.n <= SMALL_LEVEL_ROWS && L.ngroups > 0) {
    mcgs_small_kernel<<<1, 1024, 0, s>>>(L.ngroups, reverse ? 1 : 0, mode, L.d_color_ptr.p, A.prd.p, A.jpos.p, A.val.p,
                                        L.dinv.p, b, x, dl, bs);
    PNP_COUNT_LAUNCH();
    return;
  }
  for (int cc = 0; cc < L.ngroups; ++cc) {
    const int c = reverse ? L.ngroups - 1 - cc : cc;
    const int cnt = L.color_ptr[c + 1] - L.color_ptr[c];
    if (cnt == 0) continue;
    if (gs_stream_colour(L, L.color_ptr[c], L.color_ptr[c + 1], b, x, dl, mode, bs, s)) continue;   // gs_stream.cu
    const int4* rd = A.prd.p + L.color_ptr[c];
    const double* dinv = L.dinv.p + 9 * (size_t)L.color_ptr[c];      // sweep order: this colour's piece
    if (mode == 1) launch_rows<M_GS0, PNPB200_GS_CTAS, false>(cnt, cnt, rd, A.jpos.p, A, dinv, b, x, x, nullptr, s, bs, VS);
    else if (mode == 2) launch_rows<M_GSD, PNPB200_GS_CTAS, false>(cnt, cnt, rd, A.jpos.p, A, dinv, b, x, x, dl, s, bs, VS);
    else launch_rows<M_GS, PNPB200_GS_CTAS, false>(cnt, cnt, rd, A.jpos.p, A, dinv, b, x, x, nullptr, s, bs, VS);
  }
}
} // namespace

void mcgs_sweep(const Level& L, const double* b, double* x, bool reverse, cudaStream_t s, int bs) { sweep_impl(L, b, x, nullptr, reverse, 0, s, bs); }
void mcgs_sweep_from_zero(const Level& L, const double* b, double* x, cudaStream_t s, int bs) { sweep_impl(L, b, x, nullptr, false, 1, s, bs); }
void mcgs_sweep_back_delta(const Level& L, const double* b, double* x, double* delta, cudaStream_t s, int bs) { sweep_impl(L, b, x, delta, true, 2, s, bs); }

// Colour the natural graph (A.ia, ja_nat), then lay the matrix out colour by colour.
// `ws` (optional) supplies the temporaries during AMG setup; at create time they come from the pool.
void color_and_layout(Level& L, const int* ja_nat, Workspace* ws, cudaStream_t s)
{
  BsrT& A = L.A;
  const int n = A.n;
  DevBuf<int> t_color, t_rem, t_rows, t_cout, t_len, t_start, t_hist, t_newc, t_vkey, t_vkout, t_tent; DevBuf<unsigned char> t_cand, t_cub, t_cub2;
  auto geti = [&](DevBuf<int>& own, size_t cnt) -> int* { if (ws) return ws->get<int>(cnt); own.alloc(cnt, s); return own.p; };
  int* color = geti(t_color, n);
  int* rem = geti(t_rem, 2);                       // [0] remaining, [1] overflow
  unsigned char* cand;
  if (ws) cand = ws->get<unsigned char>(n); else { t_cand.alloc(n, s); cand = t_cand.p; }
  PNP_CUDA(cudaMemsetAsync(color, 0xff, (size_t)n * sizeof(int), s));
  PNP_CUDA(cudaMemsetAsync(rem, 0, 2 * sizeof(int), s));
  const bool preset = L.preset_ncolors > 0 && L.preset_color.p && L.preset_color.n >= (size_t)n;
  if (preset) PNP_CUDA(cudaMemcpyAsync(color, L.preset_color.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  int hrem = preset ? 0 : n, rounds = 0;
  int* tent = preset ? nullptr : geti(t_tent, n);
  while (hrem > 0) {
    PNP_CUDA(cudaMemsetAsync(rem, 0, sizeof(int), s));
    color_propose_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, ja_nat, color, tent, rem + 1); PNP_COUNT_LAUNCH();
    color_resolve_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, ja_nat, color, tent, cand, rem); PNP_COUNT_LAUNCH();
    color_commit_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, tent, cand, color); PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaMemcpyAsync(&hrem, rem, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (++rounds > 100000) throw Error(ERROR_UNKNOWN, "colouring did not terminate");
  }
  int hh2[2];
  PNP_CUDA(cudaMemcpyAsync(hh2, rem, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if (hh2[1]) throw Error(ERROR_DATA_STRUCTURE, "matrix graph needs more than 62 colours");
  // iterated greedy passes (reverse class order, then forward): fewer, fuller colour classes
  if (!preset) {
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
  // physical row starts, then the physical column array in [L | D | U] order per row
  // (color[] still holds the sweep key of every row; ghost rows carry the largest key)
  int* len = geti(t_len, n + 1);
  A.pia.alloc(n + 1, s);
  int* start = A.pia.p;
  rowlen_by_position_kernel<<<cdiv(n + 1, 256), 256, 0, s>>>(n, A.ia.p, A.rowid.p, len); PNP_COUNT_LAUNCH();
  exclusive_scan_int(len, start, n + 1, s);
  A.rs.alloc(n, s); A.rd.alloc(n, s); A.prd.alloc(n, s); A.pos.alloc(n, s); A.vrowid.alloc(n, s); A.vrd.alloc(n, s);
  scatter_rs_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.rowid.p, start, A.rs.p); PNP_COUNT_LAUNCH();
  {
    // vector order of the solver (BsrT::vrowid): rows sorted by (tile, colour), ghost rows last
    static const int vec_tile = getenv("PNPB200_VEC_TILE") ? atoi(getenv("PNPB200_VEC_TILE")) : 4096;
    if (vec_tile <= 0 || ntiles > 1) {
      PNP_CUDA(cudaMemcpyAsync(A.vrowid.p, A.rowid.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));   // = matrix order
    } else {
      const int nt = (n + vec_tile - 1) / vec_tile;
      const int ghost_key = nt * 64;
      int vbits = 1; while ((1 << vbits) <= ghost_key) ++vbits;
      int* vkey = geti(t_vkey, n);
      int* vkey_out = geti(t_vkout, n);
      vector_key_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, n_own, vec_tile, ghost_key, color, vkey); PNP_COUNT_LAUNCH();
      size_t vb = 0;
      PNP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, vb, vkey, vkey_out, rows_in, A.vrowid.p, n, 0, vbits, s));
      unsigned char* vtmp;
      if (ws) vtmp = ws->get<unsigned char>(vb); else { t_cub2.alloc(vb, s); vtmp = t_cub2.p; }
      PNP_CUDA(cub::DeviceRadixSort::SortPairs(vtmp, vb, vkey, vkey_out, rows_in, A.vrowid.p, n, 0, vbits, s));
      PNP_COUNT_LAUNCH();
    }
    invert_perm_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.vrowid.p, A.pos.p); PNP_COUNT_LAUNCH();
  }
  A.ja.alloc(A.nnzb, s);
  A.brow.alloc(A.nnzb, s); A.dslot.alloc(n, s); A.tslot.alloc(A.nnzb, s);
  PNP_CUDA(cudaMemsetAsync(rem, 0, sizeof(int), s));
  layout_row_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, color, ja_nat, A.ja.p, A.brow.p, A.dslot.p, A.rd.p, rem);
  PNP_COUNT_LAUNCH();
  physical_descriptor_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.rowid.p, A.pos.p, A.rd.p, A.prd.p); PNP_COUNT_LAUNCH();
  vector_descriptor_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.vrowid.p, A.rd.p, A.vrd.p); PNP_COUNT_LAUNCH();
  A.maxlen = max_row_length(A, s);
  int nodiag = 0;
  PNP_CUDA(cudaMemcpyAsync(&nodiag, rem, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if (nodiag) throw Error(ERROR_DATA_STRUCTURE, "matrix row without a diagonal block");
  tslot_kernel<<<cdiv(A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.ia.p, A.rs.p, A.dslot.p, A.ja.p, A.brow.p, A.tslot.p); PNP_COUNT_LAUNCH();
  // sweep-space index arrays of the solver kernels: columns and halo send lists as sweep positions
  A.jpos.alloc(A.nnzb, s);
  gather_pos_kernel<<<cdiv(A.nnzb, 256), 256, 0, s>>>(A.nnzb, A.pos.p, A.ja.p, A.jpos.p); PNP_COUNT_LAUNCH();
  if (L.halo.active()) {
    const int nsend = L.halo.send_ptr.back();
    L.halo.send_pos.alloc(nsend ? nsend : 1, s);
    if (nsend > 0) { gather_pos_kernel<<<cdiv(nsend, 256), 256, 0, s>>>(nsend, A.pos.p, L.halo.send_idx.p, L.halo.send_pos.p); PNP_COUNT_LAUNCH(); }
  }
  PNP_LAUNCH_CHECK();
}

} // namespace pnpb200
