// pnp-b200 — unsmoothed-aggregation AMG for BSR(3) on the device.
// Stands in for FASP's fasp_amg_setup_ua_bsr + fasp_solver_mgcycle_bsr [EXT], called through
// fasp_solver_dbsr_krylov_amg at benchmarks/pnp_exact_solution/linear_pnp.cpp:101.
//   strength : a_ij^2 >= theta^2 |a_ii a_jj| on the L-inf condensed matrix (VMB criterion),
//              theta halves per level when tentative_smooth > 0 (bsr.dat:97-99)
//   aggregates: parallel distance-2 maximal independent set of the symmetrised strength graph
//              (roots), root + strong neighbours, leftovers join their strongest aggregated
//              neighbour — the parallel counterpart of VMB's three sequential steps;
//              rows without off-diagonal coupling (Dirichlet rows) stay out of the coarse grid
//   P        : boolean block prolongator, R = P^T, A_c = R A P by sorting blocks on their
//              coarse (I,J) key and a segmented sum (atomic-free, deterministic)
//   cycle    : V(nu1,nu2), multicolour block Gauss-Seidel (colours ascending / descending),
//              exact dense solve on the coarsest level (explicit inverse, one GEMV per cycle)
#include "internal.hpp"
#include <cub/cub.cuh>
#include <cmath>
#include <algorithm>

namespace pnpb200 {

namespace {

typedef unsigned long long u64;

// block infinity-norms for the strength test. PACK: the pass holds every value of the level anyway, so it also writes
// the 7-value copy the solve's matrix passes read (bsr_ops.cu: bsrt_pack7; flag raised by a block with (1,2) / (2,1) != 0)
template <bool PACK>
__global__ void condense_kernel(int nnzb, const int4* __restrict__ rd, const int* __restrict__ brow,
                                const double* __restrict__ val, double* __restrict__ pp, double* __restrict__ val7, int* __restrict__ general)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int4 d = rd[brow[b]];
  int st;
  const double* v = val + bsrt_off(d.x, d.y, d.z, b - d.x, st);
  double e[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) e[t] = v[(size_t)t * st];
  double m = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) m = fmax(m, fabs(e[3 * r]) + fabs(e[3 * r + 1]) + fabs(e[3 * r + 2]));
  pp[b] = m;
  if (PACK) {
    const int k = b - d.x;
    double* w = val7 + (7 * (size_t)d.x + (size_t)(k < d.z ? k : 6 * d.z + k));
    if (e[5] != 0.0 || e[7] != 0.0) *general = 1;
    w[0] = e[0]; w[st] = e[1]; w[2 * (size_t)st] = e[2]; w[3 * (size_t)st] = e[3]; w[4 * (size_t)st] = e[4];
    w[5 * (size_t)st] = e[6]; w[6 * (size_t)st] = e[8];
  }
}

__global__ void isolated_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                const int* __restrict__ ja, const double* __restrict__ pp, unsigned char* __restrict__ iso)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool any = false;
  for (int p = rs[i], e = rs[i] + ia[i + 1] - ia[i]; p < e; ++p) if (ja[p] != i && pp[p] != 0.0) { any = true; break; }
  iso[i] = any ? 0 : 1;
}

__global__ void strength_kernel(int nnzb, const int* __restrict__ ja, const int* __restrict__ brow,
                                const int* __restrict__ dslot, const int* __restrict__ tslot,
                                const double* __restrict__ pp, const unsigned char* __restrict__ iso,
                                double theta2, unsigned char* __restrict__ strong)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int i = brow[b], j = ja[b];
  unsigned char s = 0;
  if (i != j && !iso[i] && !iso[j]) {
    const double dd = theta2 * fabs(pp[dslot[i]] * pp[dslot[j]]);
    const double a = pp[b];
    const int t = tslot[b];
    const double at = t >= 0 ? pp[t] : 0.0;
    s = ((a > 0.0 && a * a >= dd) || (at > 0.0 && at * at >= dd)) ? 1 : 0;
  }
  strong[b] = s;
}

// a non-isolated row without any strong neighbour keeps its largest coupling (select, then
// apply, so the outcome does not depend on thread timing)
__global__ void force_select_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                    const int* __restrict__ ja, const double* __restrict__ pp, const unsigned char* __restrict__ iso,
                                    const unsigned char* __restrict__ strong, int* __restrict__ sel)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int best = -1; double bv = 0.0;
  if (!iso[i])
    for (int p = rs[i], e = rs[i] + ia[i + 1] - ia[i]; p < e; ++p) {
      if (strong[p]) { best = -1; break; }
      const int j = ja[p];
      if (j == i || iso[j]) continue;
      if (pp[p] > bv) { bv = pp[p]; best = p; }
    }
  sel[i] = best;
}

__global__ void force_apply_kernel(int n, const int* __restrict__ tslot, const int* __restrict__ sel,
                                   unsigned char* __restrict__ strong)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = sel[i];
  if (b >= 0) { strong[b] = 1; if (tslot[b] >= 0) strong[tslot[b]] = 1; }
}

__device__ __forceinline__ u64 mis_key(int rank, int i)
{
  return ((u64)rank << 62) | ((u64)(hash32((unsigned)i) & 0x3fffffffu) << 32) | (u64)(unsigned)i;
}

// state: 0 undecided, 1 root, 2 covered / excluded
__global__ void mis_init_kernel(int n, const unsigned char* __restrict__ iso, int* __restrict__ state)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) state[i] = iso[i] ? 2 : 0;
}

// Only rows that can still influence a decision are visited: the second propagation (FIRST =
// false) feeds mis_decide_kernel, which looks at undecided rows only; the first one must be current
// for the strong neighbours of undecided rows (need[] is set by mis_decide_kernel for every row it
// leaves undecided and for that row's strong neighbours). After three rounds < 10 % of the rows
// are active, so the 16 rounds of the 256^3 fine level cost about as much as four full ones.
template <bool FIRST>
__global__ void mis_prop_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                const int* __restrict__ ja, const unsigned char* __restrict__ strong, const int* __restrict__ state,
                                const unsigned char* __restrict__ need, const u64* __restrict__ tin, u64* __restrict__ tout)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (FIRST ? !need[i] : state[i] != 0) return;
  auto key = [&](int j) -> u64 {
    if (!FIRST) return tin[j];
    const int st = state[j];
    return mis_key(st == 1 ? 2 : st == 0 ? 1 : 0, j);
  };
  u64 m = key(i);
  for (int p = rs[i], e = rs[i] + ia[i + 1] - ia[i]; p < e; ++p)
    if (strong[p]) { const u64 k = key(ja[p]); m = k > m ? k : m; }
  tout[i] = m;
}

__global__ void mis_decide_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                  const int* __restrict__ ja, const unsigned char* __restrict__ strong,
                                  const u64* __restrict__ t2, int* __restrict__ state,
                                  unsigned char* __restrict__ need_next, int* __restrict__ undecided)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || state[i] != 0) return;
  const u64 m = t2[i];
  if ((int)(m & 0xffffffffu) == i) state[i] = 1;
  else if ((m >> 62) == 2) state[i] = 2;
  else {
    atomicAdd(undecided, 1);
    need_next[i] = 1;                 // benign races: every writer stores 1
    for (int p = rs[i], e = rs[i] + ia[i + 1] - ia[i]; p < e; ++p) if (strong[p]) need_next[ja[p]] = 1;
  }
}

__global__ void root_flag_kernel(int n, const int* __restrict__ state, int* __restrict__ flag)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = state[i] == 1 ? 1 : 0;
  if (i == n) flag[n] = 0;
}

// pass 0: roots take their id; pass >=1: unassigned rows join the strongest strong neighbour
// that was assigned in an EARLIER pass (agg_in snapshot) — deterministic
__global__ void agg_root_kernel(int n, const int* __restrict__ state, const int* __restrict__ rootid,
                                const unsigned char* __restrict__ iso, int* __restrict__ agg)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  agg[i] = state[i] == 1 ? rootid[i] : (iso[i] ? -1 : -2);
}

__global__ void agg_join_kernel(int n, const int* __restrict__ ia, const int* __restrict__ rs,
                                const int* __restrict__ ja, const unsigned char* __restrict__ strong, const double* __restrict__ pp,
                                const int* __restrict__ agg_in, int* __restrict__ agg_out, int* __restrict__ left)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = agg_in[i];
  if (a == -2) {
    double bv = -1.0;
    for (int p = rs[i], e = rs[i] + ia[i + 1] - ia[i]; p < e; ++p) {
      if (!strong[p]) continue;
      const int aj = agg_in[ja[p]];
      if (aj >= 0 && pp[p] > bv) { bv = pp[p]; a = aj; }
    }
    if (a == -2) atomicAdd(left, 1);
  }
  agg_out[i] = a;
}

__global__ void agg_finalize_kernel(int n, int* __restrict__ agg)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && agg[i] < -1) agg[i] = -1;
}

__global__ void gather_int_kernel(int n, const int* __restrict__ idx, const int* __restrict__ v, int* __restrict__ out)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = v[idx[e]];
}

__global__ void agg_count_kernel(int n, const int* __restrict__ agg, int* __restrict__ cnt)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && agg[i] >= 0) atomicAdd(&cnt[agg[i]], 1);
}

__global__ void iota_kernel(int n, int* __restrict__ a)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

__global__ void agg_key_kernel(int n, const int* __restrict__ agg, unsigned* __restrict__ key)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = agg[i] >= 0 ? (unsigned)agg[i] : 0xffffffffu;
}

// ---- Galerkin product with the boolean prolongator ------------------------------------------
// key of fine block b = (coarse row, coarse column); only coarse rows owned by this rank are formed
// here. Entries nnzb .. nnzb+nghost-1 are the identity diagonals of the coarse ghost rows (idx = -1).
__global__ void rap_key_kernel(int nnzb, int nghost, int nown_c, const int* __restrict__ ja, const int* __restrict__ brow,
                               const int* __restrict__ agg, u64 nagg, u64* __restrict__ key, int* __restrict__ idx)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb + nghost) return;
  if (b >= nnzb) { const u64 G = (u64)(nown_c + (b - nnzb)); key[b] = G * nagg + G; idx[b] = -1; return; }
  const int I = agg[brow[b]], J = agg[ja[b]];
  key[b] = (I >= 0 && I < nown_c && J >= 0) ? (u64)I * nagg + (u64)J : ~0ull;
  idx[b] = b;
}

__global__ void rap_head_kernel(int nnzb, const u64* __restrict__ key, int* __restrict__ head)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nnzb) return;
  if (p == nnzb) { head[p] = 0; return; }
  const u64 k = key[p];
  head[p] = (k != ~0ull && (p == 0 || key[p - 1] != k)) ? 1 : 0;
}

__global__ void rap_structure_kernel(int nnzb, const u64* __restrict__ key, const int* __restrict__ head,
                                     const int* __restrict__ slot, u64 nagg, int* __restrict__ cja,
                                     int* __restrict__ crow, int* __restrict__ seg, int* __restrict__ rowcnt,
                                     int* __restrict__ nvalid)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnzb) return;
  const u64 k = key[p];
  if (k == ~0ull) { if (p == 0 || key[p - 1] != ~0ull) *nvalid = p; return; }
  if (p == nnzb - 1) *nvalid = nnzb;
  if (head[p]) {
    const int s = slot[p];
    cja[s] = (int)(k % nagg);
    crow[s] = (int)(k / nagg);
    seg[s] = p;
    atomicAdd(&rowcnt[(int)(k / nagg)], 1);
  }
}

// natural coarse block s (row I, column cja_nat[s]) -> its physical block in the [L | D | U] row; its fine segment
__global__ void rap_remap_kernel(int cnnzb, const int* __restrict__ crow, const int* __restrict__ cja_nat,
                                 const int* __restrict__ cia, const int* __restrict__ crs, const int* __restrict__ cja,
                                 const int* __restrict__ seg, int* __restrict__ beg, int* __restrict__ end)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cnnzb) return;
  const int I = crow[s], J = cja_nat[s];
  int b = crs[I];
  const int e = b + (cia[I + 1] - cia[I]);
  while (b < e && cja[b] != J) ++b;
  beg[b] = seg[s]; end[b] = seg[s + 1];
}

__global__ void rap_numeric_kernel(int cnnzb, const int4* __restrict__ crd, const int* __restrict__ cbrow,
                                   const int* __restrict__ beg, const int* __restrict__ end,
                                   const int* __restrict__ perm, const int4* __restrict__ frd,
                                   const int* __restrict__ fbrow, const double* __restrict__ fval,
                                   double* __restrict__ cval)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * (size_t)cnnzb) return;
  const int s = (int)(e / 9), t = (int)(e % 9);          // s = physical coarse block
  double acc = 0.0;
  for (int p = beg[s]; p < end[s]; ++p) {
    const int b = perm[p];
    if (b < 0) { acc += (t == 0 || t == 4 || t == 8) ? 1.0 : 0.0; continue; }   // identity row of a coarse ghost
    const int4 d = frd[fbrow[b]];                        // (start, length, nl, row) of the fine row
    int st;
    const size_t o = bsrt_off(d.x, d.y, d.z, b - d.x, st);
    acc += fval[o + (size_t)t * st];
  }
  const int4 c = crd[cbrow[s]];
  int cst;
  const size_t co = bsrt_off(c.x, c.y, c.z, s - c.x, cst);
  cval[co + (size_t)t * cst] = acc;
}

// ---- Galerkin product, row-wise form ------------------------------------------------------------
// A half-warp owns one coarse row I = one aggregate. With piecewise-constant P the coarse row is the
// sum of the aggregate's fine rows with the columns renamed j -> agg[j], so the fine matrix is read
// exactly once, as whole contiguous rows (lane k = block k of the fine row, like the matvec). No
// global sort of the 253 M (I,J) keys of the 256^3 fine level (22 ms + 15 ms of scattered gathers by
// its permutation): measured 37 -> 9 ms for the first coarsening step.
//   symbolic: the distinct coarse columns of the row go through a per-half-warp hash set in shared
//             memory (count pass, then fill pass + rank sort -> ascending natural order);
//   numeric : per-half-warp accumulators acc[slot][9] in shared memory, slot = position of the
//             coarse column in the row's physical [L | D | U] order. Lanes of a fine row that map to
//             the same coarse column add one after the other in lane order (match_any groups), fine
//             rows in ascending order: the sum order is fixed, the result deterministic.
// Rows with more than RAPW_MAXCOLS distinct columns are not handled here: the level then takes the
// sort-based path below (nothing in the BASELINE configs does).
constexpr int RAPW_HASH = 256;          // hash-set slots per half-warp (symbolic)
constexpr int RAPW_MAXCOLS = 192;       // longest coarse row of the row-wise path

constexpr int RAPW_TMPW = 64;           // rows up to this length hand their column set from the count pass to the fill pass

// FILL = false: count pass. The distinct coarse columns of row I are collected in the hash set, the
//   count goes to rowcnt[I]; sets of <= RAPW_TMPW columns are also parked (unsorted) in tmpcols so
//   that the fill pass does not hash the row a second time.
// FILL = true : fill pass. Column set from tmpcols (or hashed again for the few long rows), rank
//   sort -> ascending natural order into cja[cia[I] ..].
template <bool FILL>
__global__ void __launch_bounds__(256)
rapw_symbolic_kernel(int nown_c, int nghost_c, const int* __restrict__ agg_ptr, const int* __restrict__ agg_rows,
                     const int4* __restrict__ rd, const int* __restrict__ ja, const int* __restrict__ agg,
                     int* __restrict__ rowcnt, int* __restrict__ maxlen, int* __restrict__ tmpcols,
                     const int* __restrict__ cia, int* __restrict__ cja)
{
  __shared__ int tab[16][RAPW_HASH];
  __shared__ int list[16][RAPW_HASH];
  __shared__ int cnt[16];
  const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const int I = blockIdx.x * 16 + hw;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);
  if (I >= nown_c + nghost_c) return;                       // whole half-warp
  if (I >= nown_c) {                                        // coarse ghost row: identity diagonal
    if (lane == 0) { if (FILL) cja[cia[I]] = I; else rowcnt[I] = 1; }
    return;
  }
  int n = FILL ? cia[I + 1] - cia[I] : 0;
  if (FILL && n <= RAPW_TMPW) {
    for (int e = lane; e < n; e += 16) list[hw][e] = tmpcols[(size_t)I * RAPW_TMPW + e];
  } else {
    for (int e = lane; e < RAPW_HASH; e += 16) tab[hw][e] = -1;
    if (lane == 0) cnt[hw] = 0;
    __syncwarp(hmask);
    for (int m = agg_ptr[I]; m < agg_ptr[I + 1]; ++m) {
      const int4 d = rd[agg_rows[m]];
      for (int k = lane; k < d.y; k += 16) {
        const int J = agg[ja[d.x + k]];
        if (J < 0) continue;
        unsigned h = hash32((unsigned)J) & (RAPW_HASH - 1);
        for (int probe = 0; probe < RAPW_HASH; ++probe) {
          const int old = atomicCAS(&tab[hw][h], -1, J);
          if (old == -1) { list[hw][atomicAdd(&cnt[hw], 1)] = J; break; }      // first sighting: append to the dense list
          if (old == J) break;
          h = (h + 1) & (RAPW_HASH - 1);
        }
      }
    }
    __syncwarp(hmask);
    n = cnt[hw];
  }
  if (!FILL) {
    if (lane == 0) { rowcnt[I] = n; atomicMax(maxlen, n); }
    if (n <= RAPW_TMPW) for (int e = lane; e < n; e += 16) tmpcols[(size_t)I * RAPW_TMPW + e] = list[hw][e];
    return;
  }
  __syncwarp(hmask);
  // ascending natural order: rank of an entry = number of smaller entries
  const int base = cia[I];
  for (int e = lane; e < n; e += 16) {
    const int J = list[hw][e];
    int rank = 0;
    for (int q = 0; q < n; ++q) rank += list[hw][q] < J ? 1 : 0;
    cja[base + rank] = J;
  }
}

template <int CAP>
__global__ void __launch_bounds__(16 * (CAP <= 32 ? 12 : CAP <= 64 ? 6 : 2))
rapw_numeric_kernel(int nown_c, int nghost_c, const int* __restrict__ agg_ptr, const int* __restrict__ agg_rows,
                    const int4* __restrict__ frd, const int* __restrict__ fja, const int* __restrict__ agg,
                    const double* __restrict__ fval, const int4* __restrict__ crd, const int* __restrict__ cja,
                    double* __restrict__ cval)
{
  constexpr int HW = CAP <= 32 ? 12 : CAP <= 64 ? 6 : 2;     // half-warps per CTA (static shared memory <= 48 KB)
  constexpr int HT = CAP <= 32 ? 64 : CAP <= 64 ? 128 : 512;  // hash map coarse column -> slot (power of two, > CAP)
  constexpr int AS = 10;                                      // doubles per accumulator slot: 9 + pad, so that a slot is five aligned 16-byte words
  __shared__ __align__(16) double acc[HW][CAP * AS];
  __shared__ int2 map[HW][HT];                                // (coarse column, slot), .x = -1: empty
  const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const int I = blockIdx.x * HW + hw;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);
  if (I >= nown_c + nghost_c) return;
  const int4 c = crd[I];                                      // (start, len, nl, row) of the coarse row
  if (I >= nown_c) {                                          // coarse ghost row: identity block
    if (lane < 9) { int st; const size_t o = bsrt_off(c.x, c.y, c.z, c.z, st); cval[o + (size_t)lane * st] = (lane == 0 || lane == 4 || lane == 8) ? 1.0 : 0.0; }
    return;
  }
  const int len = c.y;
  for (int e = lane; e < HT; e += 16) map[hw][e] = make_int2(-1, -1);
  for (int e = lane; e < AS * len; e += 16) acc[hw][e] = 0.0;
  __syncwarp(hmask);
  for (int q = lane; q < len; q += 16) {                      // slot = position in the row's physical [L | D | U] order
    const int J = cja[c.x + q];
    unsigned h = hash32((unsigned)J) & (HT - 1);
    while (atomicCAS(&map[hw][h].x, -1, J) != -1) h = (h + 1) & (HT - 1);      // columns of a row are distinct
    map[hw][h].y = q;
  }
  __syncwarp(hmask);
  // The (fine row, 16-block pass) items of the aggregate are walked with a one-item look-ahead: the
  // loads of the next item (descriptor -> columns -> aggregate ids, and the 9 values of the lane's
  // block) are in flight while the current one is added up. m, ps, d, dn are uniform over the half-warp.
  int m = agg_ptr[I], ps = 0;
  const int mend = agg_ptr[I + 1];
  const int4 none = make_int4(0, 0, 0, 0);
  int4 d = m < mend ? frd[agg_rows[m]] : none;
  int4 dn = m + 1 < mend ? frd[agg_rows[m + 1]] : none;
  auto fetch = [&](int& J, double (&v)[9]) {       // item (m, ps) into registers, then step to the next item
    J = -1;
    const int k = ps * 16 + lane;
    if (k < d.y) {
      J = agg[fja[d.x + k]];
      int st;
      const size_t o = bsrt_off(d.x, d.y, d.z, k, st);
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = fval[o + (size_t)t * st];
    }
    if ((ps + 1) * 16 < d.y) ++ps;
    else { ++m; ps = 0; d = dn; dn = m + 1 < mend ? frd[agg_rows[m + 1]] : none; }
  };
  auto accumulate = [&](int J, double (&v)[9]) {
    int slot = -1;
    if (J >= 0) {
      unsigned h = hash32((unsigned)J) & (HT - 1);
      int2 e = map[hw][h];
      while (e.x != J && e.x != -1) { h = (h + 1) & (HT - 1); e = map[hw][h]; }
      slot = e.x == J ? e.y : -1;
    }
    // lanes with the same slot add in lane order, one member of every group per round (tried: summing a
    // group in registers with shuffles first — 17 instead of 10 ms at 256^3, the shuffles cost more than
    // the shared-memory updates they save)
    const unsigned peers = __match_any_sync(hmask, slot) & hmask;
    const int rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    int rounds = slot >= 0 ? __popc(peers) : 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { const int t = __shfl_xor_sync(hmask, rounds, o); rounds = t > rounds ? t : rounds; }
    for (int r = 0; r < rounds; ++r) {
      if (slot >= 0 && rank == r) {
        double2* a2 = reinterpret_cast<double2*>(&acc[hw][slot * AS]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { double2 w = a2[t]; w.x += v[2 * t]; w.y += v[2 * t + 1]; a2[t] = w; }
        acc[hw][slot * AS + 8] += v[8];
      }
      __syncwarp(hmask);
    }
  };
  bool have = m < mend;
  int Jc = -1; double vc[9];
  if (have) fetch(Jc, vc);
  while (have) {
    const bool more = m < mend;
    int Jn = -1; double vn[9];
    if (more) fetch(Jn, vn);
    accumulate(Jc, vc);
    have = more; Jc = Jn;
#pragma unroll
    for (int t = 0; t < 9; ++t) vc[t] = vn[t];
  }
  // coarse row out, BSR-T layout
  for (int q = lane; q < len; q += 16) {
    int st;
    const size_t o = bsrt_off(c.x, c.y, c.z, q, st);
#pragma unroll
    for (int t = 0; t < 9; ++t) cval[o + (size_t)t * st] = acc[hw][q * AS + t];
  }
}

// ---- grid transfer (sweep space, internal.hpp) --------------------------------------------------
// coarse position P <- sum of the fine positions of aggregate c_rowid[P] (coarsest level: P itself).
// Fine vectors hold VS doubles per row (sweep space); the coarse vector holds CS: VS, or 3 when the
// next level is the coarsest with the dense solve (row-id order, packed).
template <int CS>
__global__ void restrict_kernel(int nagg, const int* __restrict__ c_rowid, const int* __restrict__ agg_ptr,
                                const int* __restrict__ agg_rows_pos, const double* __restrict__ r, double* __restrict__ bc)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= CS * nagg) return;
  const int P = e / CS, k = e % CS;
  if (k == 3) { bc[e] = 0.0; return; }
  const int I = c_rowid ? c_rowid[P] : P;
  double a = 0.0;
  for (int p = agg_ptr[I]; p < agg_ptr[I + 1]; ++p) a += r[(size_t)VS * (size_t)agg_rows_pos[p] + k];
  bc[e] = a;
}

template <int CS>
__global__ void prolong_kernel(int n, const int* __restrict__ agg_pos, const double* __restrict__ ec,
                               double* __restrict__ x)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= VS * n) return;
  const int k = e % VS;
  if (k == 3) return;
  const int a = agg_pos[e / VS];
  if (a >= 0) x[e] += ec[(size_t)CS * (size_t)a + k];
}

// agg_pos[p] = coarse position of the aggregate of row rowid[p]; agg_rows_pos[q] = position of member agg_rows[q]
__global__ void transfer_maps_kernel(int n, const int* __restrict__ rowid, const int* __restrict__ pos,
                                     const int* __restrict__ agg, const int* __restrict__ agg_rows,
                                     const int* __restrict__ c_pos, int* __restrict__ agg_pos, int* __restrict__ agg_rows_pos)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int a = agg[rowid[p]];
  agg_pos[p] = a < 0 ? -1 : (c_pos ? c_pos[a] : a);
  agg_rows_pos[p] = pos[agg_rows[p]];      // entries behind the last aggregate's members are unused rows (agg < 0)
}

// ---- coarsest level: dense inverse -------------------------------------------------------------
__global__ void identity_kernel(int nd, double* __restrict__ D)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (size_t)nd * nd) D[e] = (e / nd == e % nd) ? 1.0 : 0.0;
}

// Gauss-Jordan with partial pivoting, one CTA; A is destroyed, Inv receives A^{-1}
__global__ void __launch_bounds__(1024) dense_invert_kernel(int nd, double* __restrict__ A, double* __restrict__ Inv,
                                                           double* __restrict__ colk, int* __restrict__ singular)
{
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < nd; ++k) {
    // pivot search
    double bv = -1.0; int bi = k;
    for (int r = k + tid; r < nd; r += nt) { const double v = fabs(A[(size_t)r * nd + k]); if (v > bv) { bv = v; bi = r; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      bv = tid < (nt >> 5) ? s_val[tid] : -1.0; bi = tid < (nt >> 5) ? s_idx[tid] : k;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (tid == 0) { s_piv = bi; if (!(bv > 0.0) || !isfinite(bv)) *singular = 1; }
    }
    __syncthreads();
    const int pv = s_piv;
    if (pv != k)
      for (int c = tid; c < nd; c += nt) {
        double t = A[(size_t)k * nd + c]; A[(size_t)k * nd + c] = A[(size_t)pv * nd + c]; A[(size_t)pv * nd + c] = t;
        t = Inv[(size_t)k * nd + c]; Inv[(size_t)k * nd + c] = Inv[(size_t)pv * nd + c]; Inv[(size_t)pv * nd + c] = t;
      }
    __syncthreads();
    const double ip = 1.0 / A[(size_t)k * nd + k];
    __syncthreads();
    for (int c = tid; c < nd; c += nt) { A[(size_t)k * nd + c] *= ip; Inv[(size_t)k * nd + c] *= ip; }
    for (int r = tid; r < nd; r += nt) colk[r] = A[(size_t)r * nd + k];
    __syncthreads();
    for (size_t e = tid; e < (size_t)nd * nd; e += nt) {
      const int r = (int)(e / nd), c = (int)(e % nd);
      if (r == k) continue;
      const double f = colk[r];
      if (f == 0.0) continue;
      if (c >= k) A[e] -= f * A[(size_t)k * nd + c];
      Inv[e] -= f * Inv[(size_t)k * nd + c];
    }
    __syncthreads();
  }
}

// ---- K-cycle (nonlinear AMLI, FASP cycle_type NL_AMLI [EXT]; Notay's 2-step GCR form) ----------
// All scalars stay on the device so that one preconditioner application is a fixed launch
// sequence without host synchronisation.
//   out[q] = <a_q, b_q>, q < NP; single launch, deterministic (last block sums the partials in order)
template <int NP>
__global__ void __launch_bounds__(256)
kdots_kernel(size_t n, const double* __restrict__ a0, const double* __restrict__ b0,
             const double* __restrict__ a1, const double* __restrict__ b1,
             const double* __restrict__ a2, const double* __restrict__ b2,
             double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ out)
{
  __shared__ double sm[32];
  __shared__ bool last;
  double acc[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) acc[q] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    acc[0] += a0[i] * b0[i];
    if (NP > 1) acc[1] += a1[i] * b1[i];
    if (NP > 2) acc[2] += a2[i] * b2[i];
  }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const double r = block_sum(acc[q], sm);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NP + q] = r;
  }
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x < NP) {
      double t = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) t += partial[(size_t)b * NP + threadIdx.x];
      out[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// sc[0] = <v,b>, sc[1] = <v,v>:  rt = b - (sc0/sc1) v
__global__ void kc_step1_kernel(size_t n, const double* __restrict__ sc, const double* __restrict__ b,
                                const double* __restrict__ v, double* __restrict__ rt)
{
  const double a = sc[1] != 0.0 ? sc[0] / sc[1] : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    rt[i] = b[i] - a * v[i];
}

// sc[2] = <w,v>, sc[3] = <w,w>, sc[4] = <w,rt>:  x = (a1 - g*a2/(r1*r2)) c + (a2/r2) d
__global__ void kc_final_kernel(size_t n, const double* __restrict__ sc, const double* __restrict__ c,
                                const double* __restrict__ d, double* __restrict__ x)
{
  const double r1 = sc[1], a1 = r1 != 0.0 ? sc[0] / r1 : 0.0;
  const double g = sc[2];
  const double r2 = r1 != 0.0 ? sc[3] - g * g / r1 : 0.0;
  double k1 = a1, k2 = 0.0;
  if (r2 > 0.0 && r1 != 0.0) { k2 = sc[4] / r2; k1 = a1 - g * k2 / r1; }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = k1 * c[i] + k2 * d[i];
}

template <typename K, typename V>
void sort_pairs(Workspace& ws, K* kin, K* kout, V* vin, V* vout, int n, int end_bit, cudaStream_t s)
{
  size_t bytes = 0;
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
  unsigned char* tmp = ws.get<unsigned char>(bytes);
  PNP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
  PNP_COUNT_LAUNCH();
}

int bits_for(u64 v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

// setup phase timing (print_level >= 3 only): host clock around a stream sync
struct Phase {
  cudaStream_t s; bool on; const char* name; int lvl; double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  Phase(cudaStream_t st, bool o, const char* n, int l) : s(st), on(o), name(n), lvl(l), t0(0) { if (on) { cudaStreamSynchronize(s); t0 = now(); } }
  ~Phase() { if (on) { cudaStreamSynchronize(s); printf("    [setup l%d] %-18s %8.2f ms\n", lvl, name, now() - t0); } }
};
bool g_phase = false; int g_phase_lvl = 0;
#define PHASE(name) Phase phase__(s, g_phase, name, g_phase_lvl)

// symbolic part of one coarsening step: aggregates of L and the structure of the next level
bool coarsen_symbolic(Handle& h, Level& L, Level& C, double theta)
{
  cudaStream_t s = h.stream;
  BsrT& A = L.A;
  const int n = A.n, nnzb = A.nnzb;
  // all large temporaries of this coarsening step live in the handle's workspace
  Workspace& ws = h.ws;
  // (the sort-based fallback of the Galerkin product needs 60 bytes per block; it re-reserves below)
  static const bool rowwise_on = getenv("PNPB200_RAP_ROWWISE") ? atoi(getenv("PNPB200_RAP_ROWWISE")) != 0 : true;
  ws.reserve((size_t)nnzb * (rowwise_on ? 14 : 60) + (size_t)n * 112 + (64u << 20), s);
  ws.reset();
  struct P1 { double* p; } pp; struct P2 { unsigned char* p; } iso, strong; struct P3 { int* p; void zero(cudaStream_t st) { cudaMemsetAsync(p, 0, sizeof(int), st); } } state, cnt;
  struct P4 { u64* p; } t1, t2;
  {
  PHASE("strength");
  pp.p = ws.get<double>(nnzb);
  if (bsrt_pack7_wanted(h, A)) {
    A.val7.alloc(7 * (size_t)nnzb, s);
    A.pack_flag.alloc(1, s); A.pack_flag.zero();
    condense_kernel<true><<<cdiv(nnzb, 256), 256, 0, s>>>(nnzb, A.rd.p, A.brow.p, A.val.p, pp.p, A.val7.p, A.pack_flag.p);
    A.pack7_fresh = true;
  } else condense_kernel<false><<<cdiv(nnzb, 256), 256, 0, s>>>(nnzb, A.rd.p, A.brow.p, A.val.p, pp.p, nullptr, nullptr);
  PNP_COUNT_LAUNCH();
  iso.p = ws.get<unsigned char>(n); strong.p = ws.get<unsigned char>(nnzb);
  isolated_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, pp.p, iso.p); PNP_COUNT_LAUNCH();
  strength_kernel<<<cdiv(nnzb, 256), 256, 0, s>>>(nnzb, A.ja.p, A.brow.p, A.dslot.p, A.tslot.p, pp.p, iso.p,
                                                 theta * theta, strong.p); PNP_COUNT_LAUNCH();
  {
    struct { int* p; } sel; sel.p = ws.get<int>(n);
    force_select_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, pp.p, iso.p, strong.p, sel.p); PNP_COUNT_LAUNCH();
    force_apply_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.tslot.p, sel.p, strong.p); PNP_COUNT_LAUNCH();
  }
  }
  // distance-2 MIS
  int rounds = 0;
  {
  PHASE("mis2");
  state.p = ws.get<int>(n); cnt.p = ws.get<int>(1);
  t1.p = ws.get<u64>(n); t2.p = ws.get<u64>(n);
  unsigned char* need = ws.get<unsigned char>(n); unsigned char* need_next = ws.get<unsigned char>(n);
  PNP_CUDA(cudaMemsetAsync(need, 1, n, s));
  mis_init_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, iso.p, state.p); PNP_COUNT_LAUNCH();
  int und = 1;
  while (und > 0) {
    cnt.zero(s);
    PNP_CUDA(cudaMemsetAsync(need_next, 0, n, s));
    mis_prop_kernel<true><<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, strong.p, state.p, need, nullptr, t1.p); PNP_COUNT_LAUNCH();
    mis_prop_kernel<false><<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, strong.p, state.p, need, t1.p, t2.p); PNP_COUNT_LAUNCH();
    mis_decide_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, strong.p, t2.p, state.p, need_next, cnt.p); PNP_COUNT_LAUNCH();
    std::swap(need, need_next);
    PNP_CUDA(cudaMemcpyAsync(&und, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (++rounds > 1000) throw Error(ERROR_AMG_COARSEING, "MIS-2 did not terminate");
  }
  }
  if (g_phase) printf("    [setup l%d] mis2 rounds %d\n", g_phase_lvl, rounds);
  std::unique_ptr<Phase> ph(new Phase(s, g_phase, "aggregate", g_phase_lvl));
  // aggregate ids
  struct { int* p; } flag, rootid, agg2; flag.p = ws.get<int>(n + 1); rootid.p = ws.get<int>(n + 1); agg2.p = ws.get<int>(n);
  root_flag_kernel<<<cdiv(n + 1, 256), 256, 0, s>>>(n, state.p, flag.p); PNP_COUNT_LAUNCH();
  exclusive_scan_int(flag.p, rootid.p, n + 1, s);
  int nagg = 0;
  PNP_CUDA(cudaMemcpyAsync(&nagg, rootid.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  const bool dist = h.comm.active();
  const int n_own = owned_rows(L);
  {
    int fail = (nagg == 0 || nagg >= n_own) ? 1 : 0;
    if (dist) {      // every rank must take the same branch (the cycle is a collective)
      PNP_CUDA(cudaMemcpyAsync(cnt.p, &fail, sizeof(int), cudaMemcpyHostToDevice, s));
      comm_allreduce_int(h, cnt.p, 1, true);
      PNP_CUDA(cudaMemcpyAsync(&fail, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      PNP_CUDA(cudaStreamSynchronize(s));
    }
    if (fail) return false;
  }
  L.agg.alloc(n, s);
  agg_root_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, state.p, rootid.p, iso.p, L.agg.p); PNP_COUNT_LAUNCH();
  int left = 1; int* cur = L.agg.p; int* nxt = agg2.p;
  for (int pass = 0; pass < 8 && left > 0; ++pass) {
    cnt.zero(s);
    agg_join_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, A.ia.p, A.rs.p, A.ja.p, strong.p, pp.p, cur, nxt, cnt.p); PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaMemcpyAsync(&left, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    std::swap(cur, nxt);
  }
  if (cur != L.agg.p) PNP_CUDA(cudaMemcpyAsync(L.agg.p, cur, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  agg_finalize_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L.agg.p); PNP_COUNT_LAUNCH();
  L.nagg = nagg;                 // aggregates owned by this rank = owned coarse rows
  // multi-GPU: aggregates never cross ranks. A ghost row learns its owner's aggregate; the distinct
  // remote aggregates a neighbour references become the ghost rows of the coarse level (ascending
  // owner-local id on both sides, so sender and receiver agree on the order without talking)
  int ncoarse = nagg;
  C.n_own = -1; C.halo = Halo();
  if (dist) {
    Halo& hl = L.halo; Halo& hc = C.halo;
    halo_exchange_int(h, hl, L.agg.p);
    const int nghost = n - n_own, nsend = hl.active() ? hl.send_ptr.back() : 0;
    std::vector<int> gval(nghost > 0 ? nghost : 1), sval(nsend > 0 ? nsend : 1);
    if (nghost > 0) PNP_CUDA(cudaMemcpyAsync(gval.data(), L.agg.p + n_own, (size_t)nghost * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (nsend > 0) {
      hl.isendbuf.alloc(nsend, s);
      gather_int_kernel<<<cdiv(nsend, 256), 256, 0, s>>>(nsend, hl.send_idx.p, L.agg.p, hl.isendbuf.p); PNP_COUNT_LAUNCH();
      PNP_CUDA(cudaMemcpyAsync(sval.data(), hl.isendbuf.p, (size_t)nsend * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    PNP_CUDA(cudaStreamSynchronize(s));
    hc.nbr = hl.nbr; hc.send_ptr.assign(1, 0); hc.recv_ptr.assign(1, nagg);
    std::vector<int> csend;
    for (size_t k = 0; k < hl.nbr.size(); ++k) {
      std::vector<int> u(gval.begin() + (hl.recv_ptr[k] - n_own), gval.begin() + (hl.recv_ptr[k + 1] - n_own));
      std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
      while (!u.empty() && u.front() < 0) u.erase(u.begin());
      const int base = hc.recv_ptr.back();
      for (int g = hl.recv_ptr[k] - n_own; g < hl.recv_ptr[k + 1] - n_own; ++g)
        if (gval[g] >= 0) gval[g] = base + (int)(std::lower_bound(u.begin(), u.end(), gval[g]) - u.begin());
      hc.recv_ptr.push_back(base + (int)u.size());
      std::vector<int> v(sval.begin() + hl.send_ptr[k], sval.begin() + hl.send_ptr[k + 1]);
      std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
      while (!v.empty() && v.front() < 0) v.erase(v.begin());
      csend.insert(csend.end(), v.begin(), v.end());
      hc.send_ptr.push_back((int)csend.size());
    }
    ncoarse = hc.recv_ptr.back();
    if (nghost > 0) PNP_CUDA(cudaMemcpyAsync(L.agg.p + n_own, gval.data(), (size_t)nghost * sizeof(int), cudaMemcpyHostToDevice, s));
    hc.send_idx.alloc(csend.empty() ? 1 : csend.size(), s);
    if (!csend.empty()) hc.send_idx.upload(csend.data(), csend.size());
    hc.sendbuf.alloc(4 * (csend.empty() ? 1 : csend.size()), s);
    PNP_CUDA(cudaStreamSynchronize(s));     // gval / csend are host temporaries
    C.n_own = nagg;
  }
  const int nghost_c = ncoarse - nagg;
  // members of each aggregate (ascending rows)
  {
    struct { int* p; } c; c.p = ws.get<int>(ncoarse + 1);
    PNP_CUDA(cudaMemsetAsync(c.p, 0, (size_t)(ncoarse + 1) * sizeof(int), s));
    agg_count_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L.agg.p, c.p); PNP_COUNT_LAUNCH();
    L.agg_ptr.alloc(ncoarse + 1, s);
    exclusive_scan_int(c.p, L.agg_ptr.p, ncoarse + 1, s);
    struct { unsigned* p; } kin, kout; kin.p = ws.get<unsigned>(n); kout.p = ws.get<unsigned>(n);
    struct { int* p; } vin; vin.p = ws.get<int>(n); L.agg_rows.alloc(n, s);
    agg_key_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L.agg.p, kin.p); PNP_COUNT_LAUNCH();
    iota_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, vin.p); PNP_COUNT_LAUNCH();
    sort_pairs(ws, kin.p, kout.p, vin.p, L.agg_rows.p, n, 32, s);
  }
  // coarse structure. Row-wise path first (rapw_* kernels above): per coarse row, hash set of the
  // renamed columns -> counts -> scan -> sorted column lists.
  ph.reset(); ph.reset(new Phase(s, g_phase, "rap symbolic", g_phase_lvl));
  L.rap_rowwise = false;
  if (rowwise_on) {
    BsrT& Ac = C.A;
    // nothing taken from the workspace so far is still needed (the aggregates live in the level's own
    // buffers): start over, and make sure the parked column lists fit whatever the coarsening ratio was
    ws.reset();
    ws.reserve((size_t)ncoarse * (4 * RAPW_TMPW + 128) + (size_t)nnzb * 4 + (64u << 20), s);
    struct { int* p; } rowcnt, mx, tmpcols; rowcnt.p = ws.get<int>(ncoarse + 1); mx.p = ws.get<int>(1);
    tmpcols.p = ws.get<int>((size_t)ncoarse * RAPW_TMPW);
    PNP_CUDA(cudaMemsetAsync(rowcnt.p + ncoarse, 0, sizeof(int), s));
    PNP_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), s));
    rapw_symbolic_kernel<false><<<cdiv(ncoarse, 16), 256, 0, s>>>(nagg, nghost_c, L.agg_ptr.p, L.agg_rows.p, A.rd.p, A.ja.p, L.agg.p,
                                                                 rowcnt.p, mx.p, tmpcols.p, nullptr, nullptr); PNP_COUNT_LAUNCH();
    Ac.ia.alloc(ncoarse + 1, s);
    exclusive_scan_int(rowcnt.p, Ac.ia.p, ncoarse + 1, s);
    int hv[2] = {0, 0};
    PNP_CUDA(cudaMemcpyAsync(&hv[0], Ac.ia.p + ncoarse, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaMemcpyAsync(&hv[1], mx.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    if (hv[1] <= RAPW_MAXCOLS) {
      const int cnnzb = hv[0];
      Ac.n = ncoarse; Ac.nnzb = cnnzb;
      Ac.val.alloc(9 * (size_t)cnnzb, s);
      struct { int* p; } cja_nat; cja_nat.p = ws.get<int>(cnnzb);
      rapw_symbolic_kernel<true><<<cdiv(ncoarse, 16), 256, 0, s>>>(nagg, nghost_c, L.agg_ptr.p, L.agg_rows.p, A.rd.p, A.ja.p, L.agg.p,
                                                                  nullptr, nullptr, tmpcols.p, Ac.ia.p, cja_nat.p); PNP_COUNT_LAUNCH();
      ph.reset(); ph.reset(new Phase(s, g_phase, "coarse colour+layout", g_phase_lvl));
      color_and_layout(C, cja_nat.p, &ws, s);
      L.rap_rowwise = true; L.rap_maxcols = hv[1];
      return true;
    }
    // a coarse row is longer than the row-wise kernels hold: sort-based path, which needs the big workspace.
    // Nothing taken from the workspace so far is still needed (aggregates live in the level's own buffers).
    ws.reserve((size_t)nnzb * 60 + (size_t)n * 112 + (64u << 20), s);
    ws.reset();
  }
  // coarse structure: sort blocks by (I,J). (Tried: gathering the blocks aggregate by aggregate and
  // sorting each aggregate's ~165 entries on J with cub::DeviceSegmentedSort — the segmented sort
  // alone took 52 ms at 192^3 (107 M entries in 640 k segments) where this whole radix-sort path
  // takes 22 ms at 256^3 (253 M entries).)
  ph.reset(); ph.reset(new Phase(s, g_phase, "rap_sort+structure", g_phase_lvl));
  {
    const int nnzx = nnzb + nghost_c;      // fine blocks + identity diagonals of the coarse ghost rows
    const u64 nagg64 = (u64)ncoarse;
    struct { u64* p; } kin, kout; kin.p = ws.get<u64>(nnzx); kout.p = ws.get<u64>(nnzx);
    struct { int* p; } vin; vin.p = ws.get<int>(nnzx); L.rap_perm.alloc(nnzx, s);
    rap_key_kernel<<<cdiv(nnzx, 256), 256, 0, s>>>(nnzb, nghost_c, nagg, A.ja.p, A.brow.p, L.agg.p, nagg64, kin.p, vin.p); PNP_COUNT_LAUNCH();
    sort_pairs(ws, kin.p, kout.p, vin.p, L.rap_perm.p, nnzx, 2 * bits_for(nagg64) + 1, s);
    struct { int* p; } head, slot; head.p = ws.get<int>(nnzx + 1); slot.p = ws.get<int>(nnzx + 1);
    rap_head_kernel<<<cdiv(nnzx + 1, 256), 256, 0, s>>>(nnzx, kout.p, head.p); PNP_COUNT_LAUNCH();
    exclusive_scan_int(head.p, slot.p, nnzx + 1, s);
    int cnnzb = 0;
    PNP_CUDA(cudaMemcpyAsync(&cnnzb, slot.p + nnzx, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    BsrT& Ac = C.A;
    Ac.n = ncoarse; Ac.nnzb = cnnzb;
    Ac.ia.alloc(ncoarse + 1, s); Ac.val.alloc(9 * (size_t)cnnzb, s);
    struct { int* p; } cja_nat, crow, seg; cja_nat.p = ws.get<int>(cnnzb); crow.p = ws.get<int>(cnnzb); seg.p = ws.get<int>(cnnzb + 1);
    L.rap_beg.alloc(cnnzb, s); L.rap_end.alloc(cnnzb, s);
    struct { int* p; } rowcnt; rowcnt.p = ws.get<int>(ncoarse + 1);
    PNP_CUDA(cudaMemsetAsync(rowcnt.p, 0, (size_t)(ncoarse + 1) * sizeof(int), s));
    int* nvalid = seg.p + cnnzb;
    rap_structure_kernel<<<cdiv(nnzx, 256), 256, 0, s>>>(nnzx, kout.p, head.p, slot.p, nagg64, cja_nat.p, crow.p, seg.p,
                                                        rowcnt.p, nvalid); PNP_COUNT_LAUNCH();
    exclusive_scan_int(rowcnt.p, Ac.ia.p, ncoarse + 1, s);
    // colour the coarse graph and store it colour by colour, then re-key the numeric map
    ph.reset(); ph.reset(new Phase(s, g_phase, "coarse colour+layout", g_phase_lvl));
    color_and_layout(C, cja_nat.p, &ws, s);
    rap_remap_kernel<<<cdiv(cnnzb, 256), 256, 0, s>>>(cnnzb, crow.p, cja_nat.p, Ac.ia.p, Ac.rs.p, Ac.ja.p, seg.p, L.rap_beg.p, L.rap_end.p);
    PNP_COUNT_LAUNCH();
  }
  return true;
}

void rap_numeric(Handle& h, Level& L, Level& C)
{
  cudaStream_t s = h.stream;
  BsrT& A = L.A; BsrT& Ac = C.A;
  if (L.rap_rowwise) {
    const int nown = L.nagg, ngh = Ac.n - L.nagg;
    if (L.rap_maxcols <= 32)
      rapw_numeric_kernel<32><<<cdiv(Ac.n, 12), 192, 0, s>>>(nown, ngh, L.agg_ptr.p, L.agg_rows.p, A.rd.p, A.ja.p, L.agg.p, A.val.p, Ac.rd.p, Ac.ja.p, Ac.val.p);
    else if (L.rap_maxcols <= 64)
      rapw_numeric_kernel<64><<<cdiv(Ac.n, 6), 96, 0, s>>>(nown, ngh, L.agg_ptr.p, L.agg_rows.p, A.rd.p, A.ja.p, L.agg.p, A.val.p, Ac.rd.p, Ac.ja.p, Ac.val.p);
    else
      rapw_numeric_kernel<RAPW_MAXCOLS><<<cdiv(Ac.n, 2), 32, 0, s>>>(nown, ngh, L.agg_ptr.p, L.agg_rows.p, A.rd.p, A.ja.p, L.agg.p, A.val.p, Ac.rd.p, Ac.ja.p, Ac.val.p);
    PNP_COUNT_LAUNCH();
    return;
  }
  rap_numeric_kernel<<<cdiv(9 * (size_t)Ac.nnzb, 256), 256, 0, s>>>(Ac.nnzb, Ac.rd.p, Ac.brow.p, L.rap_beg.p, L.rap_end.p,
                                                                   L.rap_perm.p, A.rd.p, A.brow.p, A.val.p, Ac.val.p);
  PNP_COUNT_LAUNCH();
}

void alloc_level_vectors(Level& L, cudaStream_t s)
{
  const size_t m = VS * (size_t)L.A.n;
  auto fresh = [&](DevBuf<double>& v) { const double* old = v.p; v.alloc(m, s); if (v.p != old) v.zero(); };   // pad entries start at zero
  fresh(L.x); fresh(L.b); fresh(L.w);
  L.dinv.alloc(9 * (size_t)L.A.n, s);
}

} // namespace

SolverConfig make_config(const itsolver_param* it, const AMG_param* amg)
{
  SolverConfig c;
  if (it) {
    c.tol = it->tol; c.maxit = it->maxit; c.restart = it->restart > 0 ? it->restart : 30;
    c.print_level = it->print_level;
  }
  if (amg) {
    c.max_levels = amg->max_levels > 0 ? amg->max_levels : 20;
    c.coarse_dof = amg->coarse_dof > 0 ? amg->coarse_dof : 100;
    c.strong_coupled = amg->strong_coupled; c.max_aggregation = amg->max_aggregation;
    c.tentative_smooth = amg->tentative_smooth;
    c.presmooth = amg->presmooth_iter; c.postsmooth = amg->postsmooth_iter;
    if (amg->print_level > c.print_level) c.print_level = amg->print_level;
    c.cycle_type = amg->cycle_type > 0 ? amg->cycle_type : V_CYCLE;
    if (const char* e = getenv("PNPB200_KCYCLE_LEVELS")) c.kcycle_levels = atoi(e);
  } else c.use_amg = 0;
  if (c.restart > 31) c.restart = 31;      // fused orthogonalisation kernels carry <= 32 coefficients
  if (c.restart > c.maxit && c.maxit > 0) c.restart = c.maxit;
  return c;
}

namespace {

// (sum of owned rows over ranks, min of owned rows over ranks) — a collective when distributed
void global_rows(Handle& h, int own, int* sum, int* mn)
{
  *sum = own; *mn = own;
  if (!h.comm.active()) return;
  cudaStream_t s = h.stream;
  DevBuf<int> t; t.alloc(2, s);
  int v[2] = {own, -own};
  PNP_CUDA(cudaMemcpyAsync(t.p, v, 2 * sizeof(int), cudaMemcpyHostToDevice, s));
  comm_allreduce_int(h, t.p, 1, false);
  comm_allreduce_int(h, t.p + 1, 1, true);
  PNP_CUDA(cudaMemcpyAsync(v, t.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  *sum = v[0]; *mn = -v[1];
}

__global__ void gid_kernel(int n_own, int offset, int* __restrict__ gid)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_own) gid[i] = offset + i;
}

// dense rows of the owned block rows with GLOBAL column ids (row-major, ld = ndg)
__global__ void dense_fill_rows_kernel(int nnzb, int n_own, int ndg, const int* __restrict__ ia, const int* __restrict__ rs,
                                       const int* __restrict__ dslot, const int* __restrict__ ja, const int* __restrict__ brow, const int* __restrict__ gid,
                                       const double* __restrict__ val, double* __restrict__ D)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * (size_t)nnzb) return;
  const int b = (int)(e / 9), t = (int)(e % 9);
  const int i = brow[b];
  if (i >= n_own) return;
  const int s = rs[i], len = ia[i + 1] - ia[i];
  int st;
  const size_t o = bsrt_off(s, len, dslot[i] - s, b - s, st);
  D[(size_t)(3 * i + t / 3) * ndg + 3 * gid[ja[b]] + t % 3] = val[o + (size_t)t * st];
}

// bglob[3*off[r] + q] = ball[r*3*maxown + q], q < 3*cnt[r]
__global__ void compact_b_kernel(int nranks, int maxown3, const int* __restrict__ off3, const int* __restrict__ cnt3,
                                 const double* __restrict__ ball, double* __restrict__ bglob)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nranks * maxown3) return;
  const int r = e / maxown3, q = e % maxown3;
  if (q < cnt3[r]) bglob[off3[r] + q] = ball[e];
}

// x[row] = Inv[row0 + row, :] . b   for row < nrows, one warp per row
__global__ void dense_matvec_rows_kernel(int nrows, int row0, int nd, const double* __restrict__ Inv,
                                         const double* __restrict__ b, double* __restrict__ x)
{
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nrows) return;
  double a = 0.0;
  for (int c = lane; c < nd; c += 32) a += Inv[(size_t)(row0 + row) * nd + c] * b[c];
  a = warp_sum(a);
  if (lane == 0) x[row] = a;
}

// Largest coarsest level (scalar unknowns, global) that gets the explicit dense inverse. One CTA runs
// the Gauss-Jordan elimination (O(nd^3) on one SM: 0.3 ms at nd = 300, the bsr.dat case; ~40 ms at
// 1536), so anything bigger — a failed coarsening step, the multi-GPU early stop — is solved
// iteratively instead (solve_coarsest_sweeps).
constexpr int DENSE_COARSEST_MAX = 1536;

// coarsest level: every rank inverts the (small) GLOBAL coarsest matrix. Returns false if a pivot
// was zero or not finite (the inverse is unusable).
// Gather layout of the LAST distributed level Z of hierarchy H (its vectors are packed, row-id order): who owns how
// many rows, where my rows start in the global numbering, the buffers of the rhs gather; gid = global id
// of every local row (ghosts from their owners). cnt / off: per-rank row counts and offsets (host).
void gather_layout(Handle& h, Hierarchy& H, Level& Z, DevBuf<int>& gid, std::vector<int>& cnt, std::vector<int>& off)
{
  cudaStream_t s = h.stream;
  const int nr = h.comm.active() ? h.comm.nranks : 1, me = h.comm.active() ? h.comm.rank : 0;
  const int own = owned_rows(Z);
  DevBuf<int> t; t.alloc(nr + 1, s);
  PNP_CUDA(cudaMemcpyAsync(t.p + nr, &own, sizeof(int), cudaMemcpyHostToDevice, s));
  comm_allgather(h, t.p + nr, t.p, sizeof(int));
  cnt = t.to_host(nr);
  off.assign(nr + 1, 0);
  int maxown = 1;
  for (int r = 0; r < nr; ++r) { off[r + 1] = off[r] + cnt[r]; if (cnt[r] > maxown) maxown = cnt[r]; }
  const int ng = off[nr], nd = 3 * ng;
  H.nd = nd; H.c_own = own; H.c_row0 = 3 * off[me]; H.c_maxown3 = 3 * maxown; H.c_nranks = nr;
  std::vector<int> off3(nr), cnt3(nr);
  for (int r = 0; r < nr; ++r) { off3[r] = 3 * off[r]; cnt3[r] = 3 * cnt[r]; }
  H.c_off3.alloc(nr, s); H.c_cnt3.alloc(nr, s);
  H.c_off3.upload(off3.data(), nr); H.c_cnt3.upload(cnt3.data(), nr);
  gid.alloc(Z.A.n > 0 ? Z.A.n : 1, s);
  if (own > 0) { gid_kernel<<<cdiv(own, 256), 256, 0, s>>>(own, off[me], gid.p); PNP_COUNT_LAUNCH(); }
  halo_exchange_int(h, Z.halo, gid.p);
  H.c_bpad.alloc((size_t)3 * maxown, s); H.c_ball.alloc((size_t)3 * maxown * nr, s); H.c_bglob.alloc(nd > 0 ? nd : 1, s);
  PNP_CUDA(cudaStreamSynchronize(s));     // off3 / cnt3 are host temporaries
}

bool setup_coarsest(Handle& h)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = *h.cur;
  Level& Z = *H.lv.back();
  DevBuf<int> gid; std::vector<int> cnt, off;
  gather_layout(h, H, Z, gid, cnt, off);
  const int nr = H.c_nranks, own = H.c_own, nd = H.nd, maxown = H.c_maxown3 / 3;
  if (nd > DENSE_COARSEST_MAX) throw Error(ERROR_AMG_COARSEING, "coarsest level too large for the dense solver");
  // my dense rows, padded to the largest rank, then gathered and compacted
  const size_t rowblk = (size_t)3 * maxown * nd;
  DevBuf<double> Drow, Dall, D, colk; DevBuf<int> sing;
  Drow.alloc(rowblk, s); Drow.zero();
  Dall.alloc(rowblk * nr, s);
  if (Z.A.nnzb > 0) {
    dense_fill_rows_kernel<<<cdiv(9 * (size_t)Z.A.nnzb, 256), 256, 0, s>>>(Z.A.nnzb, own, nd, Z.A.ia.p, Z.A.rs.p, Z.A.dslot.p, Z.A.ja.p, Z.A.brow.p,
                                                                          gid.p, Z.A.val.p, Drow.p);
    PNP_COUNT_LAUNCH();
  }
  comm_allgather(h, Drow.p, Dall.p, rowblk * sizeof(double));
  D.alloc((size_t)nd * nd, s); colk.alloc(nd, s); sing.alloc(1, s); sing.zero();
  for (int r = 0; r < nr; ++r)
    if (cnt[r] > 0)
      PNP_CUDA(cudaMemcpyAsync(D.p + (size_t)3 * off[r] * nd, Dall.p + rowblk * r, (size_t)3 * cnt[r] * nd * sizeof(double),
                               cudaMemcpyDeviceToDevice, s));
  H.dense_inv.alloc((size_t)nd * nd, s);
  identity_kernel<<<cdiv((size_t)nd * nd, 256), 256, 0, s>>>(nd, H.dense_inv.p); PNP_COUNT_LAUNCH();
  dense_invert_kernel<<<1, 1024, 0, s>>>(nd, D.p, H.dense_inv.p, colk.p, sing.p); PNP_COUNT_LAUNCH();
  int hsing = 0;
  PNP_CUDA(cudaMemcpyAsync(&hsing, sing.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));     // host staging vectors above go out of scope
  PNP_LAUNCH_CHECK();
  return hsing == 0;
}

// Coarsest level too large (or too singular) for the dense inverse: COARSEST_SWEEPS symmetric
// multicolour Gauss-Seidel sweeps from a zero initial guess, in sweep space like every other level
// (multi-GPU: hybrid GS with a halo exchange before every sweep).
constexpr int COARSEST_SWEEPS = 8;
void solve_coarsest_sweeps(Handle& h, Level& Z, const double* b, double* x, int bs = VS)
{
  cudaStream_t s = h.stream;
  const bool dist = h.comm.active();
  const size_t m = VS * (size_t)Z.A.n, mo = VS * (size_t)owned_rows(Z);
  if (m > mo) PNP_CUDA(cudaMemsetAsync(x + mo, 0, (m - mo) * sizeof(double), s));
  mcgs_sweep_from_zero(Z, b, x, s, bs);
  for (int k = 0; k < COARSEST_SWEEPS; ++k) {
    if (dist) halo_exchange(h, Z.halo, x, true);
    mcgs_sweep(Z, b, x, true, s, bs);
    if (k + 1 == COARSEST_SWEEPS) break;
    if (dist) halo_exchange(h, Z.halo, x, true);
    mcgs_sweep(Z, b, x, false, s, bs);
  }
}

// the rhs of the last distributed level (packed, row-id order) on every rank: H.c_bglob (or b itself on one GPU)
const double* gather_rhs(Handle& h, Hierarchy& H, const double* b)
{
  cudaStream_t s = h.stream;
  if (H.c_nranks <= 1) return b;
  PNP_CUDA(cudaMemsetAsync(H.c_bpad.p, 0, (size_t)H.c_maxown3 * sizeof(double), s));
  if (H.c_own > 0) PNP_CUDA(cudaMemcpyAsync(H.c_bpad.p, b, (size_t)3 * H.c_own * sizeof(double), cudaMemcpyDeviceToDevice, s));
  comm_allgather(h, H.c_bpad.p, H.c_ball.p, (size_t)H.c_maxown3 * sizeof(double));
  compact_b_kernel<<<cdiv((size_t)H.c_nranks * H.c_maxown3, 256), 256, 0, s>>>(H.c_nranks, H.c_maxown3, H.c_off3.p, H.c_cnt3.p,
                                                                              H.c_ball.p, H.c_bglob.p);
  PNP_COUNT_LAUNCH();
  return H.c_bglob.p;
}

void solve_coarsest(Handle& h, const double* b, double* x)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = *h.cur;
  const double* bg = gather_rhs(h, H, b);
  if (H.c_own > 0) {
    dense_matvec_rows_kernel<<<cdiv((size_t)3 * H.c_own * 32, 256), 256, 0, s>>>(3 * H.c_own, H.c_row0, H.nd, H.dense_inv.p, bg, x);
    PNP_COUNT_LAUNCH();
  }
}

// ---- replicated coarse hierarchy (multi-GPU) ----------------------------------------
// Below REPLICATE_ROWS global block rows a level is all latency: a distributed visit costs 3 halo
// exchanges + 2 all-reduces of ~20 us each, and the K-cycle visits level l 2^(l-1) times per Krylov
// iteration. So the first level that small is gathered to EVERY rank (one all-gather of its matrix per
// setup, one all-gather of its rhs per visit) and everything below it — coarsening, colouring, cycle,
// coarsest solve — runs redundantly with the single-GPU code, aggregates no longer confined to a rank.
constexpr int REPLICATE_ROWS = 40000;

__global__ void rep_rowlen_kernel(int own, const int* __restrict__ ia, int* __restrict__ len)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < own) len[i] = ia[i + 1] - ia[i];
}
__global__ void rep_gid_cols_kernel(int nnz, const int* __restrict__ ja_nat, const int* __restrict__ gid, int* __restrict__ out)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) out[p] = gid[ja_nat[p]];
}

// ascending columns inside every row of a FASP-layout BSR(3) matrix (the local -> global renaming of the ghost
// columns breaks the order the layout code relies on); one thread per row, insertion sort, rows are short
__global__ void rep_sort_rows_kernel(int n, const int* __restrict__ ia, int* __restrict__ ja, double* __restrict__ val)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = ia[i], e = ia[i + 1];
  for (int a = s + 1; a < e; ++a) {
    const int key = ja[a];
    double v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = val[9 * (size_t)a + t];
    int b = a - 1;
    while (b >= s && ja[b] > key) {
      ja[b + 1] = ja[b];
#pragma unroll
      for (int t = 0; t < 9; ++t) val[9 * (size_t)(b + 1) + t] = val[9 * (size_t)b + t];
      --b;
    }
    ja[b + 1] = key;
#pragma unroll
    for (int t = 0; t < 9; ++t) val[9 * (size_t)(b + 1) + t] = v[t];
  }
}

void amg_setup_chain(Handle& h, Hierarchy& H, const SolverConfig& cfg);

// gather the last distributed level Z of h.hier into h.hier_rep.lv[0] on every rank and build the hierarchy below it
void setup_replicated(Handle& h, const SolverConfig& cfg, int level_index)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = h.hier; Hierarchy& R = h.hier_rep;
  Level& Z = *H.lv.back();
  DevBuf<int> gid; std::vector<int> cnt, off;
  gather_layout(h, H, Z, gid, cnt, off);
  const int nr = H.c_nranks, own = H.c_own, maxown = H.c_maxown3 / 3, ng = off[nr];
  // natural (FASP) form of my owned rows with GLOBAL column ids
  const BsrT& A = Z.A;
  DevBuf<int> ja_nat, iah; ja_nat.alloc(A.nnzb, s);
  bsrt_export_ja(A, ja_nat.p, s);
  std::vector<int> ia = A.ia.to_host(A.n + 1);
  const int mynnz = ia[own];
  DevBuf<int> t; t.alloc(nr + 1, s);
  PNP_CUDA(cudaMemcpyAsync(t.p + nr, &mynnz, sizeof(int), cudaMemcpyHostToDevice, s));
  comm_allgather(h, t.p + nr, t.p, sizeof(int));
  std::vector<int> nnz = t.to_host(nr);
  int maxnnz = 1; std::vector<int> nzoff(nr + 1, 0);
  for (int r = 0; r < nr; ++r) { nzoff[r + 1] = nzoff[r] + nnz[r]; if (nnz[r] > maxnnz) maxnnz = nnz[r]; }
  DevBuf<double> vf; vf.alloc(9 * (size_t)A.nnzb, s);
  bsrt_to_fasp(A, vf.p, s);
  DevBuf<int> lenp, lena, jap, jaa; DevBuf<double> valp, vala;
  lenp.alloc(maxown, s); lenp.zero(); lena.alloc((size_t)maxown * nr, s);
  jap.alloc(maxnnz, s); jap.zero(); jaa.alloc((size_t)maxnnz * nr, s);
  valp.alloc(9 * (size_t)maxnnz, s); vala.alloc(9 * (size_t)maxnnz * nr, s);
  if (own > 0) { rep_rowlen_kernel<<<cdiv(own, 256), 256, 0, s>>>(own, A.ia.p, lenp.p); PNP_COUNT_LAUNCH(); }
  if (mynnz > 0) {
    rep_gid_cols_kernel<<<cdiv(mynnz, 256), 256, 0, s>>>(mynnz, ja_nat.p, gid.p, jap.p); PNP_COUNT_LAUNCH();
    PNP_CUDA(cudaMemcpyAsync(valp.p, vf.p, 9 * (size_t)mynnz * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  comm_allgather(h, lenp.p, lena.p, (size_t)maxown * sizeof(int));
  comm_allgather(h, jap.p, jaa.p, (size_t)maxnnz * sizeof(int));
  comm_allgather(h, valp.p, vala.p, 9 * (size_t)maxnnz * sizeof(double));
  // the global matrix: row offsets on the host (tiny), columns and values compacted rank by rank
  std::vector<int> la = lena.to_host((size_t)maxown * nr);
  std::vector<int> gia(ng + 1, 0);
  for (int r = 0; r < nr; ++r) for (int i = 0; i < cnt[r]; ++i) gia[off[r] + i + 1] = la[(size_t)r * maxown + i];
  for (int i = 0; i < ng; ++i) gia[i + 1] += gia[i];
  if (gia[ng] != nzoff[nr]) throw Error(ERROR_DATA_STRUCTURE, "replicated level: row lengths and block counts disagree");
  if (R.lv.empty()) R.lv.emplace_back(new Level);
  Level& R0 = *R.lv[0];
  R0.n_own = -1; R0.halo = Halo(); R0.preset_ncolors = 0;
  BsrT& G = R0.A;
  G.n = ng; G.nnzb = gia[ng];
  G.ia.alloc(ng + 1, s); G.ia.upload(gia.data(), ng + 1);
  G.val.alloc(9 * (size_t)G.nnzb, s);
  DevBuf<int> gja; gja.alloc(G.nnzb, s);
  DevBuf<double> gval; gval.alloc(9 * (size_t)G.nnzb, s);
  for (int r = 0; r < nr; ++r)
    if (nnz[r] > 0) {
      PNP_CUDA(cudaMemcpyAsync(gja.p + nzoff[r], jaa.p + (size_t)r * maxnnz, (size_t)nnz[r] * sizeof(int), cudaMemcpyDeviceToDevice, s));
      PNP_CUDA(cudaMemcpyAsync(gval.p + 9 * (size_t)nzoff[r], vala.p + 9 * (size_t)r * maxnnz, 9 * (size_t)nnz[r] * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
  rep_sort_rows_kernel<<<cdiv(ng, 128), 128, 0, s>>>(ng, G.ia.p, gja.p, gval.p); PNP_COUNT_LAUNCH();
  PNP_CUDA(cudaStreamSynchronize(s));     // gia is a host temporary
  // from here on this rank works alone: same code as one GPU
  h.comm.suspended = true;
  Hierarchy* save = h.cur;
  try {
    h.cur = &R;
    color_and_layout(R0, gja.p, nullptr, s);
    bsrt_from_fasp(gval.p, G, s);
    R.level_offset = level_index;
    R.symbolic_valid = false;
    amg_setup_chain(h, R, cfg);
  } catch (...) { h.cur = save; h.comm.suspended = false; throw; }
  h.cur = save; h.comm.suspended = false;
  H.c_xglob.alloc(3 * (size_t)ng, s);
}

void solve_level(Handle& h, const SolverConfig& cfg, int l, const double* b, double* x);

// x = approximate solution of the last distributed level's system, computed redundantly on the replicated hierarchy
void solve_replicated(Handle& h, const SolverConfig& cfg, const double* b, double* x)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = h.hier; Hierarchy& R = h.hier_rep;
  Level& R0 = *R.lv[0];
  const double* bg = gather_rhs(h, H, b);
  const bool single = R.lv.size() == 1 && R.coarse_dense;      // the gathered level is itself the coarsest: dense solve, row-id vectors
  if (!single) vec_to_sweep(R0.A, bg, R0.b.p, s);
  h.comm.suspended = true;
  Hierarchy* save = h.cur;
  try {
    h.cur = &R;
    if (single) solve_level(h, cfg, 0, bg, H.c_xglob.p);
    else solve_level(h, cfg, 0, R0.b.p, R0.x.p);
  } catch (...) { h.cur = save; h.comm.suspended = false; throw; }
  h.cur = save; h.comm.suspended = false;
  if (!single) vec_from_sweep(R0.A, R0.x.p, H.c_xglob.p, s);
  if (H.c_own > 0) PNP_CUDA(cudaMemcpyAsync(x, H.c_xglob.p + H.c_row0, (size_t)3 * H.c_own * sizeof(double), cudaMemcpyDeviceToDevice, s));
}

} // namespace

namespace {
// build (or, with hierarchy reuse, refresh) the levels of H below its level 0. H is h.hier, or the replicated
// coarse hierarchy of a multi-GPU run (then the communicator is suspended and H.level_offset > 0).
void amg_setup_chain(Handle& h, Hierarchy& H, const SolverConfig& cfg)
{
  cudaStream_t s = h.stream;
  auto& lv = H.lv;
  Level& F = *lv[0];
  const bool reuse = h.reuse_hierarchy && H.symbolic_valid;
  if (F.ncolors == 0) throw Error(ERROR_DATA_STRUCTURE, "fine level not coloured");   // done at create / upload
  for (size_t k = 1; k < lv.size(); ++k) lv[k]->A.packed7 = false;     // coarse values are rebuilt below
  for (size_t k = 0; k < lv.size(); ++k) lv[k]->A.pack7_fresh = false;
  alloc_level_vectors(F, s);
  if (!reuse) H.symbolic_valid = false;
  H.drop_coarse_graph();           // buffers and launch sizes belong to the hierarchy being replaced
  const int min_cdof = cfg.coarse_dof > 50 ? cfg.coarse_dof : 50;
  static const int rep_rows = getenv("PNPB200_REPLICATE_ROWS") ? atoi(getenv("PNPB200_REPLICATE_ROWS")) : REPLICATE_ROWS;
  bool replicate = reuse ? H.coarse_rep : false;
  int l = 0;
  while (true) {
    Level& L = *lv[l];
    bsrt_diaginv(L.A, L.dinv.p, s);
    if (reuse) {
      if (l + 1 >= (int)lv.size()) break;
      rap_numeric(h, L, *lv[l + 1]);
      ++l; continue;
    }
    // the decision to coarsen further is taken on GLOBAL sizes so that every rank builds the
    // same number of levels (each level of the cycle is a collective)
    int nglob, nmin;
    global_rows(h, owned_rows(L), &nglob, &nmin);
    if (!(nglob > min_cdof && l + H.level_offset < cfg.max_levels - 1)) break;
    if (h.comm.active() && l >= 1 && nglob <= rep_rows) { replicate = true; break; }     // small enough: every rank takes the whole level
    if (h.comm.active() && nmin < 8) break;
    const double theta = cfg.tentative_smooth > 1e-20 ? cfg.strong_coupled * std::pow(0.5, l + H.level_offset) : cfg.strong_coupled;
    // Level objects (and their device buffers) are recycled from the previous hierarchy
    if (l + 1 >= (int)lv.size()) lv.emplace_back(new Level);
    Level& C = *lv[l + 1];
    g_phase = cfg.print_level >= 3; g_phase_lvl = l + H.level_offset;
    if (!coarsen_symbolic(h, L, C, theta)) break;
    C.halo.ipc_slot = l + 1 + H.level_offset; C.halo.ipc_state = 0;     // peer-memory exchange: slot of this level (comm.cu)
    { PHASE("rap_numeric"); rap_numeric(h, L, C); }
    alloc_level_vectors(C, s);
    ++l;
  }
  if (!reuse) lv.resize(l + 1);
  // last level: replicated coarse hierarchy (multi-GPU), explicit dense inverse when it is small (the bsr.dat
  // case, <= coarse_dof rows), Gauss-Seidel sweeps when coarsening stopped early or a pivot was unusable
  if (replicate) {
    H.coarse_rep = true; H.coarse_dense = false;
    setup_replicated(h, cfg, (int)lv.size() - 1 + H.level_offset);
  } else {
    H.coarse_rep = false;
    int nglob, nmin;
    global_rows(h, owned_rows(*lv.back()), &nglob, &nmin);
    bool dense = 3 * (long)nglob <= DENSE_COARSEST_MAX;
    if (reuse && dense != H.coarse_dense) throw Error(ERROR_SOLVER_MISC, "coarsest level changed its solver under hierarchy reuse");
    if (dense && !setup_coarsest(h)) {
      if (reuse) throw Error(ERROR_SOLVER_MISC, "coarsest level is singular");
      dense = false;
    }
    H.coarse_dense = dense;
  }
  if (!reuse) {
    // grid transfer maps in sweep space; a last level with the dense or the replicated solve keeps row-id order
    for (int k = 0; k + 1 < (int)lv.size(); ++k) {
      Level& Lk = *lv[k]; Level& Ck = *lv[k + 1];
      Lk.coarse_identity = (k + 2 == (int)lv.size()) && (H.coarse_dense || H.coarse_rep);
      const int nk = Lk.A.n;
      Lk.agg_pos.alloc(nk, s); Lk.agg_rows_pos.alloc(nk, s);
      transfer_maps_kernel<<<cdiv(nk, 256), 256, 0, s>>>(nk, Lk.A.vrowid.p, Lk.A.pos.p, Lk.agg.p, Lk.agg_rows.p,
                                                        Lk.coarse_identity ? nullptr : Ck.A.pos.p, Lk.agg_pos.p, Lk.agg_rows_pos.p);
      PNP_COUNT_LAUNCH();
    }
  }
  H.symbolic_valid = true;
  { PHASE("pack7"); bsrt_pack7(h, H); }     // the solve's matrix passes read 7 values per block where the structure allows
  if (cfg.print_level > 0) {
    PNP_CUDA(cudaStreamSynchronize(s));
    for (size_t k = 0; k < lv.size(); ++k)
      printf("  [pnpb200 amg] rank %d level %zu%s: rows %d (owned %d) nnzb %d colours %d\n", h.comm.rank, k + H.level_offset,
             H.level_offset ? " (replicated)" : "", lv[k]->A.n, owned_rows(*lv[k]), lv[k]->A.nnzb, lv[k]->ncolors);
  }
}
} // namespace

void amg_setup(Handle& h, const SolverConfig& cfg)
{
  h.cur = &h.hier;
  h.hier.level_offset = 0;
  amg_setup_chain(h, h.hier, cfg);
}

namespace {

inline int small_grid(size_t n) { const size_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 592 ? 592 : b)); }

void solve_level(Handle& h, const SolverConfig& cfg, int l, const double* b, double* x);

// Approximate solve on level l (>= 2) with rhs lv[l]->b and result lv[l]->x, replayed from a CUDA
// graph. Everything in there runs on fixed buffers with device-side scalars and no host
// synchronisation, so the launch sequence is the same at every visit of one hierarchy; the
// K-cycle makes two such visits per Krylov iteration, ~200 launches of 2-8 us kernels each.
// First visit: direct (lazy scratch allocations, function attributes, NCCL connections), second:
// capture + instantiate, later ones: one cudaGraphLaunch. PNPB200_COARSE_GRAPH=0 switches it off,
// =k moves the cut to level k. Measured: 1 GPU 256^3 783 -> 769 ms per step; 2 GPUs 128^3 per GPU
// 167 -> 159 ms (Krylov stage 130 -> 120 ms: the NCCL group launches of the coarse levels are
// CPU-bound when issued one by one).
int coarse_graph_level()
{
  static const int lvl = getenv("PNPB200_COARSE_GRAPH") ? atoi(getenv("PNPB200_COARSE_GRAPH")) : 2;
  return lvl;
}
// multi-GPU: the halo exchanges and allreduces of the coarse levels are NCCL calls on the same stream
// and are captured with the kernels (every rank captures the same sequence: the level structure is
// decided on global sizes). PNPB200_COARSE_GRAPH_DIST=0 keeps the distributed cycle uncaptured.
bool coarse_graph_dist()
{
  static const bool on = getenv("PNPB200_COARSE_GRAPH_DIST") ? atoi(getenv("PNPB200_COARSE_GRAPH_DIST")) != 0 : true;
  return on;
}
void coarse_solve_graphed(Handle& h, const SolverConfig& cfg, int l)
{
  Hierarchy& H = *h.cur;
  Level& C = *H.lv[l];
  cudaStream_t s = h.stream;
  const unsigned sig = hash32((unsigned)cfg.cycle_type * 7919u + (unsigned)cfg.kcycle_levels * 104729u + (unsigned)cfg.presmooth * 31u +
                              (unsigned)cfg.postsmooth * 131u + (unsigned)h.sweep_shortcuts * 3u + (unsigned)l * 1299709u);
  if (H.coarse_exec && H.coarse_sig != sig) H.drop_coarse_graph();
  if (H.coarse_exec) {
    PNP_CUDA(cudaGraphLaunch(H.coarse_exec, s));
    g_launches += H.coarse_nodes;
    return;
  }
  if (H.coarse_visits++ == 0) { solve_level(h, cfg, l, C.b.p, C.x.p); return; }
  const long before = g_launches;
  cudaGraph_t graph = nullptr;
  PNP_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  try { solve_level(h, cfg, l, C.b.p, C.x.p); }
  catch (...) { cudaStreamEndCapture(s, &graph); if (graph) cudaGraphDestroy(graph); throw; }
  PNP_CUDA(cudaStreamEndCapture(s, &graph));
  H.coarse_nodes = g_launches - before;
  g_launches = before;
  const cudaError_t e = cudaGraphInstantiate(&H.coarse_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { H.coarse_exec = nullptr; throw Error(ERROR_DEVICE, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
  H.coarse_sig = sig;
  PNP_CUDA(cudaGraphLaunch(H.coarse_exec, s));
  g_launches += H.coarse_nodes;
}

// one multigrid cycle on level l from a zero initial guess: x = B_l b. Multi-GPU: hybrid
// Gauss-Seidel (ghost values refreshed from their owners before every sweep / residual).
// With one pre-smoothing sweep the zero initial guess is exploited: the forward sweep reads only
// the lower parts of the rows and the residual after it only the upper parts (bsr_ops.cu).
// If Ax != nullptr (>= 1 post-smoothing sweep) the last backward sweep also records
// delta = x_new - x_old in L.w, and Ax = A x = b + L delta costs half a matrix pass; returns
// whether Ax was produced.
bool mg_cycle(Handle& h, const SolverConfig& cfg, int l, const double* b, double* x, double* Ax, int bs = VS, int axs = VS)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = *h.cur;
  Level& L = *H.lv[l];
  Level& C = *H.lv[l + 1];
  const size_t m = VS * (size_t)L.A.n;
  const int own = owned_rows(L);
  const bool dist = h.comm.active();
  if (cfg.presmooth == 1 && h.sweep_shortcuts) {
    if (m > VS * (size_t)own) PNP_CUDA(cudaMemsetAsync(x + VS * (size_t)own, 0, (m - VS * (size_t)own) * sizeof(double), s));
    // debugging aid: the forward sweep from zero must never read an entry it has not written yet; PNPB200_POISON=1
    // fills the owned part with NaN bit patterns first, so that such a read surfaces as a failed solve
    static const bool poison = getenv("PNPB200_POISON") && atoi(getenv("PNPB200_POISON"));
    if (poison && own > 0) PNP_CUDA(cudaMemsetAsync(x, 0xff, VS * (size_t)own * sizeof(double), s));
    mcgs_sweep_from_zero(L, b, x, s, bs);
    if (dist) halo_exchange(h, L.halo, x, true);
    bsrt_residual_after_sweep(L, b, x, L.w.p, s, bs);
  } else {
    vec_set(x, 0.0, m, s);
    for (int k = 0; k < cfg.presmooth; ++k) { if (dist && k > 0) halo_exchange(h, L.halo, x, true); mcgs_sweep(L, b, x, false, s, bs); }
    if (dist) halo_exchange(h, L.halo, x, true);
    bsrt_residual(L.A, b, x, L.w.p, s, bs, VS);
  }
  if (L.nagg > 0) {
    if (L.coarse_identity) restrict_kernel<3><<<cdiv(3 * (size_t)L.nagg, 256), 256, 0, s>>>(L.nagg, nullptr, L.agg_ptr.p, L.agg_rows_pos.p, L.w.p, C.b.p);
    else restrict_kernel<VS><<<cdiv(VS * (size_t)L.nagg, 256), 256, 0, s>>>(L.nagg, C.A.vrowid.p, L.agg_ptr.p, L.agg_rows_pos.p, L.w.p, C.b.p);
    PNP_COUNT_LAUNCH();
  }
  // (levels are counted globally: the replicated coarse hierarchy of a multi-GPU run continues the numbering)
  if (l + 1 + H.level_offset == coarse_graph_level() && l + 1 >= 1 && (l + 2 < (int)H.lv.size() || H.coarse_rep) && (!dist || coarse_graph_dist()) && !g_debug_sync) coarse_solve_graphed(h, cfg, l + 1);
  else solve_level(h, cfg, l + 1, C.b.p, C.x.p);
  if (own > 0) {
    if (L.coarse_identity) prolong_kernel<3><<<cdiv(VS * (size_t)own, 256), 256, 0, s>>>(own, L.agg_pos.p, C.x.p, x);
    else prolong_kernel<VS><<<cdiv(VS * (size_t)own, 256), 256, 0, s>>>(own, L.agg_pos.p, C.x.p, x);
    PNP_COUNT_LAUNCH();
  }
  const bool want_ax = Ax != nullptr && cfg.postsmooth >= 1 && h.sweep_shortcuts;
  for (int k = 0; k < cfg.postsmooth; ++k) {
    if (dist) halo_exchange(h, L.halo, x, true);
    if (want_ax && k == cfg.postsmooth - 1) mcgs_sweep_back_delta(L, b, x, L.w.p, s, bs);
    else mcgs_sweep(L, b, x, true, s, bs);
  }
  if (want_ax) {
    // multi-GPU (hybrid GS): the ghost columns kept their pre-sweep value, so their delta — the
    // owners' delta, one halo exchange — enters the product as well
    if (dist) halo_exchange(h, L.halo, L.w.p, true);
    bsrt_apply_after_sweep(L, b, L.w.p, Ax, s, bs, axs);
  }
  return want_ax;
}

// approximate solve of A_l x = b on a coarse level l >= 1
void solve_level(Handle& h, const SolverConfig& cfg, int l, const double* b, double* x)
{
  cudaStream_t s = h.stream;
  Hierarchy& H = *h.cur;
  const int nl = (int)H.lv.size();
  Level& L = *H.lv[l];
  if (l == nl - 1) {
    if (H.coarse_rep) solve_replicated(h, cfg, b, x);
    else if (H.coarse_dense) solve_coarsest(h, b, x);
    else solve_coarsest_sweeps(h, L, b, x);
    return;
  }
  const bool kcyc = cfg.cycle_type != V_CYCLE && l + H.level_offset >= 1 && l + H.level_offset <= cfg.kcycle_levels;
  if (!kcyc) { mg_cycle(h, cfg, l, b, x, nullptr); return; }
  // K-cycle: two GCR steps preconditioned by the cycle of this level (device scalars only;
  // multi-GPU: the dot products run over the owned rows and are summed with ncclAllReduce)
  const size_t m = VS * (size_t)L.A.n, mo = VS * (size_t)owned_rows(L);
  const bool dist = h.comm.active();
  { const double* old = L.kc.p; L.kc.alloc(5 * m + 8, s); if (L.kc.p != old) L.kc.zero(); }
  L.ksc.alloc(8 + 3 * 592, s);
  if (!L.kcnt.p) { L.kcnt.alloc(1, s); L.kcnt.zero(); }
  double* c = L.kc.p; double* v = c + m; double* rt = v + m; double* d = rt + m; double* w = d + m;
  double* sc = L.ksc.p; double* partial = sc + 8;
  const int g = small_grid(mo);
  if (!mg_cycle(h, cfg, l, b, c, v)) {
    if (dist) halo_exchange(h, L.halo, c, true);
    bsrt_spmv(L.A, c, v, s);
  }
  kdots_kernel<2><<<g, 256, 0, s>>>(mo, v, b, v, v, nullptr, nullptr, partial, L.kcnt.p, sc); PNP_COUNT_LAUNCH();
  if (dist) comm_allreduce(h, sc, 2, false);
  kc_step1_kernel<<<g, 256, 0, s>>>(mo, sc, b, v, rt); PNP_COUNT_LAUNCH();
  if (!mg_cycle(h, cfg, l, rt, d, w)) {
    if (dist) halo_exchange(h, L.halo, d, true);
    bsrt_spmv(L.A, d, w, s);
  }
  kdots_kernel<3><<<g, 256, 0, s>>>(mo, w, v, w, w, w, rt, partial, L.kcnt.p, sc + 2); PNP_COUNT_LAUNCH();
  if (dist) comm_allreduce(h, sc + 2, 3, false);
  kc_final_kernel<<<g, 256, 0, s>>>(mo, sc, c, d, x); PNP_COUNT_LAUNCH();
}

} // namespace

// z = M^{-1} r : one multigrid cycle from a zero initial guess (fasp_precond_dbsr_amg [EXT]):
// V(nu1,nu2), or with AMG_cycle_type != V the K-cycle on the first cfg.kcycle_levels coarse levels
bool amg_vcycle(Handle& h, const SolverConfig& cfg, const double* r, double* z, double* Az, int rs, int azs)
{
  cudaStream_t s = h.stream;
  h.cur = &h.hier;
  auto& lv = h.hier.lv;
  const int nl = (int)lv.size();
  if (nl == 1) {       // no coarse grid at all: the "cycle" is the coarsest-level solve on the fine matrix
    Level& Z = *lv[0];
    if (h.hier.coarse_dense) {      // the dense solve works in row-id order
      double* t0 = Z.b.p; double* t1 = Z.x.p;     // level work vectors: unused without a cycle
      vec_from_sweep(Z.A, r, t0, s, rs);
      solve_coarsest(h, t0, t1);
      vec_to_sweep(Z.A, t1, z, s);
    } else solve_coarsest_sweeps(h, Z, r, z, rs);
    (void)cfg; (void)s;
    return false;
  }
  return mg_cycle(h, cfg, 0, r, z, Az, rs, azs);
}

} // namespace pnpb200
