// pnp-b200 — streamed multicolour Gauss-Seidel colour launch for sm_100a: the matrix stream of a colour
// (values, column indices, row descriptors, block inverses — all CONTIGUOUS in sweep order, 72 of every
// ~100 bytes a row touches) is moved by the bulk-copy engine (cp.async.bulk = UBLKCP, completion on an
// mbarrier) into a ring of shared-memory stages, GST_STAGES chunks ahead of the arithmetic; only the
// gather of x (one 256-bit load per block) and the 24 bytes of b still go through the LSU path.
// Why: the one-row-per-half-warp kernel of bsr_ops.cu is LATENCY bound in the colour launches (ncu at
// 256^3: long-scoreboard stall 58 per issue, L1 data pipe 40 %, 4.9 of 6.5 TB/s): every row pays the
// chain descriptor -> columns -> gather, and the 9 value loads of a lane sit behind the first link.
// Here the first two links are shared-memory reads of data that landed microseconds ago.
//   grid   : persistent, 3 CTAs per SM, CTA = 16 consumer half-warps (one row of the chunk each)
//            + 1 producer warp (one lane issues the copies)
//   chunk  : 16 consecutive rows of the colour = one contiguous piece of val / jpos / prd / dinv
//   stage  : full barrier (producer arrives with expect_tx, the copies complete_tx), empty barrier
//            (one arrival per consumer half-warp once its shared-memory reads are in registers)
// Rows longer than GST_MAXLEN blocks do not fit a stage: levels that have one keep the plain kernel.
#include "internal.hpp"

namespace pnpb200 {

namespace {

constexpr int GST_ROWS = 16;
constexpr int GST_THREADS = GST_ROWS * 16 + 32;
// Two instantiations by the longest row of the level (ML blocks): 16 (the fine level of a Kuhn mesh: 15) and 32 (the
// first coarse levels: mean 17.5, longest 29). A half-warp takes a row in ML / 16 rounds of 16 blocks.
//   ML = 16: 57 KB per 3-stage ring, 3 CTAs per SM;   ML = 32: 81 KB per 2-stage ring, 2 CTAs per SM
template <int ML> struct GstCfg {
  static constexpr int STAGES = ML <= 16 ? 3 : 2;
  static constexpr int CTAS = ML <= 16 ? 3 : 2;
  static constexpr int ROWSLOT = ML * 9 + 2;      // doubles per row slot of a stage (a multiple of 16 bytes)
};

template <int ML>
struct alignas(128) GstStage {
  double val[GST_ROWS * GstCfg<ML>::ROWSLOT];     // + slack: a copy starts at the 16-byte boundary below its first byte
  double dinv[GST_ROWS * 9 + 2];
  int4 prd[GST_ROWS];
  int ja[GST_ROWS * ML + 4];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned addresses, size a multiple of 16), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The passes of bsr_ops.cu (same modes, same arithmetic, same reduction tree) over the rows [p0, p1) in
// MATRIX order. Which part of a row a mode reads decides how the producer feeds the ring:
//   full rows  (S_SPMV, S_RES, S_GS, S_GSD): the chunk's rows are one contiguous piece of val: ONE copy;
//   lower part (S_GS0, S_LDELTA) / upper part (S_RESU): the part of row r is contiguous inside the row
//              ([L | D | U] storage): producer lane r copies the part of row r into row slot r, so that
//              half a matrix pass really moves half the bytes. The lanes fetch the descriptors of the
//              NEXT chunk of their CTA one chunk ahead, so the producer never waits for them.
enum { S_SPMV = 0, S_RES = 1, S_GS = 2, S_GS0 = 3, S_RESU = 4, S_GSD = 5, S_LDELTA = 6 };

template <int MODE, int ML, int NV>
__global__ void __launch_bounds__(GST_THREADS, GstCfg<ML>::CTAS)
gs_stream_kernel(int p0, int p1, const int4* __restrict__ prd, const int* __restrict__ pia, const int* __restrict__ jpos,
                 const double* __restrict__ val, const double* __restrict__ dinv, const double* __restrict__ b,
                 const double* x, double* y, double* __restrict__ dl, int bs, int ys)
{
  constexpr bool LOWER = MODE == S_GS0 || MODE == S_LDELTA, UPPER = MODE == S_RESU, PART = LOWER || UPPER;
  constexpr bool SMOOTH = MODE == S_GS || MODE == S_GS0 || MODE == S_GSD;
  constexpr int GST_STAGES = GstCfg<ML>::STAGES, GST_ROWSLOT = GstCfg<ML>::ROWSLOT;
  using Stage = GstStage<ML>;
  extern __shared__ __align__(128) unsigned char gst_smem[];
  Stage* stages = reinterpret_cast<Stage*>(gst_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(gst_smem + GST_STAGES * sizeof(Stage));
  unsigned long long* empty = full + GST_STAGES;
  if (threadIdx.x == 0) {
    for (int q = 0; q < GST_STAGES; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], GST_ROWS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int nchunks = (p1 - p0 + GST_ROWS - 1) / GST_ROWS;

  if (threadIdx.x >= GST_ROWS * 16) {
    // ---- producer warp
    const int pl = threadIdx.x - GST_ROWS * 16;                 // lane
    if (!PART && pl != 0) return;
    int4 dn = make_int4(0, 0, 0, -1);
    if (PART && pl < GST_ROWS) { const int r = p0 + blockIdx.x * GST_ROWS + pl; if (r < p1) dn = prd[r]; }
    int it = 0;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
      const int stage = it % GST_STAGES;
      const int4 d = dn;
      if (PART) {                                               // descriptors of the CTA's next chunk, one chunk ahead
        dn = make_int4(0, 0, 0, -1);
        const int r = p0 + (c + (int)gridDim.x) * GST_ROWS + pl;
        if (pl < GST_ROWS && c + (int)gridDim.x < nchunks && r < p1) dn = prd[r];
      }
      if (it >= GST_STAGES) mbar_wait(&empty[stage], ((it / GST_STAGES) - 1) & 1);      // the previous use of the slot is consumed
      Stage& st = stages[stage];
      const int r0 = p0 + c * GST_ROWS, r1 = min(r0 + GST_ROWS, p1);
      const unsigned np = 16u * (unsigned)(r1 - r0);
      const unsigned ld = (r0 & 1) * 8u;
      const unsigned nd = SMOOTH ? (72u * (unsigned)(r1 - r0) + ld + 15u) & ~15u : 0u;
      if (!PART) {
        const int b0 = pia[r0], b1 = pia[r1];
        // every copy starts at the 16-byte boundary at or below its first byte and is rounded up to 16 bytes
        const unsigned lv = (b0 & 1) * 8u, lj = (b0 & 3) * 4u;
        const unsigned nv = (8u * NV * (unsigned)(b1 - b0) + lv + 15u) & ~15u;
        const unsigned nj = (4u * (unsigned)(b1 - b0) + lj + 15u) & ~15u;
        mbar_expect_tx(&full[stage], nv + nj + nd + np);
        bulk_g2s(st.val, reinterpret_cast<const unsigned char*>(val) + 8ull * NV * (unsigned long long)b0 - lv, nv, &full[stage]);
        bulk_g2s(st.ja, reinterpret_cast<const unsigned char*>(jpos) + 4ull * (unsigned long long)b0 - lj, nj, &full[stage]);
        if (SMOOTH) bulk_g2s(st.dinv, reinterpret_cast<const unsigned char*>(dinv) + 72ull * (unsigned long long)r0 - ld, nd, &full[stage]);
        bulk_g2s(st.prd, prd + r0, np, &full[stage]);
      } else {
        // my row's part: lower = blocks [s, s + nl), upper = the D + U part [s + nl, s + len)
        const int blk = LOWER ? d.x : d.x + d.z, nblk = d.w < 0 ? 0 : (LOWER ? d.z : d.y - d.z);
        const unsigned lv = (blk & 1) * 8u;
        const unsigned nv = nblk > 0 ? (8u * NV * (unsigned)nblk + lv + 15u) & ~15u : 0u;
        unsigned tot = nv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        // the chunk's column indices are one contiguous piece (first block of its first row .. end of its last row)
        const int b0 = __shfl_sync(0xffffffffu, d.x, 0);
        const int last = r1 - r0 - 1;
        const int b1 = __shfl_sync(0xffffffffu, d.x + d.y, last);
        const unsigned lj = (b0 & 3) * 4u;
        const unsigned nj = (4u * (unsigned)(b1 - b0) + lj + 15u) & ~15u;
        if (pl == 0) mbar_expect_tx(&full[stage], tot + nj + nd + np);
        __syncwarp();
        if (nv > 0) bulk_g2s(st.val + pl * GST_ROWSLOT, reinterpret_cast<const unsigned char*>(val) + 8ull * NV * (unsigned long long)blk - lv, nv, &full[stage]);
        if (pl == 0) {
          bulk_g2s(st.ja, reinterpret_cast<const unsigned char*>(jpos) + 4ull * (unsigned long long)b0 - lj, nj, &full[stage]);
          if (SMOOTH) bulk_g2s(st.dinv, reinterpret_cast<const unsigned char*>(dinv) + 72ull * (unsigned long long)r0 - ld, nd, &full[stage]);
          bulk_g2s(st.prd, prd + r0, np, &full[stage]);
        }
      }
    }
    return;
  }

  // ---- consumers: half-warp hw owns row hw of every chunk of this CTA
  const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const int cc = lane >> 2, ii = lane & 3;
  int it = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int stage = it % GST_STAGES;
    mbar_wait(&full[stage], (it / GST_STAGES) & 1);
    const Stage& st = stages[stage];
    const int r0 = p0 + c * GST_ROWS;
    const int rows = min(GST_ROWS, p1 - r0);
    const int b0 = st.prd[0].x;
    int row = -1, s = 0, len = 0, nl = 0;
    if (hw < rows) { const int4 d = st.prd[hw]; s = d.x; len = d.y; nl = d.z; row = d.w; }
    double bq = 0.0, dq = 0.0, xold = 0.0;
    if (row >= 0 && cc < 3) {
      if (SMOOTH) { if (ii < 3) { bq = b[(size_t)bs * (size_t)row + cc]; dq = st.dinv[(r0 & 1) + 9 * hw + 3 * ii + cc]; } }
      else if (MODE == S_RES || MODE == S_LDELTA) { if (ii == 0) bq = b[(size_t)bs * (size_t)row + cc]; }
    }
    if (MODE == S_GSD && row >= 0 && lane < 3) xold = x[4 * (size_t)row + lane];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    // block k = lane + 16 * round; the sums are formed exactly as in bsr_ops.cu (a += products of block k, k ascending)
#pragma unroll
    for (int rnd = 0; rnd < ML / 16; ++rnd) {
      const int k = lane + 16 * rnd;
      const bool mine = LOWER ? k < nl : UPPER ? (k > nl && k < len) : (k < len && (!(MODE == S_GS || MODE == S_GSD) || k != nl));
      if (mine) {
        const int jc = st.ja[(b0 & 3) + (s - b0) + k];
        const double* v; int sv;
        if (!PART) {
          const bool lo = k < nl;
          v = st.val + (b0 & 1) + NV * (s - b0) + (lo ? k : (NV - 1) * nl + k);
          sv = lo ? nl : len - nl;
        } else if (LOWER) { v = st.val + hw * GST_ROWSLOT + (s & 1) + k; sv = nl; }
        else { v = st.val + hw * GST_ROWSLOT + ((s + nl) & 1) + (k - nl); sv = len - nl; }
        double x0, x1, x2;
        load_row3<4>(x, jc, x0, x1, x2);
        block_mac<NV>(v, sv, x0, x1, x2, a0, a1, a2);
      }
    }
    const double a = half_sum3(a0, a1, a2, lane);
    if (!SMOOTH) {
      if (row >= 0 && ii == 0 && cc < ys) {        // cc == 3 (padded y): a == 0 and bq == 0, the pad entry stays zero
        const size_t o = (size_t)ys * (size_t)row + cc;
        if (MODE == S_SPMV) y[o] = a;
        else if (MODE == S_RES) y[o] = bq - a;
        else if (MODE == S_RESU) y[o] = -a;
        else y[o] = bq + a;
      }
    } else {
      double part = dq * (bq - a);                 // lanes outside (c < 3, i < 3): dq = 0
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 8);
      if (row >= 0 && lane < 4) {                  // lane 3: part == 0, the pad entry of the padded vectors
        const size_t o = 4 * (size_t)row + lane;
        if (MODE == S_GSD) dl[o] = lane < 3 ? part - xold : 0.0;
        y[o] = part;
      }
    }
    // Hand the slot back only HERE, after the stores that consumed every value read from it, and behind a
    // generic -> async proxy fence. Releasing right after the shared-memory loads were ISSUED (round-2 first cut)
    // let the producer's next bulk copy overwrite the stage while a load was still in flight: about one colour
    // launch in a hundred computed one row from the next chunk's numbers (caught as Krylov counts that changed
    // from run to run; tests/test_gpu_solver.py::test_streamed_passes_are_reproducible). No measurable cost.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
}

__global__ void max_rowlen_kernel(int n, const int* __restrict__ ia, int* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int m = i < n ? ia[i + 1] - ia[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

} // namespace

int max_row_length(const BsrT& A, cudaStream_t s)
{
  DevBuf<int> m; m.alloc(1, s); m.zero();
  max_rowlen_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, m.p); PNP_COUNT_LAUNCH();
  int h = 0; m.download(&h, 1);
  return h;
}

template <int MODE, int ML, int NV>
void launch_stream_ml(const BsrT& A, const double* dinv, int p0, int p1, const double* b, const double* x, double* y, double* dl, int bs, int ys, cudaStream_t s)
{
  constexpr size_t SMEM = GstCfg<ML>::STAGES * sizeof(GstStage<ML>) + 2 * GstCfg<ML>::STAGES * sizeof(unsigned long long);
  static bool configured[64] = {false};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    PNP_CUDA(cudaFuncSetAttribute(gs_stream_kernel<MODE, ML, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  static int sms = 0;
  if (sms == 0) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int nchunks = (p1 - p0 + GST_ROWS - 1) / GST_ROWS;
  const int cap = GstCfg<ML>::CTAS * sms;
  const int grid = nchunks < cap ? nchunks : cap;
  gs_stream_kernel<MODE, ML, NV><<<grid, GST_THREADS, SMEM, s>>>(p0, p1, A.prd.p, A.pia.p, A.jpos.p, NV == 7 ? A.val7.p : A.val.p, dinv, b, x, y, dl, bs, ys);
  PNP_COUNT_LAUNCH();
}

template <int MODE>
void launch_stream(const BsrT& A, const double* dinv, int p0, int p1, const double* b, const double* x, double* y, double* dl, int bs, int ys, cudaStream_t s)
{
  if (A.maxlen <= 16) {
    if (A.packed7) launch_stream_ml<MODE, 16, 7>(A, dinv, p0, p1, b, x, y, dl, bs, ys, s);
    else launch_stream_ml<MODE, 16, 9>(A, dinv, p0, p1, b, x, y, dl, bs, ys, s);
  } else {
    if (A.packed7) launch_stream_ml<MODE, 32, 7>(A, dinv, p0, p1, b, x, y, dl, bs, ys, s);
    else launch_stream_ml<MODE, 32, 9>(A, dinv, p0, p1, b, x, y, dl, bs, ys, s);
  }
}

constexpr int GST_MAXLEN = 32;          // longest row the streamed kernels take

// Does the level qualify for the streamed kernels? Read at every call so that a test can switch the path
// inside one process: PNPB200_GS_STREAM=0 keeps the plain kernels, PNPB200_GS_STREAM_MIN = smallest row
// range that is streamed (default 4096: below that a launch is latency, not bandwidth).
bool stream_ok(const BsrT& A, int nrows)
{
  const char* e_on = getenv("PNPB200_GS_STREAM"); const char* e_min = getenv("PNPB200_GS_STREAM_MIN");
  const int on = e_on ? atoi(e_on) : 1, min_rows = e_min ? atoi(e_min) : 4096;
  // Rows of 17 ... 32 blocks (coarse levels) have their own instantiation (two rounds per row, 2-stage ring), but it
  // does not pay: measured on level 1 of the 256^3 hierarchy (1.5 M rows, 11 colours): GS + delta 0.605 vs 0.649 ms,
  // forward sweep from zero 0.500 vs 0.417 ms, Krylov stage unchanged — those launches are bound by their size
  // (137 k rows each), not by load latency. PNPB200_GS_STREAM_MAXLEN=32 switches it on (tests do).
  const char* e_ml = getenv("PNPB200_GS_STREAM_MAXLEN");
  const int maxlen = e_ml ? atoi(e_ml) : 16;
  return on && VS == 4 && A.maxlen > 0 && A.maxlen <= (maxlen < GST_MAXLEN ? maxlen : GST_MAXLEN) && nrows >= min_rows;
}

// One colour launch [p0, p1) (matrix order) of a sweep. mode 0: plain, 1: forward from a zero initial guess
// (lower parts only), 2: with delta. Returns false if the level does not qualify.
bool gs_stream_colour(const Level& L, int p0, int p1, const double* b, double* x, double* dl, int mode, int bs, cudaStream_t s)
{
  if (!stream_ok(L.A, p1 - p0)) return false;
  if (mode == 2) launch_stream<S_GSD>(L.A, L.dinv.p, p0, p1, b, x, x, dl, bs, VS, s);
  else if (mode == 1) launch_stream<S_GS0>(L.A, L.dinv.p, p0, p1, b, x, x, nullptr, bs, VS, s);
  else launch_stream<S_GS>(L.A, L.dinv.p, p0, p1, b, x, x, nullptr, bs, VS, s);
  return true;
}

// Natural passes over the SWEPT rows of a level (matrix positions [0, p1): every row but the ghost rows of a
// multi-GPU level), which: 0 y = A x, 1 y = b - A x, 4 y = -U x (residual right after a forward sweep
// from zero), 6 y = b + L x. The caller runs the plain kernel on the rows behind p1.
bool gs_stream_pass(const BsrT& A, int which, int p1, const double* b, const double* x, double* y, int bs, int ys, cudaStream_t s)
{
  if (!stream_ok(A, p1)) return false;
  // The half-row passes in natural order stay on the software-pipelined plain kernel by default: their
  // producer issues one ~500-byte copy per row, and at that size the copy engine is slower than the LSU
  // path (measured at 256^3: r = -U x 2.69 vs 2.30 ms, y = b + L x 2.62 vs 2.26 ms), while the forward
  // sweep from zero gains (2.81 vs 3.09 ms: its colour launches were latency bound). PNPB200_GS_STREAM_PART=1
  // streams them as well.
  if (which == 4 || which == 6) { const char* e = getenv("PNPB200_GS_STREAM_PART"); if (!e || !atoi(e)) return false; }
  switch (which) {
    case 0: launch_stream<S_SPMV>(A, nullptr, 0, p1, b, x, y, nullptr, bs, ys, s); break;
    case 1: launch_stream<S_RES>(A, nullptr, 0, p1, b, x, y, nullptr, bs, ys, s); break;
    case 4: launch_stream<S_RESU>(A, nullptr, 0, p1, b, x, y, nullptr, bs, ys, s); break;
    case 6: launch_stream<S_LDELTA>(A, nullptr, 0, p1, b, x, y, nullptr, bs, ys, s); break;
    default: return false;
  }
  return true;
}

} // namespace pnpb200
