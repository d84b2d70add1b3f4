// pnp-b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace pnpb200 {

// doubles per block row in the solver's SWEEP-SPACE vectors (internal.hpp): 4 = padded to one
// 32-byte sector per row (the pad entry is kept at zero), 3 = packed
#ifndef PNPB200_VSTRIDE
#define PNPB200_VSTRIDE 4
#endif
constexpr int VS = PNPB200_VSTRIDE;

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PNP_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      char buf__[512];                                                                     \
      snprintf(buf__, sizeof buf__, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), \
               __FILE__, __LINE__, #call);                                                 \
      throw ::pnpb200::Error(-120, buf__);                                                 \
    }                                                                                      \
  } while (0)

#define PNP_LAUNCH_CHECK() PNP_CUDA(cudaGetLastError())

// launch counter (bench.py's gpu_launches claim)
extern thread_local long g_launches;
struct Error;
extern int g_debug_sync;      // PNPB200_DEBUG_SYNC=1: synchronise + check after every launch
inline void count_launch(const char* file, int line)
{
  ++g_launches;
  if (g_debug_sync) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      char buf[512];
      snprintf(buf, sizeof buf, "CUDA error %s after the launch at %s:%d", cudaGetErrorString(e), file, line);
      throw Error(-120, buf);
    }
  }
}
#define PNP_COUNT_LAUNCH() ::pnpb200::count_launch(__FILE__, __LINE__)

inline int cdiv(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// Stream-ordered device buffer (cudaMallocAsync pool: hierarchy rebuilds every Newton step
// re-use the same blocks without driver round trips).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept
  {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count, cudaStream_t stream)
  {
    if (p && n >= count && s == stream) return;     // keep a big-enough block
    release();
    count += count / 16;                            // slack: sizes drift a little between Newton steps
    s = stream; n = count;
    if (count == 0) return;
    PNP_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), stream));
  }
  void release()
  {
    if (p) cudaFreeAsync(p, s);
    p = nullptr; n = 0;
  }
  void zero() { if (p) PNP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count) { PNP_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void download(T* h, size_t count) const
  {
    PNP_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<T> to_host(size_t count) const { std::vector<T> v(count); if (count) download(v.data(), count); return v; }
};

// Bump allocator over one persistent device block: the AMG setup's large temporaries (sort keys,
// strength flags, CUB scratch) come from here so that a hierarchy rebuild never goes back to the
// driver for memory.
struct Workspace {
  DevBuf<unsigned char> buf;
  size_t off = 0;
  void reserve(size_t bytes, cudaStream_t s)
  {
    if (buf.n >= bytes) return;
    PNP_CUDA(cudaStreamSynchronize(s));
    buf.release();
    buf.alloc(bytes, s);
  }
  void reset() { off = 0; }
  template <typename T> T* get(size_t count)
  {
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > buf.n) throw Error(-15, "pnpb200 workspace exhausted");
    T* r = reinterpret_cast<T*>(buf.p + off);
    off += bytes;
    return r;
  }
};

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the 16 lanes of a half-warp group (all 32 lanes must participate)
__device__ __forceinline__ double half_sum(double v)
{
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sums of a0, a1, a2 over the 16 lanes of a half-warp with 5 instead of 12 double shuffles (the
// shuffles share the L1 data pipe with the loads, which ncu shows at 55-75 % in the row kernels):
// lane L of the half-warp returns the total of a_c, c = (L & 15) >> 2 (c = 3: 0). Every step halves
// the lanes a value lives on while it doubles the lanes it has been summed over.
__device__ __forceinline__ double half_sum3(double a0, double a1, double a2, int lane)
{
  const bool hi = (lane & 8) != 0;
  const double r1 = __shfl_xor_sync(0xffffffffu, hi ? a0 : a2, 8);
  const double r2 = __shfl_xor_sync(0xffffffffu, hi ? a1 : 0.0, 8);
  const double p = (hi ? a2 : a0) + r1;
  const double q = hi ? 0.0 : a1 + r2;
  const bool b2 = (lane & 4) != 0;
  double v = (b2 ? q : p) + __shfl_xor_sync(0xffffffffu, b2 ? p : q, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum(double v, double* smem32)
{
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem32[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}
__device__ __forceinline__ double block_max(double v, double* smem32)
{
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem32[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.0;
  if (w == 0) v = warp_max(v);
  return v;
}
// BSR-T/LDU addressing (internal.hpp): a row that starts at block s with len blocks, of which the
// first nl are "lower" (their columns are swept before the row), stores the lower part
// component-major with stride nl and the diagonal + upper part component-major with stride len-nl.
// Returns the offset of component 0 of block k (k-th block of the row, physical order) and the
// component stride: component t lives at val[off + t*stride].
#ifndef PNPB200_SPLIT_ROWS
#define PNPB200_SPLIT_ROWS 1
#endif
__device__ __forceinline__ size_t bsrt_off(int s, int len, int nl, int k, int& stride)
{
#if PNPB200_SPLIT_ROWS
  if (k < nl) { stride = nl; return 9 * (size_t)s + (size_t)k; }
  stride = len - nl;
  return 9 * (size_t)s + 9 * (size_t)nl + (size_t)(k - nl);
#else
  (void)nl;
  stride = len;
  return 9 * (size_t)s + (size_t)k;
#endif
}
// a += (3x3 block) * x for one stored block: v = its first value inside the row part, st = the part's component stride.
// NV = 9: all nine entries. NV = 7: the PNP block structure — the linearised PNP system has no cation-anion coupling,
// entries (1,2) and (2,1) of EVERY block are structurally zero on every level (unsmoothed aggregation sums blocks) —
// stored as (0,0) (0,1) (0,2) (1,0) (1,1) (2,0) (2,2). Explicit FMA chains in a fixed order, so that the two forms
// (and every kernel that uses them) give the same bits when the two entries are zero.
template <int NV>
__device__ __forceinline__ void block_mac(const double* v, int st, double x0, double x1, double x2, double& a0, double& a1, double& a2)
{
  a0 = __fma_rn(v[2 * st], x2, __fma_rn(v[st], x1, __fma_rn(v[0], x0, a0)));
  if (NV == 9) {
    a1 = __fma_rn(v[5 * st], x2, __fma_rn(v[4 * st], x1, __fma_rn(v[3 * st], x0, a1)));
    a2 = __fma_rn(v[8 * st], x2, __fma_rn(v[7 * st], x1, __fma_rn(v[6 * st], x0, a2)));
  } else {
    a1 = __fma_rn(v[4 * st], x1, __fma_rn(v[3 * st], x0, a1));
    a2 = __fma_rn(v[6 * st], x2, __fma_rn(v[5 * st], x0, a2));
  }
}
// gather the three components of block row j from a vector with XS doubles per row. Sweep-space
// vectors are padded to XS = 4 doubles (32 bytes, one DRAM sector) per row so that the gather is ONE
// 256-bit load (LDG.E.256 on sm_100a) instead of three 8-byte loads that each walk the same 16
// scattered sectors of a half-warp: the gathers were half of the L1 data-pipe wavefronts of the row
// kernels (profiles/README.md, round 2). Row-id vectors (C ABI, assembly) stay packed (XS = 3).
template <int XS>
__device__ __forceinline__ void load_row3(const double* x, int j, double& x0, double& x1, double& x2)
{
  if (XS == 4) {
    double pad;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(pad) : "l"(x + 4 * (size_t)j) : "memory");
  } else {
    const double* p = x + (size_t)XS * (size_t)j;
    x0 = p[0]; x1 = p[1]; x2 = p[2];
  }
}
__host__ __device__ __forceinline__ unsigned hash32(unsigned x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
#endif

} // namespace pnpb200
