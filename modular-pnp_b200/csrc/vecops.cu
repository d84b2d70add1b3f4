// pnp-b200 — dense vector kernels for the Krylov solver:
// fused multi-dot (all Gram-Schmidt projections of one Arnoldi step in a single pass over the
// basis, block + warp-shuffle reductions, deterministic two-stage sum) and fused multi-axpy
// with the norm of the result. fasp_blas_array_* [EXT FASP] equivalents.
#include "internal.hpp"

namespace pnpb200 {

namespace {

constexpr int VB = 256;          // threads per block
constexpr int DOT_GROUP = 8;     // dot products accumulated per pass over w

__global__ void set_kernel(double* __restrict__ x, double v, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void axpy_kernel(double a, const double* __restrict__ x, double* __restrict__ y, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void scale_kernel(double a, double* __restrict__ x, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= a;
}

// partial[(g0+q)*nblk + blk] = sum over this block's slice of V_{g0+q} . w   (q < cnt);
// if WITH_WW also the slice of w.w at index nvec
template <bool WITH_WW>
__global__ void __launch_bounds__(VB)
multi_dot_kernel(const double* __restrict__ V, size_t ld, int g0, int cnt, int nvec,
                 const double* __restrict__ w, size_t n, double* __restrict__ partial)
{
  __shared__ double sm[32];
  double acc[DOT_GROUP], ww = 0.0;
#pragma unroll
  for (int q = 0; q < DOT_GROUP; ++q) acc[q] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
    if (WITH_WW) ww += wi * wi;
#pragma unroll
    for (int q = 0; q < DOT_GROUP; ++q)
      if (q < cnt) acc[q] += V[(size_t)(g0 + q) * ld + i] * wi;
  }
#pragma unroll
  for (int q = 0; q < DOT_GROUP; ++q) {
    if (q < cnt) {
      const double r = block_sum(acc[q], sm);
      if (threadIdx.x == 0) partial[(size_t)(g0 + q) * gridDim.x + blockIdx.x] = r;
    }
  }
  if (WITH_WW) {
    const double r = block_sum(ww, sm);
    if (threadIdx.x == 0) partial[(size_t)nvec * gridDim.x + blockIdx.x] = r;
  }
}

// out[q] = sum_blk partial[q*nblk + blk]; one block per q
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, int nblk,
                                                             double* __restrict__ out)
{
  __shared__ double sm[32];
  double a = 0.0;
  for (int t = threadIdx.x; t < nblk; t += blockDim.x) a += partial[(size_t)blockIdx.x * nblk + t];
  a = block_sum(a, sm);
  if (threadIdx.x == 0) out[blockIdx.x] = a;
}

struct Coefs { double c[32]; };

// w += sum_i c_i V_i ; optional partial of ||w||^2
template <bool WITH_NORM>
__global__ void __launch_bounds__(VB)
multi_axpy_kernel(const double* __restrict__ V, size_t ld, int nvec, Coefs cf, double* __restrict__ w, size_t n,
                  double* __restrict__ partial)
{
  __shared__ double sm[32];
  double ss = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double a = w[i];
    for (int q = 0; q < nvec; ++q) a += cf.c[q] * V[(size_t)q * ld + i];
    w[i] = a;
    if (WITH_NORM) ss += a * a;
  }
  if (WITH_NORM) {
    ss = block_sum(ss, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = ss;
  }
}

inline int vec_grid(size_t n)
{
  const size_t b = (n + VB - 1) / VB;
  const size_t cap = 148 * 8;            // 8 resident CTAs of 256 threads per SM
  return (int)(b < cap ? (b ? b : 1) : cap);
}

} // namespace

void vec_set(double* x, double v, size_t n, cudaStream_t s)
{
  if (v == 0.0) { PNP_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), s)); return; }
  set_kernel<<<vec_grid(n), VB, 0, s>>>(x, v, n); PNP_COUNT_LAUNCH();
}
void vec_copy(double* dst, const double* src, size_t n, cudaStream_t s)
{
  PNP_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void vec_axpy(double a, const double* x, double* y, size_t n, cudaStream_t s)
{
  axpy_kernel<<<vec_grid(n), VB, 0, s>>>(a, x, y, n); PNP_COUNT_LAUNCH();
}
void vec_scale(double a, double* x, size_t n, cudaStream_t s)
{
  scale_kernel<<<vec_grid(n), VB, 0, s>>>(a, x, n); PNP_COUNT_LAUNCH();
}

namespace {
__global__ void restride_kernel(size_t nrows, const double* __restrict__ in, int is, double* __restrict__ out, int os)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nrows * (size_t)os) return;
  const size_t r = e / os; const int k = (int)(e % os);
  out[e] = k < is && k < 3 ? in[r * is + k] : 0.0;
}
}
void vec_restride(const double* in, int is, double* out, int os, size_t nrows, cudaStream_t s)
{
  if (is == os) { vec_copy(out, in, nrows * (size_t)is, s); return; }
  restride_kernel<<<cdiv(nrows * (size_t)os, 256), 256, 0, s>>>(nrows, in, is, out, os); PNP_COUNT_LAUNCH();
}

void multi_dot(Handle& h, const double* V, size_t ld, int nvec, const double* w, size_t n, double* out_host)
{
  cudaStream_t s = h.stream;
  const int g = vec_grid(n);
  h.red.alloc((size_t)(nvec + 1) * g + 64, s);
  double* partial = h.red.p + 64;
  if (nvec == 0) {
    multi_dot_kernel<true><<<g, VB, 0, s>>>(V, ld, 0, 0, 0, w, n, partial); PNP_COUNT_LAUNCH();
  }
  for (int g0 = 0; g0 < nvec; g0 += DOT_GROUP) {
    const int cnt = nvec - g0 < DOT_GROUP ? nvec - g0 : DOT_GROUP;
    if (g0 == 0) multi_dot_kernel<true><<<g, VB, 0, s>>>(V, ld, g0, cnt, nvec, w, n, partial);
    else multi_dot_kernel<false><<<g, VB, 0, s>>>(V, ld, g0, cnt, nvec, w, n, partial);
    PNP_COUNT_LAUNCH();
  }
  reduce_partials_kernel<<<nvec + 1, 256, 0, s>>>(partial, g, h.red.p); PNP_COUNT_LAUNCH();
  comm_allreduce(h, h.red.p, nvec + 1, false);   // one allreduce for the whole Gram-Schmidt batch
  PNP_CUDA(cudaMemcpyAsync(h.h_red, h.red.p, (nvec + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i <= nvec; ++i) out_host[i] = h.h_red[i];
}

void multi_axpy(Handle& h, const double* V, size_t ld, int nvec, const double* coef_host, double* w, size_t n,
                double* norm2_out)
{
  cudaStream_t s = h.stream;
  if (nvec > 32) throw Error(ERROR_INPUT_PAR, "multi_axpy: more than 32 vectors");
  Coefs cf;
  for (int i = 0; i < 32; ++i) cf.c[i] = i < nvec ? coef_host[i] : 0.0;
  const int g = vec_grid(n);
  if (norm2_out) {
    h.red.alloc((size_t)g + 64, s);
    multi_axpy_kernel<true><<<g, VB, 0, s>>>(V, ld, nvec, cf, w, n, h.red.p + 64); PNP_COUNT_LAUNCH();
    reduce_partials_kernel<<<1, 256, 0, s>>>(h.red.p + 64, g, h.red.p); PNP_COUNT_LAUNCH();
    comm_allreduce(h, h.red.p, 1, false);
    PNP_CUDA(cudaMemcpyAsync(h.h_red, h.red.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    PNP_CUDA(cudaStreamSynchronize(s));
    *norm2_out = h.h_red[0];
  } else {
    multi_axpy_kernel<false><<<g, VB, 0, s>>>(V, ld, nvec, cf, w, n, nullptr); PNP_COUNT_LAUNCH();
  }
}

double vec_norm2(Handle& h, const double* x, size_t n)
{
  double out[1];
  multi_dot(h, x, 0, 0, x, n, out);
  return sqrt(out[0]);
}

} // namespace pnpb200
