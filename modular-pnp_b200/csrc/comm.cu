// pnp-b200 — multi-GPU layer: one process per GPU, block rows partitioned by vertex,
// ghost vertices appended after the owned ones, halo exchange of 3-vectors with grouped
// ncclSend/ncclRecv and fp64 allreduce of the fused dot-product batches. NCCL is resolved at run
// time (dlopen) so that the process hosting us (bench.py under torchrun, which has loaded torch's
// libnccl.so.2 already) does not end up with two copies.
#include "internal.hpp"
#include <dlfcn.h>
#include <cstring>

namespace pnpb200 {

namespace {
struct ncclUniqueId_t { char internal[128]; };
typedef int (*fn_getid)(ncclUniqueId_t*);
typedef int (*fn_init)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_csendrecv)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
constexpr int NCCL_DOUBLE = 8, NCCL_INT32 = 2, NCCL_INT8 = 0, NCCL_SUM = 0, NCCL_MAX = 2;

struct Api {
  void* lib = nullptr;
  fn_getid getid; fn_init init; fn_destroy destroy; fn_allreduce allreduce; fn_csendrecv send; fn_sendrecv recv;
  fn_group gstart, gend; fn_allgather allgather;
};
Api& api()
{
  static Api a;
  if (!a.lib) {
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) throw Error(ERROR_DEVICE, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) { void* f = dlsym(a.lib, n); if (!f) throw Error(ERROR_DEVICE, std::string("NCCL symbol missing: ") + n); return f; };
    a.getid = (fn_getid)sym("ncclGetUniqueId"); a.init = (fn_init)sym("ncclCommInitRank");
    a.destroy = (fn_destroy)sym("ncclCommDestroy"); a.allreduce = (fn_allreduce)sym("ncclAllReduce");
    a.send = (fn_csendrecv)sym("ncclSend"); a.recv = (fn_sendrecv)sym("ncclRecv");
    a.gstart = (fn_group)sym("ncclGroupStart"); a.gend = (fn_group)sym("ncclGroupEnd");
    a.allgather = (fn_allgather)sym("ncclAllGather");
  }
  return a;
}
inline void nccl_check(int rc, const char* what)
{
  if (rc != 0) throw Error(ERROR_DEVICE, std::string("NCCL failure in ") + what + " (code " + std::to_string(rc) + ")");
}

// sendbuf[XS*q + k] = vec[XS*idx[q] + k]   (XS doubles per block row: 3 in row-id space, VS in sweep space)
template <int XS>
__global__ void pack_rows_kernel(int n, const int* __restrict__ idx, const double* __restrict__ vec, double* __restrict__ buf)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= XS * n) return;
  buf[e] = vec[(size_t)XS * (size_t)idx[e / XS] + e % XS];
}
} // namespace

Comm::~Comm() { if (nccl) { try { api().destroy(nccl); } catch (...) {} } }

void comm_allreduce(Handle& h, double* dev, int count, bool is_max)
{
  if (!h.comm.active()) return;
  nccl_check(api().allreduce(dev, dev, (size_t)count, NCCL_DOUBLE, is_max ? NCCL_MAX : NCCL_SUM, h.comm.nccl, h.stream), "allreduce");
}
void comm_allreduce_int(Handle& h, int* dev, int count, bool is_max)
{
  if (!h.comm.active()) return;
  nccl_check(api().allreduce(dev, dev, (size_t)count, NCCL_INT32, is_max ? NCCL_MAX : NCCL_SUM, h.comm.nccl, h.stream), "allreduce(int)");
}
void comm_allgather(Handle& h, const void* send, void* recv, size_t bytes_per_rank)
{
  if (!h.comm.active()) { PNP_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, h.stream)); return; }
  nccl_check(api().allgather(send, recv, bytes_per_rank, NCCL_INT8, h.comm.nccl, h.stream), "allgather");
}

__global__ void pack1_kernel(int n, const int* __restrict__ idx, const int* __restrict__ vec, int* __restrict__ buf)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) buf[e] = vec[idx[e]];
}

// refresh the ghost entries of a block vector (3 doubles per row) from their owners
void halo_exchange(Handle& h, Halo& c, double* vec, bool sweep_space)
{
  if (!h.comm.active() || !c.active()) return;
  cudaStream_t s = h.stream;
  const int nsend = c.send_ptr.back();
  if (sweep_space && nsend > 0 && !c.send_pos.p) throw Error(ERROR_UNKNOWN, "halo_exchange: level has no sweep-space send list");
  const int* idx = sweep_space ? c.send_pos.p : c.send_idx.p;      // ghost rows are the tail in both orders
  const size_t xs = sweep_space ? VS : 3;      // doubles per block row of this vector
  if (c.sendbuf.n < xs * (size_t)(nsend ? nsend : 1)) c.sendbuf.alloc(xs * (size_t)(nsend ? nsend : 1), s);
  if (nsend > 0) {
    if (sweep_space) pack_rows_kernel<VS><<<cdiv(VS * (size_t)nsend, 256), 256, 0, s>>>(nsend, idx, vec, c.sendbuf.p);
    else pack_rows_kernel<3><<<cdiv(3 * (size_t)nsend, 256), 256, 0, s>>>(nsend, idx, vec, c.sendbuf.p);
    PNP_COUNT_LAUNCH();
  }
  Api& a = api();
  nccl_check(a.gstart(), "groupStart");
  for (size_t k = 0; k < c.nbr.size(); ++k) {
    const int ns = c.send_ptr[k + 1] - c.send_ptr[k], nr = c.recv_ptr[k + 1] - c.recv_ptr[k];
    if (ns > 0) nccl_check(a.send(c.sendbuf.p + xs * (size_t)c.send_ptr[k], xs * (size_t)ns, NCCL_DOUBLE, c.nbr[k], h.comm.nccl, s), "send");
    if (nr > 0) nccl_check(a.recv(vec + xs * (size_t)c.recv_ptr[k], xs * (size_t)nr, NCCL_DOUBLE, c.nbr[k], h.comm.nccl, s), "recv");
  }
  nccl_check(a.gend(), "groupEnd");
}

void halo_exchange_int(Handle& h, Halo& c, int* vec)
{
  if (!h.comm.active() || !c.active()) return;
  cudaStream_t s = h.stream;
  const int nsend = c.send_ptr.back();
  c.isendbuf.alloc(nsend ? nsend : 1, s);
  if (nsend > 0) { pack1_kernel<<<cdiv(nsend, 256), 256, 0, s>>>(nsend, c.send_idx.p, vec, c.isendbuf.p); PNP_COUNT_LAUNCH(); }
  Api& a = api();
  nccl_check(a.gstart(), "groupStart");
  for (size_t k = 0; k < c.nbr.size(); ++k) {
    const int ns = c.send_ptr[k + 1] - c.send_ptr[k], nr = c.recv_ptr[k + 1] - c.recv_ptr[k];
    if (ns > 0) nccl_check(a.send(c.isendbuf.p + c.send_ptr[k], (size_t)ns, NCCL_INT32, c.nbr[k], h.comm.nccl, s), "send(int)");
    if (nr > 0) nccl_check(a.recv(vec + c.recv_ptr[k], (size_t)nr, NCCL_INT32, c.nbr[k], h.comm.nccl, s), "recv(int)");
  }
  nccl_check(a.gend(), "groupEnd");
}

} // namespace pnpb200

using namespace pnpb200;
struct pnpb200_handle { Handle h; };

extern "C" {

int pnpb200_comm_unique_id(unsigned char id[128])
{
  try {
    ncclUniqueId_t u;
    nccl_check(api().getid(&u), "ncclGetUniqueId");
    memcpy(id, u.internal, 128);
    return FASP_SUCCESS;
  } catch (const Error& e) { fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return e.code; }
}

int pnpb200_comm_init(pnpb200_handle* ph, const unsigned char id[128], int rank, int n_ranks)
{
  if (!ph || !id || rank < 0 || rank >= n_ranks) return ERROR_INPUT_PAR;
  try {
    Handle& h = ph->h;
    PNP_CUDA(cudaSetDevice(h.device));
    ncclUniqueId_t u; memcpy(u.internal, id, 128);
    nccl_check(api().init(&h.comm.nccl, n_ranks, u, rank), "ncclCommInitRank");
    h.comm.rank = rank; h.comm.nranks = n_ranks;
    return FASP_SUCCESS;
  } catch (const Error& e) { fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return e.code; }
}

int pnpb200_set_partition(pnpb200_handle* ph, int n_owned, int n_neighbors, const int* neighbor_ranks,
                          const int* send_ptr, const int* send_idx, const int* recv_ptr)
{
  if (!ph || n_owned < 1 || n_neighbors < 0) return ERROR_INPUT_PAR;
  try {
    Handle& h = ph->h; cudaStream_t s = h.stream;
    PNP_CUDA(cudaSetDevice(h.device));
    if (n_owned > h.nv) throw Error(ERROR_INPUT_PAR, "n_owned exceeds the local vertex count");
    if (n_neighbors > 0 && (!neighbor_ranks || !send_ptr || !recv_ptr)) throw Error(ERROR_INPUT_PAR, "neighbour lists missing");
    if (n_neighbors > 0 && !h.comm.active()) throw Error(ERROR_INPUT_PAR, "pnpb200_set_partition with neighbours needs pnpb200_comm_init first");
    if (n_neighbors > 0 && send_ptr[0] != 0) throw Error(ERROR_INPUT_PAR, "send_ptr must start at 0");
    for (int k = 0; k < n_neighbors; ++k) {
      if (neighbor_ranks[k] < 0 || neighbor_ranks[k] >= h.comm.nranks || neighbor_ranks[k] == h.comm.rank)
        throw Error(ERROR_INPUT_PAR, "neighbour rank out of range (or this rank itself)");
      for (int q = 0; q < k; ++q) if (neighbor_ranks[q] == neighbor_ranks[k]) throw Error(ERROR_INPUT_PAR, "neighbour rank listed twice");
      if (send_ptr[k + 1] < send_ptr[k] || recv_ptr[k + 1] < recv_ptr[k]) throw Error(ERROR_INPUT_PAR, "send_ptr / recv_ptr must be non-decreasing");
    }
    if (n_neighbors > 0 && send_ptr[n_neighbors] > 0 && !send_idx) throw Error(ERROR_INPUT_PAR, "send_idx missing");
    if (n_neighbors > 0 && (recv_ptr[0] != n_owned || recv_ptr[n_neighbors] != h.nv))
      throw Error(ERROR_INPUT_PAR, "the ghost ranges must tile [n_owned, n_local)");
    if (n_neighbors == 0 && n_owned != h.nv) throw Error(ERROR_INPUT_PAR, "ghost vertices without a neighbour");
    Level& F = *h.hier.lv[0];
    Halo& c = F.halo;
    F.n_own = n_owned;
    c.nbr.assign(neighbor_ranks, neighbor_ranks + n_neighbors);
    c.send_ptr.assign(send_ptr, send_ptr + n_neighbors + 1);
    c.recv_ptr.assign(recv_ptr, recv_ptr + n_neighbors + 1);
    const int nsend = n_neighbors ? c.send_ptr.back() : 0;
    for (int q = 0; q < nsend; ++q) if (send_idx[q] < 0 || send_idx[q] >= n_owned) throw Error(ERROR_INPUT_PAR, "send index is not an owned vertex");
    for (int k = 0; k <= n_neighbors; ++k) if (c.recv_ptr[k] < n_owned || c.recv_ptr[k] > h.nv) throw Error(ERROR_INPUT_PAR, "ghost range outside [n_owned, n_local]");
    c.send_idx.alloc(nsend ? nsend : 1, s);
    if (nsend) c.send_idx.upload(send_idx, nsend);
    c.sendbuf.alloc(4 * (size_t)(nsend ? nsend : 1), s);
    // ghost rows must not be part of the Gauss-Seidel colour groups: lay the fine level out again
    relayout_fine_level(h);
    // ghost dofs become identity rows of the local matrix (like Dirichlet rows)
    const size_t N = 3 * (size_t)h.nv;
    std::vector<unsigned char> f = h.bc_user;
    f.resize(N, 0);
    for (size_t i = 3 * (size_t)n_owned; i < N; ++i) f[i] = 1;
    h.bcflag.upload(f.data(), N);
    h.rhs_current = false;
    // owner rank of every local vertex: a cell that several ranks hold (ghost layer) belongs to the
    // lowest owner rank among its vertices, a rule every holder evaluates alike (error_norms)
    std::vector<int> own((size_t)h.nv, h.comm.rank);
    for (int k = 0; k < n_neighbors; ++k) for (int g = c.recv_ptr[k]; g < c.recv_ptr[k + 1]; ++g) own[g] = c.nbr[k];
    h.vowner.alloc(h.nv, s); h.vowner.upload(own.data(), h.nv);
    PNP_CUDA(cudaStreamSynchronize(s));
    return FASP_SUCCESS;
  } catch (const Error& e) { fprintf(stderr, "### pnpb200 ERROR: %s\n", e.what()); return e.code; }
}

}
