// pnp-b200 — multi-GPU layer: one process per GPU, block rows partitioned by vertex,
// ghost vertices appended after the owned ones, halo exchange of 3-vectors with grouped
// ncclSend/ncclRecv and fp64 allreduce of the fused dot-product batches. NCCL is resolved at run
// time (dlopen) so that the process hosting us (bench.py under torchrun, which has loaded torch's
// libnccl.so.2 already) does not end up with two copies.
//
// Peer-memory halo exchange (round 2, SURVEY §2.3 "fused with the collective that follows"): on one node every GPU
// reaches every other through NVLink / NVSwitch, so the exchange of a level's boundary rows is ONE kernel per rank
// (halo_ipc_kernel): it packs the rows of each neighbour's send list, stores them straight into that neighbour's
// receive slot (memory exported with CUDA IPC, mapped once per communicator), publishes an epoch flag there with
// system scope, waits for the neighbours' flags in its own slot and copies the received rows into the ghost tail of
// the vector. No NCCL call, no rendezvous on the host, one launch instead of pack + grouped send / recv; it is captured
// into the coarse-level CUDA graph like any other kernel. NCCL stays the fallback (PNPB200_HALO_IPC=0, more than 32
// neighbours, slot too small, IPC unavailable) and the referee: each level's first exchanges are run through both
// paths and compared before the peer path is trusted.
#include "internal.hpp"
#include <dlfcn.h>
#include <cstring>
#include <algorithm>

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

// ---- peer-memory halo exchange ------------------------------------------------------------------------------------
constexpr int IPC_NSLOTS = 6;                   // levels 0 .. 5 (deeper distributed levels keep NCCL)
constexpr size_t IPC_CTL_BYTES = 1024;          // 128 x 8 bytes of flags / counters in front of every slot
}  // namespace

// one allocation per communicator, carved into per-level slots of equal size on every rank:
//   slot = [ctl: 128 x u64][receive buffer: 2 halves x cap rows x 4 doubles]
struct IpcArena {
  unsigned char* base = nullptr;
  size_t slot_bytes = 0; int cap = 0;
  std::vector<unsigned char*> peer;             // per rank: this process' mapping of that rank's arena (nullptr = self)
  bool failed = false;
  int checks[IPC_NSLOTS] = {0, 0, 0, 0, 0, 0};  // exchanges of the slot verified against NCCL so far
  ~IpcArena()
  {
    for (unsigned char* q : peer) if (q) cudaIpcCloseMemHandle(q);
    if (base) cudaFree(base);
  }
};

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// One launch = one halo exchange of a level. CTAs [blk_beg[k], blk_beg[k+1]) serve neighbour k:
//   phase A  pack this rank's boundary rows for k and store them into k's receive slot (half = epoch parity) over
//            NVLink; the group's last CTA publishes the epoch in k's flag word (release, system scope)
//   phase B  wait until k's epoch has arrived in this rank's flag word (acquire), copy k's rows into the ghost tail
// Every CTA finishes its stores before it waits and the grid is far below the resident-CTA limit, so no CTA waits for
// one that cannot run. Two halves suffice: a neighbour can be at most one exchange ahead (it needs this rank's flag of
// exchange e to finish e), and then it writes the other half. A wait longer than 5 s sets the error word instead of
// hanging the device (comm_check_ipc_error).
template <int XS>
__global__ void __launch_bounds__(256) halo_ipc_kernel(HaloIpcArgs a, const int* __restrict__ send_idx, double* __restrict__ vec)
{
  unsigned long long* ctl = a.ctl;
  __shared__ unsigned long long s_epoch;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(ctl + 32) + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  int k = 0;
  while ((int)blockIdx.x >= a.blk_beg[k + 1]) ++k;
  const int nb = a.blk_beg[k + 1] - a.blk_beg[k], b = (int)blockIdx.x - a.blk_beg[k];
  const size_t half = (size_t)(e & 1) * (size_t)a.cap * 4;
  {
    double* dst = a.peer_recv[k] + half + (size_t)XS * a.peer_goff[k];
    const int s0 = a.send_beg[k], cnt = XS * (a.send_beg[k + 1] - s0);
    for (int i = b * blockDim.x + threadIdx.x; i < cnt; i += nb * blockDim.x)
      dst[i] = vec[(size_t)XS * (size_t)send_idx[s0 + i / XS] + i % XS];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long sent = atomicAdd(ctl + 64 + k, 1ull) + 1;
    if (sent == (unsigned long long)nb) { ctl[64 + k] = 0; __threadfence_system(); st_release_sys(a.peer_flag[k], e); }
    int ok = 1;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(ctl + k) < e) {
      if (global_ns() - t0 > 5000000000ull) { ok = 0; atomicExch(ctl + 34, 1ull); break; }
      __nanosleep(100);
    }
    s_ok = ok;
  }
  __syncthreads();
  if (s_ok) {
    const double* src = a.recv + half + (size_t)XS * a.recv_off[k];
    double* out = vec + (size_t)XS * (size_t)(a.ghost0 + a.recv_off[k]);
    const int cnt = XS * a.recv_cnt[k];
    for (int i = b * blockDim.x + threadIdx.x; i < cnt; i += nb * blockDim.x) out[i] = __ldcg(src + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(ctl + 33, 1ull) + 1;
    if (done == (unsigned long long)gridDim.x) { ctl[33] = 0; __threadfence(); st_release_sys(ctl + 32, e); }
  }
}

__global__ void ipc_fill_kernel(int n, int n_own, int rank, int xs, double* v)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * xs) v[i] = (i / xs < n_own) ? 1.0e7 * (rank + 1) + 3.0 * (i / xs) + (i % xs) * 0.25 : -1.0;
}
__global__ void ipc_compare_kernel(size_t cnt, const double* a, const double* b, int* bad)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt && a[i] != b[i]) atomicExch(bad, 1);
}

bool ipc_enabled()
{
  const char* e = getenv("PNPB200_HALO_IPC");
  return !e || atoi(e) != 0;
}

void halo_exchange_nccl(Handle& h, Halo& c, double* vec, bool sweep_space);

// every rank agrees (min over ranks) on a yes / no decision taken on the host
bool all_agree(Handle& h, bool mine)
{
  DevBuf<int> d; d.alloc(1, h.stream);
  int v = mine ? 1 : 0;
  d.upload(&v, 1);
  nccl_check(api().allreduce(d.p, d.p, 1, NCCL_INT32, 1 /* ncclProd */, h.comm.nccl, h.stream), "allreduce(agree)");
  d.download(&v, 1);
  return v != 0;
}

// collective, once per communicator: size the slots for the finest level's ghost rows, export, exchange and map
void ipc_arena_init(Handle& h, int ghost_rows_level0)
{
  Comm& cm = h.comm;
  cm.ipc = new IpcArena;
  IpcArena& A = *cm.ipc;
  cudaStream_t s = h.stream;
  DevBuf<int> d; d.alloc(1, s);
  d.upload(&ghost_rows_level0, 1);
  comm_allreduce_int(h, d.p, 1, true);
  int gmax = 0; d.download(&gmax, 1);
  A.cap = ((gmax + gmax / 8 + 255) / 256) * 256;
  A.slot_bytes = IPC_CTL_BYTES + 2 * (size_t)A.cap * 4 * sizeof(double);
  cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof mine);
  bool ok = cudaMalloc((void**)&A.base, IPC_NSLOTS * A.slot_bytes) == cudaSuccess;
  if (!ok) { A.base = nullptr; cudaGetLastError(); }
  if (ok) ok = cudaMemset(A.base, 0, IPC_NSLOTS * A.slot_bytes) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine, A.base) == cudaSuccess;
  if (!ok) cudaGetLastError();
  DevBuf<unsigned char> hs, hr; hs.alloc(sizeof mine, s); hr.alloc(sizeof mine * (size_t)cm.nranks, s);
  hs.upload((const unsigned char*)&mine, sizeof mine);
  comm_allgather(h, hs.p, hr.p, sizeof mine);
  std::vector<unsigned char> all = hr.to_host(sizeof mine * (size_t)cm.nranks);
  ok = all_agree(h, ok);
  A.peer.assign(cm.nranks, nullptr);
  if (ok) {
    for (int r = 0; r < cm.nranks && ok; ++r) {
      if (r == cm.rank) continue;
      cudaIpcMemHandle_t hd; memcpy(&hd, all.data() + sizeof hd * (size_t)r, sizeof hd);
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      A.peer[r] = (unsigned char*)q;
    }
    ok = all_agree(h, ok);
  }
  A.failed = !ok;
  if (A.failed && cm.rank == 0) fprintf(stderr, "pnpb200: peer-memory halo exchange unavailable (CUDA IPC), using NCCL send / recv\n");
}

// collective, once per halo (level 0: once per partition; coarse levels: once per hierarchy): tell every neighbour where
// its rows land in this rank's slot and which flag word it owns there, learn the same from it, agree on the path
void ipc_prepare(Handle& h, Halo& c)
{
  Comm& cm = h.comm; cudaStream_t s = h.stream;
  const int nn = (int)c.nbr.size();
  // a rank without neighbours on this level (nn == 0, empty lists) still takes part in the collective decisions below
  const int ghost0 = c.recv_ptr.empty() ? 0 : c.recv_ptr.front(), nghost = c.recv_ptr.empty() ? 0 : c.recv_ptr.back() - ghost0;
  if (!cm.ipc) ipc_arena_init(h, c.ipc_slot == 0 ? nghost : 0);
  IpcArena& A = *cm.ipc;
  bool ok = !A.failed && (nn == 0 || (c.ipc_slot >= 0 && c.ipc_slot < IPC_NSLOTS && nn <= IPC_MAX_NBR && nghost <= A.cap));
  std::vector<int> out(3 * (size_t)(nn ? nn : 1)), in(3 * (size_t)(nn ? nn : 1), 0);
  for (int k = 0; k < nn; ++k) { out[3 * k] = c.recv_ptr[k] - ghost0; out[3 * k + 1] = k; out[3 * k + 2] = ok ? 1 : 0; }
  DevBuf<int> ds, dr; ds.alloc(out.size(), s); dr.alloc(in.size(), s);
  ds.upload(out.data(), out.size());
  Api& a = api();
  nccl_check(a.gstart(), "groupStart");
  for (int k = 0; k < nn; ++k) {
    nccl_check(a.send(ds.p + 3 * k, 3, NCCL_INT32, c.nbr[k], cm.nccl, s), "send(handshake)");
    nccl_check(a.recv(dr.p + 3 * k, 3, NCCL_INT32, c.nbr[k], cm.nccl, s), "recv(handshake)");
  }
  nccl_check(a.gend(), "groupEnd");
  dr.download(in.data(), in.size());
  for (int k = 0; k < nn; ++k) ok = ok && in[3 * k + 2] == 1 && in[3 * k + 1] >= 0 && in[3 * k + 1] < IPC_MAX_NBR;
  ok = all_agree(h, ok);
  if (!ok) { c.ipc_state = -1; return; }
  if (nn == 0) { c.ipc_state = 2; return; }     // peer path agreed; nothing to exchange here, but the self-check is collective
  HaloIpcArgs& g = c.ipc;
  g = HaloIpcArgs();
  g.nnbr = nn; g.cap = A.cap; g.ghost0 = ghost0;
  unsigned char* slot = A.base + (size_t)c.ipc_slot * A.slot_bytes;
  g.ctl = (unsigned long long*)slot; g.recv = (double*)(slot + IPC_CTL_BYTES);
  g.send_beg[0] = 0; g.blk_beg[0] = 0;
  for (int k = 0; k < nn; ++k) {
    unsigned char* ps = A.peer[c.nbr[k]] + (size_t)c.ipc_slot * A.slot_bytes;
    g.peer_recv[k] = (double*)(ps + IPC_CTL_BYTES);
    g.peer_flag[k] = (unsigned long long*)ps + in[3 * k + 1];
    g.peer_goff[k] = in[3 * k];
    g.send_beg[k + 1] = c.send_ptr[k + 1];
    g.recv_off[k] = c.recv_ptr[k] - ghost0; g.recv_cnt[k] = c.recv_ptr[k + 1] - c.recv_ptr[k];
  }
  // CTAs per neighbour: one per 256 rows (the pack is a dependent gather: index, then row — it needs many threads in
  // flight, not many iterations per thread), at most 592 in total = 4 per SM, far below what is resident at once
  int per_cta = 256;
  for (;;) {
    int total = 0;
    for (int k = 0; k < nn; ++k) total += std::max(1, (std::max(c.send_ptr[k + 1] - c.send_ptr[k], g.recv_cnt[k]) + per_cta - 1) / per_cta);
    if (total <= 592) break;
    per_cta *= 2;
  }
  for (int k = 0; k < nn; ++k) {
    const int rows = std::max(c.send_ptr[k + 1] - c.send_ptr[k], g.recv_cnt[k]);
    g.blk_beg[k + 1] = g.blk_beg[k] + std::max(1, (rows + per_cta - 1) / per_cta);
  }
  c.ipc_state = 1;
}

void halo_exchange_ipc(Handle& h, Halo& c, double* vec, bool sweep_space)
{
  const int* idx = sweep_space ? c.send_pos.p : c.send_idx.p;
  const int grid = c.ipc.blk_beg[c.ipc.nnbr];
  if (sweep_space) halo_ipc_kernel<VS><<<grid, 256, 0, h.stream>>>(c.ipc, idx, vec);
  else halo_ipc_kernel<3><<<grid, 256, 0, h.stream>>>(c.ipc, idx, vec);
  PNP_LAUNCH_CHECK(); PNP_COUNT_LAUNCH();
}

// the level's first exchanges run through BOTH paths on a marker vector; any difference in the ghost rows (or a
// timed-out wait) sends the level back to NCCL on every rank
void ipc_selfcheck(Handle& h, Halo& c, bool sweep_space)
{
  cudaStream_t s = h.stream;
  if (!c.active()) { if (!all_agree(h, true)) c.ipc_state = -1; return; }
  const int n = c.recv_ptr.back(), n_own = c.recv_ptr.front();
  const int xs = sweep_space ? VS : 3;
  DevBuf<double> t1, t2; t1.alloc((size_t)xs * n, s); t2.alloc((size_t)xs * n, s);
  DevBuf<int> bad; bad.alloc(1, s); bad.zero();
  ipc_fill_kernel<<<cdiv((size_t)xs * n, 256), 256, 0, s>>>(n, n_own, h.comm.rank, xs, t1.p); PNP_COUNT_LAUNCH();
  PNP_CUDA(cudaMemcpyAsync(t2.p, t1.p, (size_t)xs * n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  halo_exchange_ipc(h, c, t1.p, sweep_space);
  halo_exchange_nccl(h, c, t2.p, sweep_space);
  const size_t cnt = (size_t)xs * (size_t)(n - n_own);
  if (cnt) { ipc_compare_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(cnt, t1.p + (size_t)xs * n_own, t2.p + (size_t)xs * n_own, bad.p); PNP_COUNT_LAUNCH(); }
  int hb = 0; bad.download(&hb, 1);
  unsigned long long err = 0;
  PNP_CUDA(cudaMemcpy(&err, c.ipc.ctl + 34, sizeof err, cudaMemcpyDeviceToHost));
  const bool ok = all_agree(h, hb == 0 && err == 0);
  if (!ok) {
    if (h.comm.rank == 0) fprintf(stderr, "pnpb200: peer-memory halo exchange of level slot %d disagrees with NCCL, falling back\n", c.ipc_slot);
    c.ipc_state = -1;
    if (err) { err = 0; PNP_CUDA(cudaMemcpy(c.ipc.ctl + 34, &err, sizeof err, cudaMemcpyHostToDevice)); }
  }
}

} // namespace

Comm::~Comm()
{
  delete ipc; ipc = nullptr;
  if (nccl) { try { api().destroy(nccl); } catch (...) {} }
}

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
  if (!h.comm.active()) return;
  if (c.active() && sweep_space && c.send_ptr.back() > 0 && !c.send_pos.p) throw Error(ERROR_UNKNOWN, "halo_exchange: level has no sweep-space send list");
  // the choice of path is a collective decision (all_agree): a rank that has no neighbour on this level goes through it
  // with the others and only then returns
  if (c.ipc_state == 0) {
    // first exchange with these lists (never inside a graph capture: the first visit of a level is direct)
    if (ipc_enabled()) ipc_prepare(h, c); else c.ipc_state = -1;
  }
  if ((c.ipc_state == 1 || c.ipc_state == 2) && h.comm.ipc && c.ipc_slot >= 0 && c.ipc_slot < IPC_NSLOTS && h.comm.ipc->checks[c.ipc_slot] < 2) {
    ++h.comm.ipc->checks[c.ipc_slot];
    ipc_selfcheck(h, c, sweep_space);
  }
  if (!c.active()) return;
  if (c.ipc_state == 1) { halo_exchange_ipc(h, c, vec, sweep_space); ++h.comm.n_ipc; }
  else { halo_exchange_nccl(h, c, vec, sweep_space); ++h.comm.n_nccl; }
}

void comm_check_ipc_error(Handle& h)
{
  if (!h.comm.ipc || !h.comm.ipc->base) return;
  for (int sl = 0; sl < IPC_NSLOTS; ++sl) {
    unsigned long long err = 0;
    PNP_CUDA(cudaMemcpy(&err, h.comm.ipc->base + (size_t)sl * h.comm.ipc->slot_bytes + 34 * sizeof(unsigned long long), sizeof err, cudaMemcpyDeviceToHost));
    if (err) throw Error(ERROR_DEVICE, "peer-memory halo exchange timed out waiting for a neighbour (level slot " + std::to_string(sl) + ")");
  }
}

namespace {
void halo_exchange_nccl(Handle& h, Halo& c, double* vec, bool sweep_space)
{
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
} // namespace

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

int pnpb200_comm_halo_counts(const pnpb200_handle* ph, long long out[2])
{
  if (!ph || !out) return ERROR_INPUT_PAR;
  out[0] = ph->h.comm.n_ipc; out[1] = ph->h.comm.n_nccl;
  return FASP_SUCCESS;
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
    c.ipc_slot = 0; c.ipc_state = 0;          // new lists: the peer-memory path is prepared again at the first exchange
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
