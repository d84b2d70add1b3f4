// pnp-b200 — internal host-side declarations shared by the .cu translation units.
#pragma once
#include "common.cuh"
#include "pnpb200/pnpb200.h"
#include <memory>

namespace pnpb200 {

// BSR(3) matrix on the device in "BSR-T/LDU" layout. Rows are stored in Gauss-Seidel sweep order
// (colour by colour). Inside a row the blocks are ordered [L | D | U]: first the nl blocks whose
// columns are swept BEFORE the row (lower part, columns ascending), then the diagonal block, then
// the blocks whose columns are swept after it (upper part, columns ascending). Each of the two
// parts is stored component-major,
//     lower:  val[9*s + t*nl + k]                    k <  nl
//     D + U:  val[9*s + 9*nl + t*(len-nl) + (k-nl)]  k >= nl      (t = 3*r + c, s = rs[i])
// so that lane k of a row group loads component t of its block with unit stride AND the lower /
// upper part of a row is one contiguous piece of memory: a forward sweep from a zero initial
// guess streams only L, the residual after it only U, and A*x after a backward sweep is
// b + L*(x_new - x_old) (bsr_ops.cu). common.cuh: bsrt_off().
struct BsrT {
  int n = 0;        // block rows
  int nnzb = 0;     // blocks
  int maxlen = 0;   // longest row (blocks), set by color_and_layout
  DevBuf<int> ia;       // natural CSR offsets (row lengths, FASP import/export)
  DevBuf<int> rs;       // physical start of each row's blocks: rows are STORED colour by colour, so that a
                        // Gauss-Seidel colour sweep streams one contiguous piece of ja/val
  DevBuf<int> rowid;    // physical row position -> row id (= rows sorted by colour)
  DevBuf<int> pia;      // CSR offsets in PHYSICAL order: row rowid[p] owns blocks [pia[p], pia[p+1])
  DevBuf<int4> rd;      // row descriptor of row i: (physical start, length, nl, i) — one 16-byte load
  DevBuf<int4> prd;     // the descriptors in matrix (sweep) order: prd[p] = (start, length, nl, pos[rowid[p]])
  DevBuf<int> vrowid;   // VECTOR order of the solver: entry q of a solver vector belongs to row vrowid[q]. Rows are
                        // ordered by (tile of VEC_TILE consecutive row ids, colour): a colour launch then reads b and
                        // writes x in contiguous runs of ~VEC_TILE/ncolors rows, while a pass in vector order still
                        // gathers x from a mesh-local window (pure colour order costs the matvec 37 %: measured)
  DevBuf<int> pos;      // row id -> vector position (inverse of vrowid)
  DevBuf<int4> vrd;     // descriptors in vector order: vrd[q] = (start, length, nl, q) of row vrowid[q]
  DevBuf<int> ja;       // column (ROW ID) of each block, physical order ([L | D | U], ascending inside each part)
  DevBuf<int> jpos;     // column as SWEEP POSITION: what the solver kernels gather with
  DevBuf<int> brow;     // row of each block
  DevBuf<int> dslot;    // position (absolute block index) of the diagonal block of each row (= rs[i] + nl_i)
  DevBuf<int> tslot;    // block index of the transposed block (j,i), -1 if absent
  DevBuf<double> val;   // 9*nnzb
  // PNP blocks have no cation-anion coupling: entries (1,2) and (2,1) are zero in every block of every level. After the
  // AMG set-up the matrix passes of the solve read a 7-value copy in the same [L | D | U] component-major layout
  // (7 instead of 9 components per part): 60 instead of 76 bytes per block. Built by bsrt_pack7 when every block of the
  // level has the two zeros (a general BSR(3) matrix does not, and keeps the 9-value passes); dropped whenever val changes.
  DevBuf<double> val7;  // 7*nnzb
  bool packed7 = false;
  DevBuf<int> pack_flag;      // raised by the packing pass if a block is not of the PNP structure
  bool pack7_fresh = false;   // val7 was written during this set-up by the strength pass (amg.cu: condense_kernel<true>)
};

// halo lists of one level (block-row granularity): neighbour k gets the owned rows
// send_idx[send_ptr[k] .. send_ptr[k+1]) and fills the local ghost rows [recv_ptr[k], recv_ptr[k+1])
// (ghost rows are the last rows in row-id AND in sweep order)
// Peer-memory halo exchange (comm.cu): what ONE kernel launch needs to pack the boundary rows, store them into the
// neighbours' receive buffers over NVLink, publish an epoch flag, wait for the neighbours' flags and unpack. Passed to the
// kernel by value (captured into the coarse-level CUDA graph like any other argument).
constexpr int IPC_MAX_NBR = 32;
struct HaloIpcArgs {
  int nnbr = 0;
  double* peer_recv[IPC_MAX_NBR];                 // neighbour's receive buffer (mapped through CUDA IPC)
  unsigned long long* peer_flag[IPC_MAX_NBR];     // the flag slot the neighbour watches for this rank
  int peer_goff[IPC_MAX_NBR];                     // ghost-row offset of this rank's rows in the neighbour's receive buffer
  int send_beg[IPC_MAX_NBR + 1];                  // this rank's send list, per neighbour
  int recv_off[IPC_MAX_NBR], recv_cnt[IPC_MAX_NBR];    // ghost rows of neighbour k: offset behind the owned rows, count
  int blk_beg[IPC_MAX_NBR + 1];                   // CTAs [blk_beg[k], blk_beg[k+1]) serve neighbour k
  double* recv = nullptr; int cap = 0;            // this rank's receive buffer: 2 halves x cap rows x 4 doubles (cap is the same on every rank)
  int ghost0 = 0;                                 // first ghost row of the level (= owned rows)
  unsigned long long* ctl = nullptr;              // [0..31] incoming epoch flags, [32] epoch, [33] CTAs done, [34] error, [64..95] CTAs that have sent
};

struct Halo {
  std::vector<int> nbr, send_ptr, recv_ptr;
  DevBuf<int> send_idx;         // row ids
  DevBuf<int> send_pos;         // the same rows as sweep positions (solver vectors)
  DevBuf<double> sendbuf;
  DevBuf<int> isendbuf;
  int ipc_slot = -1;            // level slot of the persistent peer-memory buffers (Comm::ipc), -1 = none
  int ipc_state = 0;            // 0 = not prepared for these lists yet, 1 = ready, 2 = ready but no neighbour here, -1 = NCCL send / recv
  HaloIpcArgs ipc;
  bool active() const { return !nbr.empty(); }
};

struct Level {
  BsrT A;
  int n_own = -1;               // owned block rows (first n_own local rows); -1 = all rows (single GPU).
                                // Ghost rows are identity rows of A; their vector entries come from the owner.
  Halo halo;
  DevBuf<double> dinv;          // 9*n, AoS: inverse of the diagonal blocks in SWEEP order (entry p = row A.rowid[p])
  // multicolour ordering
  int ncolors = 0;              // colours of the graph
  int ngroups = 0;              // (tile, colour) groups of one sweep, in order
  std::vector<int> color_ptr;   // host, ngroups+1 offsets into A.rowid
  DevBuf<int> d_color_ptr;      // the same offsets on the device (single-CTA sweep of small levels)
  DevBuf<int> preset_color;     // optional: a proper colouring known in closed form (box meshes, mesh_setup.cu)
  int preset_ncolors = 0;       // 0 = colour the graph with Jones-Plassmann + iterated greedy
  // aggregation to the next level
  int nagg = 0;
  DevBuf<int> agg;              // n, -1 = isolated
  DevBuf<int> agg_ptr, agg_rows;   // members of each aggregate (ascending row ids)
  // grid transfer in sweep space (built when the hierarchy is complete): coarse position of a fine
  // position (-1: none), and the members of agg_rows as fine positions
  DevBuf<int> agg_pos, agg_rows_pos;
  bool coarse_identity = false;    // the next level is the coarsest (dense solve): its vectors stay in row-id order
  // numeric RAP map: coarse block s sums fine blocks rap_perm[rap_seg[s] .. rap_seg[s+1])
  DevBuf<int> rap_beg, rap_end, rap_perm;
  bool rap_rowwise = false;     // Galerkin product by the row-wise kernels (amg.cu: rapw_*), no sort map
  int rap_maxcols = 0;          // longest coarse row (selects the shared-memory accumulator size)
  // work vectors (3*n)
  DevBuf<double> x, b, w;
  // K-cycle scratch: 5 vectors, device scalars + partials, block counter
  DevBuf<double> kc, ksc;
  DevBuf<unsigned> kcnt;
};

struct SolverConfig {
  // Krylov
  double tol = 1e-6; int maxit = 100; int restart = 30; int print_level = 0;
  // AMG
  int max_levels = 20; int coarse_dof = 100; double strong_coupled = 0.08; int max_aggregation = 20;
  double tentative_smooth = 0.67; int presmooth = 1, postsmooth = 1;
  int use_amg = 1;
  int use_ilu = 0, ilu_lfil = 0;   // block ILU(k) as the Krylov preconditioner (ilu.cu) when use_amg == 0
  int cycle_type = 1;       // V_CYCLE; anything else = K-cycle (nonlinear AMLI)
  int kcycle_levels = 3;    // coarse levels 1..kcycle_levels get the Krylov acceleration (measured best at 256^3)
};
SolverConfig make_config(const itsolver_param* it, const AMG_param* amg);

struct Hierarchy {
  std::vector<std::unique_ptr<Level>> lv;   // lv[0] = fine level (assembled Jacobian), lv[1..] rebuilt by amg_setup
  // coarsest level dense inverse
  int nd = 0;                   // scalar size of the GLOBAL coarsest level
  DevBuf<double> dense_inv;     // nd*nd, row-major, replicated on every rank
  // distributed coarsest solve: my rows [c_row0, c_row0 + 3*c_own), gather layout of the rhs
  int c_own = 0, c_row0 = 0, c_maxown3 = 0, c_nranks = 1;
  DevBuf<int> c_off3, c_cnt3;
  DevBuf<double> c_bpad, c_ball, c_bglob;
  bool symbolic_valid = false;
  bool coarse_dense = true;     // coarsest level: explicit dense inverse (else Gauss-Seidel sweeps, amg.cu)
  bool coarse_rep = false;      // multi-GPU: the last level is gathered to every rank and continued by Handle::hier_rep
  int level_offset = 0;         // global index of lv[0] (> 0 for the replicated coarse hierarchy)
  DevBuf<double> c_xglob;       // replicated solve: the gathered level's solution, row-id order
  // CUDA graph of the coarse part of a cycle (everything below level COARSE_GRAPH_LEVEL, amg.cu):
  // the K-cycle visits those levels several times per Krylov iteration with ~200 launches of
  // microsecond kernels each; captured once per hierarchy, replayed with one cudaGraphLaunch
  cudaGraphExec_t coarse_exec = nullptr;
  int coarse_visits = 0;        // visits since the hierarchy was built (the first one runs uncaptured)
  long coarse_nodes = 0;        // kernels in the graph (launch counter)
  unsigned coarse_sig = 0;      // cycle parameters the graph was captured with
  void drop_coarse_graph() { if (coarse_exec) { cudaGraphExecDestroy(coarse_exec); coarse_exec = nullptr; } coarse_visits = 0; coarse_nodes = 0; }
  ~Hierarchy() { if (coarse_exec) cudaGraphExecDestroy(coarse_exec); }
};

struct Krylov {
  int n = 0, restart = 0;
  DevBuf<double> V, Z;          // (restart+1)*n and restart*n
  DevBuf<double> r, partial;    // scratch
  DevBuf<double> bs, xs;        // rhs and iterate in sweep order
  DevBuf<double> dots;          // device results
  double* h_dots = nullptr;     // pinned host mirror
};

// block ILU(k) factors of the fine-level matrix (ilu.cu); rows in row-id order, columns ascending
struct IluFactor {
  int n = 0, nnz = 0, maxlen = 0, lfil = -1, pattern_nnzb = -1, epoch = 0;
  bool symbolic_valid = false;
  DevBuf<int> ia, ja, diag, src;     // factor pattern, position of the diagonal, source block in A (-1 = fill)
  DevBuf<double> val, dinv, y, x;    // L\U blocks (row-major 3x3), inverse diagonal blocks, work vectors
  DevBuf<int> flag, ticket;          // done-flags (numeric / forward / backward) and row tickets of the sync-free sweeps
};

struct Stats {
  int krylov_its = 0;
  double ms_assemble = 0, ms_setup = 0, ms_krylov = 0, ms_residual = 0;
  double ms_spmv = 0, ms_vcycle = 0, ms_ortho = 0;
  long launches = 0;
};

// first member of Handle => destroyed last, after every DevBuf has queued its cudaFreeAsync
struct StreamOwner {
  cudaStream_t s = nullptr;
  ~StreamOwner() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};

// multi-GPU state: one NCCL communicator per handle
struct Comm {
  void* nccl = nullptr;
  int rank = 0, nranks = 1;
  bool suspended = false;       // inside the replicated coarse hierarchy every rank works alone
  bool active() const { return nccl != nullptr && nranks > 1 && !suspended; }
  long n_ipc = 0, n_nccl = 0;           // halo exchanges issued through the peer-memory kernel / through NCCL send-recv
  struct IpcArena* ipc = nullptr;       // peer-memory halo exchange: IPC-exported receive slots per level + the peers' mappings (comm.cu)
  ~Comm();
};

struct Handle {
  StreamOwner owner;
  int device = 0;
  cudaStream_t stream = nullptr;
  // mesh
  int nv = 0, nc = 0;
  int box_n[3] = {0, 0, 0};     // grid of a mesh generated by generate_box_mesh (0: mesh given by the caller)
  Comm comm;
  std::vector<unsigned char> bc_user;   // Dirichlet flags as set by the caller (ghost flags are added on top)
  // caller's dof numbering (pnpb200_set_dof_map): internal 3*v + k <-> dofmap[3*v + k]; scalars v <-> sdofmap[v]; empty = identity
  std::vector<int> dofmap, dofmap_inv, sdofmap;
  DevBuf<double> xyz;           // 3*nv
  DevBuf<int> cells;            // 4*nc
  DevBuf<int> v2c_ptr, v2c_idx; // vertex -> incident cells (ascending)
  // coefficients / state (dof order 3*v + k)
  DevBuf<double> eps, fixed, diff, val, reac, uu, du, rhs;
  DevBuf<double> uu_base;       // solution the last pnpb200_solve_update started from (damped Newton: trial = base + alpha du)
  bool have_update = false;
  DevBuf<unsigned char> bcflag; // 3*nv
  DevBuf<int> vowner;           // multi-GPU: owner rank of every local vertex (set by pnpb200_set_partition)
  double q_eafe[3] = {0, 0, 0}; // valency function at dofs 0..2 (linear_pnp.cpp:220-224)
  bool have_coef = false;
  bool rhs_current = false;     // h.rhs holds the residual of the CURRENT solution (reused as the next step's rhs)
  int use_eafe = 1;
  double poisson_scale = 1.0;   // diode forms: scale of the charge terms in the Poisson row
  int conc_cutoff = 0;          // diode forms: exp(w) -> (w > -50 ? exp(w) : 0) in the Jacobian's diffusion terms
  // fine level matrix + per-pattern tables
  DevBuf<double> omega;         // nnzb: EAFE edge weights  sum_T S^T_ij
  DevBuf<double> k00;           // nnzb: permittivity stiffness (solution independent)
  DevBuf<double> cellrec;       // scatter records of the assembly (assemble.cu): 4 doubles per (cell, vertex pair) for the
                                // Jacobian, 3 per (cell, vertex) for the residual, 68 per cell for the Galerkin-only Jacobian
  int share_scratch = -1;       // the records are dead once the matrix is assembled and the Krylov basis is dead once the
                                // update is known: when the device cannot hold both (320^3: 67 + 56 GB next to 90 GB of
                                // matrix, tables and hierarchy) the basis lives in the record buffer. -1 = not decided yet
                                // (assemble_system decides from the free device memory), 0 = own buffers, 1 = shared
  DevBuf<int> cvslot;           // 4*nc : residual record slot of (cell, local vertex)   = position in the vertex' cell list
  DevBuf<int> cpslot;           // 10*nc: Jacobian record index of (cell, local vertex pair)
  DevBuf<int2> pblk;            // nnzb : (first record, count) of the pair run a block sums (physical block index)
  DevBuf<double> eafe_tab;      // 8*nv : per (vertex, ion) operands of the EAFE edge weights (assemble.cu)
  size_t n_pair_records = 0;
  DevBuf<double> eta;           // DG0 entropy indicator of the last pnpb200_entropy_indicator call (nc)
  bool eta_valid = false;
  // solver
  Hierarchy hier;
  Hierarchy hier_rep;           // multi-GPU: the small coarse levels, replicated on every rank (amg.cu: setup_replicated)
  Hierarchy* cur = &hier;       // the hierarchy the cycle / setup helpers of amg.cu are working on
  Krylov kry;
  IluFactor ilu;
  int precond = 0, precond_lfil = 2;   // pnpb200_set_preconditioner: 0 = AMG (bsr.dat), 1 = block ILU(k) (pnp_diode/bsr.dat)
  int reuse_hierarchy = 0;
  int sweep_shortcuts = getenv("PNPB200_NO_SWEEP_SHORTCUTS") ? 0 : 1;   // L-only / U-only / L*delta passes in the cycle
  int timing_detail = 0;
  SolverConfig last_cfg;
  Stats stats;
  Workspace ws;                 // AMG setup temporaries
  // scratch for reductions
  DevBuf<double> red;
  double* h_red = nullptr;      // pinned
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  Handle() {}
  Handle(const Handle&) = delete;
  ~Handle()
  {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (h_red) cudaFreeHost(h_red);
    for (int i = 0; i < 8; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
  }
};

// ---- mesh_setup.cu
void build_mesh_tables(Handle& h);                 // v2c, pattern, brow/dslot/tslot, omega
void compute_k00(Handle& h);
void build_assembly_tables(Handle& h);                // slot tables of the atomic-free scatter for the current fine-level layout
void relayout_fine_level(Handle& h);                // after the partition (ghost rows) is known
void set_fine_coloring(Handle& h, int ncolors, const int* host_color);   // caller-supplied proper colouring
void generate_box_mesh(Handle& h, int nx, int ny, int nz, double lx, double ly, double lz);   // domain_build
// colour the graph given in natural CSR (ia, ja_nat), lay the rows out colour by colour (rs, rowid,
// physical ja) and build brow/dslot/tslot
void color_and_layout(Level& L, const int* ja_nat, Workspace* ws, cudaStream_t s);
void exclusive_scan_int(const int* in, int* out, size_t n, cudaStream_t s);
// ---- assemble.cu
void assemble_system(Handle& h);                   // J (BSR-T) and rhs
void assemble_residual(Handle& h, DevBuf<double>& out, double* l2, double* linf);
void error_norms(Handle& h, const double* d_exact, double* l2, double* h1);   // src/error.cpp:58-103
void total_charge(Handle& h, double* d_out);                                  // linear_pnp.cpp:322-351
// min / max over the first n entries of a, of b and of a + b (out[6] on the host; all ranks' owned dofs)
void update_ranges(Handle& h, const double* a, const double* b, size_t n, double* out6);
void set_trial(Handle& h, double alpha);                                      // uu = uu_base + alpha * du
void entropy_indicator(Handle& h, double* d_eta);                             // src/mesh_refiner.cpp:212-271
long mark_cells(Handle& h, const double* d_eta, double tol, long target_size, unsigned char* d_marker, double* tol_used);
// ---- bsr_ops.cu
// Two index spaces. The assembly, the AMG setup and the C ABI address vectors by ROW ID (vertex
// order on the fine level). The solver (cycle, Krylov) keeps every level's vectors in SWEEP order
// (entry q belongs to row A.vrowid[q], see BsrT): a Gauss-Seidel colour then reads b and writes x
// in long contiguous runs. Same kernels, different index arrays: (rd, ja) for row-id space,
// (vrd or prd, jpos) for sweep space.
// Sweep-space vectors hold VS doubles per row (common.cuh: padded to one 32-byte sector so that the
// gather of x is one 256-bit load). Exception: the Krylov basis V of FGMRES stays PACKED (3 per row:
// every Arnoldi step streams the whole basis twice), so the passes that read the cycle's rhs (b) or
// write A*z (y) take the stride of those vectors as bs / ys.
void bsrt_spmv(const BsrT& A, const double* x, double* y, cudaStream_t s, int ys = VS);                  // y = A x      (sweep space)
void bsrt_residual(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s, int bs = VS, int ys = VS); // r = b - A x  (sweep space)
void bsrt_spmv_rowid(const BsrT& A, const double* x, double* y, cudaStream_t s);            // y = A x      (row-id space)
void bsrt_residual_rowid(const BsrT& A, const double* b, const double* x, double* r, cudaStream_t s);
void vec_to_sweep(const BsrT& A, const double* in_rowid, double* out_sweep, cudaStream_t s, int vs = VS);
void vec_from_sweep(const BsrT& A, const double* in_sweep, double* out_rowid, cudaStream_t s, int vs = VS);
void bsrt_diaginv(const BsrT& A, double* dinv, cudaStream_t s);
void bsrt_from_fasp(const double* d_val_fasp, BsrT& A, cudaStream_t s);
void bsrt_export_ja(const BsrT& A, int* d_ja_nat, cudaStream_t s);
void bsrt_to_fasp(const BsrT& A, double* d_val_fasp, cudaStream_t s);
void mcgs_sweep(const Level& L, const double* b, double* x, bool reverse, cudaStream_t s, int bs = VS);
// forward sweep from a ZERO initial guess (x need not be initialised on the swept rows): lower parts only
void mcgs_sweep_from_zero(const Level& L, const double* b, double* x, cudaStream_t s, int bs = VS);
// r = b - A x right after mcgs_sweep_from_zero (swept rows: r = -U x; ghost rows: full row)
void bsrt_residual_after_sweep(const Level& L, const double* b, const double* x, double* r, cudaStream_t s, int bs = VS);
// backward sweep that also returns delta = x_new - x_old, and y = A x_new = b + L delta after it
void mcgs_sweep_back_delta(const Level& L, const double* b, double* x, double* delta, cudaStream_t s, int bs = VS);
void bsrt_apply_after_sweep(const Level& L, const double* b, const double* delta, double* y, cudaStream_t s, int bs = VS, int ys = VS);
// pack the levels [0, nlevels) that the row kernels sweep: sets A.packed7 where the block structure allows it
void bsrt_pack7(Handle& h, Hierarchy& H);
bool bsrt_pack7_wanted(const Handle& h, const BsrT& A);
// ---- gs_stream.cu: bulk-copy (TMA) streamed colour launch for levels with short rows; false = not applicable
int max_row_length(const BsrT& A, cudaStream_t s);
bool gs_stream_colour(const Level& L, int p0, int p1, const double* b, double* x, double* dl, int mode, int bs, cudaStream_t s);
bool gs_stream_pass(const BsrT& A, int which, int p1, const double* b, const double* x, double* y, int bs, int ys, cudaStream_t s);
// ---- vecops.cu
void vec_set(double* x, double v, size_t n, cudaStream_t s);
void vec_copy(double* dst, const double* src, size_t n, cudaStream_t s);
void vec_axpy(double a, const double* x, double* y, size_t n, cudaStream_t s);      // y += a x
void vec_scale(double a, double* x, size_t n, cudaStream_t s);
void vec_restride(const double* in, int is, double* out, int os, size_t nrows, cudaStream_t s);   // 3 <-> 4 doubles per row
double vec_norm2(Handle& h, const double* x, size_t n);                              // sync
// dots[i] = <V_i, w> for i < nv, dots[nv] = <w, w>; V_i = V + i*ld. Synchronises.
void multi_dot(Handle& h, const double* V, size_t ld, int nvec, const double* w, size_t n, double* out_host);
// w += sum_i coef[i] * V_i ; if norm2_out != null also returns ||w||^2 (sync)
void multi_axpy(Handle& h, const double* V, size_t ld, int nvec, const double* coef_host, double* w, size_t n,
                double* norm2_out);
// ---- comm.cu
// 3 doubles per block row; sweep_space: vec is a solver vector (indexed by sweep position), else by row id
void halo_exchange(Handle& h, Halo& halo, double* vec, bool sweep_space = false);
void halo_exchange_int(Handle& h, Halo& halo, int* vec);     // 1 int per block row
void comm_check_ipc_error(Handle& h);                        // throws if a peer-memory exchange timed out
void comm_allreduce(Handle& h, double* dev, int count, bool is_max);
void comm_allreduce_int(Handle& h, int* dev, int count, bool is_max);
void comm_allgather(Handle& h, const void* send, void* recv, size_t bytes_per_rank);
inline int owned_rows(const Level& L) { return L.n_own >= 0 ? L.n_own : L.A.n; }
inline size_t owned_dofs(const Handle& h) { return 3 * (size_t)owned_rows(*h.hier.lv[0]); }
// ---- block_solver.cu: PNP-Stokes 2x2 block system (benchmarks/pnp_stokes/linear_pnp_ns.cpp:137-235)
// scalar CSR on the device (the coupling blocks and the Stokes block; the PNP block and the
// velocity block additionally live as BSR(3) hierarchies in their own handles)
struct DevCsr {
  int nrow = 0, ncol = 0, nnz = 0;
  DevBuf<int> ia, ja;
  DevBuf<double> val;
};
void csr_upload(const dCSRmat* A, DevCsr& D, cudaStream_t s);
// y[i - r0] = beta * y[i - r0] + alpha * sum_{c0 <= j < c1} A_ij x[j - c0]   for rows r0 <= i < r1
void csr_spmv(const DevCsr& A, int r0, int r1, int c0, int c1, double alpha, const double* x, double beta, double* y,
              cudaStream_t s);
struct BlockStats { int outer_its = 0, pnp_solves = 0, pnp_its = 0, stokes_solves = 0, stokes_its = 0, vel_solves = 0, vel_its = 0; double relres = 0; };
struct BlockSystem {
  Handle* outer = nullptr;      // stream, reduction scratch and Krylov storage of the outer iteration
  Handle* pnp = nullptr;        // PNP block as BSR-T (3 dofs per vertex) + its AMG hierarchy
  Handle* vel = nullptr;        // velocity block as BSR-T (dofs grouped in threes, identity padding) + hierarchy
  DevCsr App, Aps, Asp, Ass;    // the four blocks, scalar CSR (App only for the outer matvec)
  int np = 0, ns = 0, nu = 0, npres = 0;
  DevBuf<double> sinv;          // npres: inverse of the diagonal of the SIMPLE Schur complement
  SolverConfig cfg_outer, cfg_pnp, cfg_ns, cfg_vel;
  bool vel_krylov = true;       // velocity block: inner AMG-FGMRES (ns.dat's _v parameters) instead of one AMG cycle
};
// outer FGMRES with the block upper-triangular preconditioner; b, x device vectors of length np + ns
int block_pnp_stokes_solve(BlockSystem& S, const double* b, double* x, BlockStats* st);

// ---- amg.cu
void amg_setup(Handle& h, const SolverConfig& cfg);
// z = M^{-1} r. If Az != nullptr and the cycle can deliver it for half a matrix pass (at least one
// post-smoothing sweep), Az = A z is written too and true is returned.
// r carries rs doubles per row, Az azs (the Krylov basis is packed); z is a padded sweep-space vector.
bool amg_vcycle(Handle& h, const SolverConfig& cfg, const double* r, double* z, double* Az = nullptr, int rs = VS, int azs = VS);
// ---- ilu.cu
void ilu_setup(Handle& h, int lfil);                                         // symbolic (cached per pattern) + numeric
void ilu_apply(Handle& h, const double* r, int rs, double* z, int zs, bool sweep_space);
// ---- fgmres.cu
// zero_guess: x is known to be all zeros on entry (linear_pnp.cpp:93), so r0 = b needs no matvec
int fgmres_solve(Handle& h, const SolverConfig& cfg, const double* b, double* x, bool zero_guess = false);

} // namespace pnpb200
