// pnp-b200 — per-mesh tables built on the device:
// vertex->cell adjacency, the BSR(3) sparsity pattern DOLFIN's assembler would build
// (src/pde.cpp:545: all vertex pairs that share a cell, columns ascending), row / diagonal /
// transpose slot maps, the EAFE edge table omega_ij = sum_T S^T_ij (include/EAFE.h:4825-4849,
// SURVEY App. A.3) and the solution-independent permittivity stiffness (forms.ufl:28).
#include "internal.hpp"
#include "fem.cuh"
#include <cub/cub.cuh>

namespace pnpb200 {

thread_local long g_launches = 0;
int g_debug_sync = getenv("PNPB200_DEBUG_SYNC") ? atoi(getenv("PNPB200_DEBUG_SYNC")) : 0;

namespace {

constexpr int MAX_ROW = 128;   // max blocks per row the pattern builder supports

__global__ void count_v2c_kernel(int nc, const int4* __restrict__ cells, int* __restrict__ deg)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int4 cv = cells[c];
  atomicAdd(&deg[cv.x], 1); atomicAdd(&deg[cv.y], 1); atomicAdd(&deg[cv.z], 1); atomicAdd(&deg[cv.w], 1);
}

__global__ void fill_v2c_kernel(int nc, const int4* __restrict__ cells, int* __restrict__ cursor,
                                int* __restrict__ idx)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int4 cv = cells[c];
  idx[atomicAdd(&cursor[cv.x], 1)] = c; idx[atomicAdd(&cursor[cv.y], 1)] = c;
  idx[atomicAdd(&cursor[cv.z], 1)] = c; idx[atomicAdd(&cursor[cv.w], 1)] = c;
}

// ascending order per vertex makes every later gather deterministic
__global__ void sort_v2c_kernel(int nv, const int* __restrict__ ptr, int* __restrict__ idx)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int s = ptr[v], e = ptr[v + 1];
  for (int a = s + 1; a < e; ++a) {
    const int key = idx[a];
    int b = a - 1;
    while (b >= s && idx[b] > key) { idx[b + 1] = idx[b]; --b; }
    idx[b + 1] = key;
  }
}

// sorted-unique neighbour set of one vertex; FILL=false counts, FILL=true writes JA
template <bool FILL>
__global__ void pattern_kernel(int nv, const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                               const int4* __restrict__ cells, int* __restrict__ rowlen,
                               const int* __restrict__ ia, int* __restrict__ ja, int* __restrict__ err)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  int nb[MAX_ROW];
  int cnt = 0;
  for (int p = v2c_ptr[v]; p < v2c_ptr[v + 1]; ++p) {
    const int4 cv = cells[v2c_idx[p]];
    const int w4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int w = w4[r];
      int lo = 0;
      while (lo < cnt && nb[lo] < w) ++lo;
      if (lo < cnt && nb[lo] == w) continue;
      if (cnt >= MAX_ROW) { *err = 1; continue; }
      for (int t = cnt; t > lo; --t) nb[t] = nb[t - 1];
      nb[lo] = w; ++cnt;
    }
  }
  if (cnt == 0) { nb[0] = v; cnt = 1; }     // vertex without cells keeps a diagonal
  if (!FILL) rowlen[v] = cnt;
  else { const int s = ia[v]; for (int t = 0; t < cnt; ++t) ja[s + t] = nb[t]; }
}

// omega_b = sum over cells containing (i,j) of the P1 stiffness entry (det/6) grad l_i . grad l_j
// k00_b   = sum over cells of (sum_ip W24 eps(ip)) * det grad l_i . grad l_j
template <bool WITH_EPS>
__global__ void edge_table_kernel(int nnzb, const int* __restrict__ ja, const int* __restrict__ brow,
                                  const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                                  const int4* __restrict__ cells, const double* __restrict__ xyz,
                                  const double* __restrict__ eps, double* __restrict__ out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int i = brow[b], j = ja[b];
  double acc = 0.0;
  for (int p = v2c_ptr[i]; p < v2c_ptr[i + 1]; ++p) {
    const int4 cv = cells[v2c_idx[p]];
    const int lj = local_index(cv, j);
    if (lj < 0) continue;
    const int li = local_index(cv, i);
    double x[4][3], g[4][3], det;
    load_cell_xyz(xyz, cv, x);
    tet_geom(x, g, det);
    const double G = det * dot3(g[li], g[lj]);
    if (!WITH_EPS) acc += 0.166666666666667 * G;
    else {
      const double e4[4] = {eps[cv.x], eps[cv.y], eps[cv.z], eps[cv.w]};
      double we = 0.0;
#pragma unroll 4
      for (int ip = 0; ip < 24; ++ip) {
        const double e_ip = c_P24[ip][0] * e4[0] + c_P24[ip][1] * e4[1] + c_P24[ip][2] * e4[2] + c_P24[ip][3] * e4[3];
        we += c_W24[ip] * e_ip;
      }
      acc += we * G;
    }
  }
  out[b] = acc;
}

// ---- slot tables of the atomic-free scatter (assemble.cu) ----------------------------------------
// Residual: the 3 numbers cell c contributes to its r-th vertex v go to slot p = position of c in v's
// ascending cell list (v2c), so that the per-vertex sum reads 3 * |cells(v)| contiguous doubles.
__global__ void cell_vertex_slot_kernel(int nv, const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                                        const int4* __restrict__ cells, int* __restrict__ cvslot)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  for (int p = v2c_ptr[v]; p < v2c_ptr[v + 1]; ++p) {
    const int c = v2c_idx[p];
    const int r = local_index(cells[c], v);
    if (r >= 0) cvslot[4 * (size_t)c + r] = p;
  }
}

// Jacobian: a vertex pair {i, j} (i <= j, diagonal included) owns one run of records, one per cell that
// contains both vertices, ascending cell id. The run belongs to the UPPER block (i, j); the block (j, i)
// reads the same run. Pass 1 counts the cells of every upper block, pass 2 hands every (cell, local pair)
// its record index and every upper block its (start, count).
__global__ void pair_count_kernel(int nnzb, const int* __restrict__ ja, const int* __restrict__ brow,
                                  const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                                  const int4* __restrict__ cells, int* __restrict__ cnt)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nnzb) return;
  if (b == nnzb) { cnt[b] = 0; return; }
  const int i = brow[b], j = ja[b];
  int n = 0;
  if (j >= i)
    for (int p = v2c_ptr[i]; p < v2c_ptr[i + 1]; ++p) n += local_index(cells[v2c_idx[p]], j) >= 0 ? 1 : 0;
  cnt[b] = n;
}

__global__ void pair_fill_kernel(int nnzb, const int* __restrict__ ja, const int* __restrict__ brow,
                                 const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                                 const int4* __restrict__ cells, const int* __restrict__ start,
                                 int* __restrict__ cpslot, int2* __restrict__ pblk)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int i = brow[b], j = ja[b];
  if (j < i) return;
  int q = start[b];
  for (int p = v2c_ptr[i]; p < v2c_ptr[i + 1]; ++p) {
    const int c = v2c_idx[p];
    const int4 cv = cells[c];
    const int lj = local_index(cv, j);
    if (lj < 0) continue;
    cpslot[10 * (size_t)c + sym4(local_index(cv, i), lj)] = q++;
  }
  pblk[b] = make_int2(start[b], q - start[b]);
}

__global__ void pair_lower_kernel(int nnzb, const int* __restrict__ ja, const int* __restrict__ brow,
                                  const int* __restrict__ tslot, int2* __restrict__ pblk)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb || ja[b] >= brow[b]) return;
  const int t = tslot[b];
  pblk[b] = t >= 0 ? pblk[t] : make_int2(0, 0);        // the pattern of a mesh is symmetric: t >= 0
}

} // namespace

// cvslot / cpslot / pblk for the current layout of the fine level (pblk is indexed by PHYSICAL block)
void build_assembly_tables(Handle& h)
{
  cudaStream_t s = h.stream;
  const BsrT& A = h.hier.lv[0]->A;
  const int4* cells = reinterpret_cast<const int4*>(h.cells.p);
  const int nnzb = A.nnzb;
  if (10 * (size_t)h.nc > 0x7fffffffull) throw Error(ERROR_INPUT_PAR, "mesh too large: more than 2^31 (cell, vertex pair) records");
  h.cvslot.alloc(4 * (size_t)h.nc, s);
  cell_vertex_slot_kernel<<<cdiv(h.nv, 128), 128, 0, s>>>(h.nv, h.v2c_ptr.p, h.v2c_idx.p, cells, h.cvslot.p); PNP_COUNT_LAUNCH();
  h.cpslot.alloc(10 * (size_t)h.nc, s);
  h.pblk.alloc(nnzb, s);
  DevBuf<int> cnt, start; cnt.alloc((size_t)nnzb + 1, s); start.alloc((size_t)nnzb + 1, s);
  pair_count_kernel<<<cdiv((size_t)nnzb + 1, 128), 128, 0, s>>>(nnzb, A.ja.p, A.brow.p, h.v2c_ptr.p, h.v2c_idx.p, cells, cnt.p); PNP_COUNT_LAUNCH();
  exclusive_scan_int(cnt.p, start.p, (size_t)nnzb + 1, s);
  pair_fill_kernel<<<cdiv(nnzb, 128), 128, 0, s>>>(nnzb, A.ja.p, A.brow.p, h.v2c_ptr.p, h.v2c_idx.p, cells, start.p, h.cpslot.p, h.pblk.p); PNP_COUNT_LAUNCH();
  pair_lower_kernel<<<cdiv(nnzb, 256), 256, 0, s>>>(nnzb, A.ja.p, A.brow.p, A.tslot.p, h.pblk.p); PNP_COUNT_LAUNCH();
  int total = 0;
  PNP_CUDA(cudaMemcpyAsync(&total, start.p + nnzb, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if ((size_t)total != 10 * (size_t)h.nc) throw Error(ERROR_DATA_STRUCTURE, "assembly tables: a vertex pair of a cell has no block in the pattern");
  h.n_pair_records = (size_t)total;
}

void exclusive_scan_int(const int* in, int* out, size_t n, cudaStream_t s)
{
  size_t bytes = 0;
  PNP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s));
  DevBuf<unsigned char> tmp; tmp.alloc(bytes, s);
  PNP_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, s));
  PNP_COUNT_LAUNCH();
}


// ---- closed-form colouring of dolfin::BoxMesh ---------------------------------------------------
// The Kuhn split (6 tetrahedra per cube around the diagonal v0-v7) couples vertex (i, j, k) with
// +-(1,0,0), +-(0,1,0), +-(0,0,1), +-(1,1,0), +-(1,0,1), +-(0,1,1), +-(1,1,1): on every one of
// these i + j + k changes by 1, 2 or 3, so (i + j + k) mod 4 is a proper colouring with FOUR
// balanced classes (the triangulation is balanced: the 4 vertices of every cell get the 4 colours).
// Jones-Plassmann + iterated greedy finds 10 unbalanced ones on the same graph. Fewer colours =
// fewer launches per Gauss-Seidel sweep and, since every colour launch re-reads most of the
// gathered x (profiles/r01_traffic_gsd_n256.json), less DRAM traffic per sweep. The colouring is
// verified against the actual pattern before it is used; every BASELINE config runs on a BoxMesh
// (src/domain.cpp:301-308 rejects anything else). Measured at 256^3: 758 -> 747 ms per step, 26 Krylov
// iterations with either colouring. PNPB200_BOX_COLORING=0 keeps the graph colouring.
namespace {
__global__ void box_color_kernel(int nx, int ny, int nv, int* __restrict__ color)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int i = v % (nx + 1), j = (v / (nx + 1)) % (ny + 1), k = v / ((nx + 1) * (ny + 1));
  color[v] = (i + j + k) & 3;
}
__global__ void check_coloring_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                      const int* __restrict__ color, int* __restrict__ bad)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = color[i];
  for (int p = ia[i]; p < ia[i + 1]; ++p) { const int j = ja[p]; if (j != i && color[j] == c) { *bad = 1; return; } }
}
#ifndef PNPB200_BOX_COLORING_DEFAULT
#define PNPB200_BOX_COLORING_DEFAULT 1
#endif
void box_coloring(Handle& h, Level& L, const int* ja_nat)
{
  L.preset_ncolors = 0;
  static const int on = getenv("PNPB200_BOX_COLORING") ? atoi(getenv("PNPB200_BOX_COLORING")) : PNPB200_BOX_COLORING_DEFAULT;
  if (!on || h.box_n[0] <= 0) return;
  cudaStream_t s = h.stream;
  const int nv = h.nv;
  L.preset_color.alloc(nv, s);
  box_color_kernel<<<cdiv(nv, 256), 256, 0, s>>>(h.box_n[0], h.box_n[1], nv, L.preset_color.p); PNP_COUNT_LAUNCH();
  DevBuf<int> bad; bad.alloc(1, s); bad.zero();
  check_coloring_kernel<<<cdiv(nv, 256), 256, 0, s>>>(nv, L.A.ia.p, ja_nat, L.preset_color.p, bad.p); PNP_COUNT_LAUNCH();
  int hb = 0; bad.download(&hb, 1);
  if (hb == 0) L.preset_ncolors = 4;      // else: not the Kuhn pattern after all, colour the graph
}
} // namespace

void build_mesh_tables(Handle& h)
{
  cudaStream_t s = h.stream;
  const int nv = h.nv, nc = h.nc;
  const int4* cells = reinterpret_cast<const int4*>(h.cells.p);
  // vertex -> cells
  DevBuf<int> deg; deg.alloc(nv + 1, s); deg.zero();
  count_v2c_kernel<<<cdiv(nc, 256), 256, 0, s>>>(nc, cells, deg.p); PNP_COUNT_LAUNCH();
  h.v2c_ptr.alloc(nv + 1, s);
  exclusive_scan_int(deg.p, h.v2c_ptr.p, nv + 1, s);
  h.v2c_idx.alloc(4 * (size_t)nc, s);
  PNP_CUDA(cudaMemcpyAsync(deg.p, h.v2c_ptr.p, (nv + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s));
  fill_v2c_kernel<<<cdiv(nc, 256), 256, 0, s>>>(nc, cells, deg.p, h.v2c_idx.p); PNP_COUNT_LAUNCH();
  sort_v2c_kernel<<<cdiv(nv, 128), 128, 0, s>>>(nv, h.v2c_ptr.p, h.v2c_idx.p); PNP_COUNT_LAUNCH();
  // pattern
  Level& L = *h.hier.lv[0];
  BsrT& A = L.A;
  A.n = nv;
  DevBuf<int> err; err.alloc(1, s); err.zero();
  DevBuf<int> rowlen; rowlen.alloc(nv + 1, s); rowlen.zero();
  pattern_kernel<false><<<cdiv(nv, 64), 64, 0, s>>>(nv, h.v2c_ptr.p, h.v2c_idx.p, cells, rowlen.p, nullptr, nullptr, err.p);
  PNP_COUNT_LAUNCH();
  A.ia.alloc(nv + 1, s);
  exclusive_scan_int(rowlen.p, A.ia.p, nv + 1, s);
  int nnzb = 0, herr = 0;
  PNP_CUDA(cudaMemcpyAsync(&nnzb, A.ia.p + nv, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  PNP_CUDA(cudaStreamSynchronize(s));
  if (herr) throw Error(ERROR_DATA_STRUCTURE, "mesh vertex with more than 128 neighbours is not supported");
  A.nnzb = nnzb;
  {
    DevBuf<int> ja_nat; ja_nat.alloc(nnzb, s);
    pattern_kernel<true><<<cdiv(nv, 64), 64, 0, s>>>(nv, h.v2c_ptr.p, h.v2c_idx.p, cells, nullptr, A.ia.p, ja_nat.p, err.p);
    PNP_COUNT_LAUNCH();
    box_coloring(h, L, ja_nat.p);
    color_and_layout(L, ja_nat.p, nullptr, s);      // fine-level colouring + colour-major row storage
  }
  A.val.alloc(9 * (size_t)nnzb, s);
  // EAFE edge table
  h.omega.alloc(nnzb, s);
  edge_table_kernel<false><<<cdiv(nnzb, 128), 128, 0, s>>>(nnzb, A.ja.p, A.brow.p, h.v2c_ptr.p, h.v2c_idx.p, cells,
                                                          h.xyz.p, nullptr, h.omega.p);
  PNP_COUNT_LAUNCH();
  build_assembly_tables(h);
  PNP_LAUNCH_CHECK();
}

// caller-supplied colouring of the fine level (pnpb200_set_coloring): verified against the pattern,
// then the level is laid out again colour by colour
void set_fine_coloring(Handle& h, int ncolors, const int* host_color)
{
  cudaStream_t s = h.stream;
  Level& L = *h.hier.lv[0];
  BsrT& A = L.A;
  if (ncolors < 1 || ncolors > 62) throw Error(ERROR_INPUT_PAR, "pnpb200_set_coloring: 1 .. 62 colours");
  for (int v = 0; v < A.n; ++v) if (host_color[v] < 0 || host_color[v] >= ncolors) throw Error(ERROR_INPUT_PAR, "pnpb200_set_coloring: colour out of range");
  L.preset_color.alloc(A.n, s);
  L.preset_color.upload(host_color, A.n);
  DevBuf<int> ja_nat; ja_nat.alloc(A.nnzb, s);
  bsrt_export_ja(A, ja_nat.p, s);
  DevBuf<int> bad; bad.alloc(1, s); bad.zero();
  check_coloring_kernel<<<cdiv(A.n, 256), 256, 0, s>>>(A.n, A.ia.p, ja_nat.p, L.preset_color.p, bad.p); PNP_COUNT_LAUNCH();
  int hb = 0; bad.download(&hb, 1);
  if (hb) { L.preset_ncolors = 0; throw Error(ERROR_INPUT_PAR, "pnpb200_set_coloring: two coupled vertices share a colour"); }
  L.preset_ncolors = ncolors;
  relayout_fine_level(h);
}

void relayout_fine_level(Handle& h)
{
  cudaStream_t s = h.stream;
  Level& L = *h.hier.lv[0];
  BsrT& A = L.A;
  DevBuf<int> ja_nat; ja_nat.alloc(A.nnzb, s);
  bsrt_export_ja(A, ja_nat.p, s);
  color_and_layout(L, ja_nat.p, nullptr, s);
  edge_table_kernel<false><<<cdiv(A.nnzb, 128), 128, 0, s>>>(A.nnzb, A.ja.p, A.brow.p, h.v2c_ptr.p, h.v2c_idx.p,
                                                            reinterpret_cast<const int4*>(h.cells.p), h.xyz.p, nullptr, h.omega.p);
  PNP_COUNT_LAUNCH();
  build_assembly_tables(h);
  if (h.have_coef) compute_k00(h);
  h.hier.symbolic_valid = false;
}

void compute_k00(Handle& h)
{
  cudaStream_t s = h.stream;
  BsrT& A = h.hier.lv[0]->A;
  h.k00.alloc(A.nnzb, s);
  edge_table_kernel<true><<<cdiv(A.nnzb, 128), 128, 0, s>>>(A.nnzb, A.ja.p, A.brow.p, h.v2c_ptr.p, h.v2c_idx.p,
                                                           reinterpret_cast<const int4*>(h.cells.p), h.xyz.p,
                                                           h.eps.p, h.k00.p);
  PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}

} // namespace pnpb200

// ---- domain_build (src/domain.cpp:290-311): centred dolfin::BoxMesh generated on the device ----
namespace pnpb200 {
namespace {
__global__ void box_vertices_kernel(int nx, int ny, int nz, double lx, double ly, double lz, double* __restrict__ xyz)
{
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvx = nx + 1, nvy = ny + 1, nv = nvx * nvy * (size_t)(nz + 1);
  if (v >= nv) return;
  const int ix = (int)(v % nvx), iy = (int)((v / nvx) % nvy), iz = (int)(v / (nvx * nvy));
  // DOLFIN BoxMesh [EXT]: x = a + ix*(b-a)/nx with a = -L/2, b = +L/2
  const double a = -lx / 2, b = lx / 2, c = -ly / 2, d = ly / 2, e = -lz / 2, f = lz / 2;
  xyz[3 * v]     = a + ((double)ix) * (b - a) / (double)nx;
  xyz[3 * v + 1] = c + ((double)iy) * (d - c) / (double)ny;
  xyz[3 * v + 2] = e + ((double)iz) * (f - e) / (double)nz;
}

__global__ void box_cells_kernel(int nx, int ny, int nz, int4* __restrict__ cells)
{
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t ncube = (size_t)nx * ny * nz;
  if (q >= ncube) return;
  const int ix = (int)(q % nx), iy = (int)((q / nx) % ny), iz = (int)(q / ((size_t)nx * ny));
  const int sx = nx + 1, sxy = (nx + 1) * (ny + 1);
  const int v0 = iz * sxy + iy * sx + ix, v1 = v0 + 1, v2 = v0 + sx, v3 = v1 + sx;
  const int v4 = v0 + sxy, v5 = v1 + sxy, v6 = v2 + sxy, v7 = v3 + sxy;
  // six tetrahedra around the v0-v7 diagonal, vertices ascending (UFC order)
  int4* o = cells + 6 * q;
  o[0] = make_int4(v0, v1, v3, v7); o[1] = make_int4(v0, v1, v5, v7); o[2] = make_int4(v0, v4, v5, v7);
  o[3] = make_int4(v0, v2, v3, v7); o[4] = make_int4(v0, v4, v6, v7); o[5] = make_int4(v0, v2, v6, v7);
}
} // namespace

void generate_box_mesh(Handle& h, int nx, int ny, int nz, double lx, double ly, double lz)
{
  cudaStream_t s = h.stream;
  const size_t nv = (size_t)(nx + 1) * (ny + 1) * (nz + 1), ncube = (size_t)nx * ny * nz;
  if (nv > 0x7fffffffull / 16 || 6 * ncube > 0x7fffffffull / 4) throw Error(ERROR_INPUT_PAR, "box mesh too large for 32-bit indices");
  h.nv = (int)nv; h.nc = (int)(6 * ncube);
  h.box_n[0] = nx; h.box_n[1] = ny; h.box_n[2] = nz;
  h.xyz.alloc(3 * nv, s); h.cells.alloc(4 * (size_t)h.nc, s);
  box_vertices_kernel<<<cdiv(nv, 256), 256, 0, s>>>(nx, ny, nz, lx, ly, lz, h.xyz.p); PNP_COUNT_LAUNCH();
  box_cells_kernel<<<cdiv(ncube, 256), 256, 0, s>>>(nx, ny, nz, reinterpret_cast<int4*>(h.cells.p)); PNP_COUNT_LAUNCH();
  PNP_LAUNCH_CHECK();
}
} // namespace pnpb200
