// TEST INFRASTRUCTURE ONLY. Compiles the product's row-assembly code (modular-pnp_b200/csrc/ns_forms.cuh, the very
// functions the CUDA kernels of assemble_ns.cu call) for the HOST, so that the CPU suite can check it against the oracle
// and the reference's compiled kernels without a GPU. Never linked into libpnpb200.so; the product has no CPU path.
#include "ns_forms.cuh"

extern "C" void shim_ns_assemble(int nv, int nc, int nf, const double* xyz, const int* cells, const int* cell_facets,
                                 const int* facet_cells, const int* vcell_ptr, const int* vcell, const double* cc,
                                 const double* uu, const double* pp, const double* par8, const int* ia, const int* ja,
                                 const unsigned char* bc, double* mom, double* val, double* rhs, int matrix)
{
  pnpb200::NsView S;
  S.nv = nv; S.nc = nc; S.nf = nf; S.ndof = 3 * nv + nf + nc;
  S.xyz = xyz; S.cells = cells; S.cell_facets = cell_facets; S.facet_cells = facet_cells;
  S.vcell_ptr = vcell_ptr; S.vcell = vcell; S.cc = cc; S.uu = uu; S.pp = pp;
  for (int i = 0; i < 8; ++i) S.par[i] = par8[i];
  S.mom = mom; S.ia = ia; S.ja = ja; S.val = val; S.rhs = rhs; S.bc = bc;
  for (int c = 0; c < nc; ++c) pnpb200::ns_cell_moments(S, c);
  for (int r = 0; r < S.ndof; ++r) pnpb200::ns_assemble_row(S, r, matrix != 0);
}
