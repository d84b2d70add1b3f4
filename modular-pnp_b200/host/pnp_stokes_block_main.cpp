// pnp-b200 host side — driver of the PNP-Stokes linear solve with the control flow of
// Linear_PNP_NS::fasp_solve (benchmarks/pnp_stokes/linear_pnp_ns.cpp:178-235) and the parameter
// set-up of benchmarks/pnp_stokes/main.cpp:55-86, written against host/linear_pnp_ns.hpp. The
// assembled mixed-space system comes from a file (DOLFIN's part): little-endian
//   int32 n, nnz, n_components(5) | int32 IA[n+1], JA[nnz] | f64 val[nnz] | f64 rhs[n]
//   | per component: int32 count, int32 dofs[count]      (components: c1, c2, phi, velocity, pressure)
// and the update is written back as f64[n] in the mixed numbering.
//   usage: pnp_stokes_block system.bin update.bin [bcsr.dat bsr.dat ns.dat]
#include <cstdio>
#include <string>
#include <vector>
#include <unistd.h>
#include "linear_pnp_ns.hpp"

template <typename T> static bool rd(FILE* f, T* p, size_t n) { return fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s system.bin update.bin [bcsr.dat bsr.dat ns.dat]\n", argv[0]); return 1; }
  char exe[4096] = {0};
  std::string here = ".";
  if (readlink("/proc/self/exe", exe, sizeof exe - 1) > 0) { here = exe; here = here.substr(0, here.find_last_of('/')); }
  const std::string bcsr = argc > 3 ? argv[3] : here + "/bcsr.dat", bsr = argc > 4 ? argv[4] : here + "/bsr.dat",
                    ns = argc > 5 ? argv[5] : here + "/ns.dat";
  // main.cpp:55-86
  input_param inpar; itsolver_param itpar; AMG_param amgpar; ILU_param ilupar;
  fasp_param_input(bcsr.c_str(), &inpar);
  fasp_param_init(&inpar, &itpar, &amgpar, &ilupar, NULL);
  input_param pnp_inpar; itsolver_param pnp_itpar; AMG_param pnp_amgpar; ILU_param pnp_ilupar; Schwarz_param pnp_schpar;
  fasp_param_input(bsr.c_str(), &pnp_inpar);
  fasp_param_init(&pnp_inpar, &pnp_itpar, &pnp_amgpar, &pnp_ilupar, &pnp_schpar);
  input_ns_param ns_inpar; itsolver_ns_param ns_itpar; AMG_ns_param ns_amgpar; ILU_param ns_ilupar; Schwarz_param ns_schpar;
  fasp_ns_param_input(ns.c_str(), &ns_inpar);
  fasp_ns_param_init(&ns_inpar, &ns_itpar, &ns_amgpar, &ns_ilupar, &ns_schpar);

  FILE* f = fopen(argv[1], "rb");
  if (!f) { fprintf(stderr, "### ERROR: cannot open %s\n", argv[1]); return 1; }
  int hdr[3];
  if (!rd(f, hdr, 3) || hdr[2] != 5) { fprintf(stderr, "### ERROR: bad header\n"); return 1; }
  const int n = hdr[0], nnz = hdr[1];
  std::vector<int> IA(n + 1), JA(nnz); std::vector<double> val(nnz), rhs(n);
  if (!rd(f, IA.data(), IA.size()) || !rd(f, JA.data(), JA.size()) || !rd(f, val.data(), val.size()) || !rd(f, rhs.data(), rhs.size())) return 1;
  std::vector<std::vector<int>> dof_map(5);
  for (int c = 0; c < 5; ++c) {
    int cnt = 0;
    if (!rd(f, &cnt, 1)) return 1;
    dof_map[c].resize(cnt);
    if (cnt && !rd(f, dof_map[c].data(), (size_t)cnt)) return 1;
  }
  fclose(f);
  dCSRmat A; A.row = A.col = n; A.nnz = nnz; A.IA = IA.data(); A.JA = JA.data(); A.val = val.data();

  Linear_PNP_NS problem(itpar, pnp_itpar, pnp_amgpar, ns_itpar, ns_amgpar);
  problem.get_dofs_fasp(dof_map, {0, 1, 2}, {3, 4});
  std::vector<double> update;
  const INT status = problem.fasp_solve(&A, rhs.data(), update);
  FILE* g = fopen(argv[2], "wb");
  if (!g || fwrite(update.data(), sizeof(double), update.size(), g) != update.size()) return 1;
  fclose(g);
  printf("BLOCK_SOLVE_STATUS %d\n", (int)status);
  return status < 0 ? 2 : 0;
}
