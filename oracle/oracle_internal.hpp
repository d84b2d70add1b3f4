// TEST INFRASTRUCTURE ONLY (oracle/) — internal declarations shared by the oracle sources.
#ifndef PNP_ORACLE_INTERNAL_HPP
#define PNP_ORACLE_INTERNAL_HPP
#include <vector>

namespace orc {

typedef void (*eafe_fn)(double*, const double*, const double*, const double*, const double*, const double*);
typedef void (*jac_fn)(double*, const double*, const double*, const double*, const double*, const double*);
typedef void (*res_fn)(double*, const double*, const double*, const double*, const double*,
                       const double*, const double*, const double*);
struct ElementKernels { eafe_fn eafe; jac_fn jac; res_fn res; int is_ref; };
extern ElementKernels kernels;

void tet_geometry(const double* x, double g[4][3], double* absdet);
void eafe_element(double* A, const double* eta, const double* alpha, const double* beta,
                  const double* gamma, const double* x);
void jac_element(double* A, const double* uu, const double* eps, const double* D, const double* q,
                 const double* x);
void res_element(double* b, const double* uu, const double* eps, const double* fc, const double* D,
                 const double* q, const double* reac, const double* x);

double marker_element(const double* mu, const double* logw, const double* D, const double* x);

// FASP dBSRmat restated (nb = 3, storage_manner 0: row-major blocks)
struct BSR {
  int ROW = 0, COL = 0, NNZ = 0;
  std::vector<int> IA, JA;
  std::vector<double> val;
};

} // namespace orc
#endif
