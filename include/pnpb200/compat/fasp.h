/* pnp-b200 — forwarding header: lets the reference's `extern "C" { #include "fasp.h" }`
 * (include/pde.h:11-14, benchmarks/pnp_exact_solution/linear_pnp.cpp:10-13) resolve to the
 * FASP-compatible surface of libpnpb200.so. Put this directory on the include path instead of
 * ${FASP_DIR}/base/include (CMakeLists.txt:40-48). See INTEGRATION.md, section A. */
#ifndef PNPB200_COMPAT_FASP_H
#define PNPB200_COMPAT_FASP_H
#include "../fasp_compat.h"
#endif
