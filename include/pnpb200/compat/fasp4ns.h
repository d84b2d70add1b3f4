/* pnp-b200 — forwarding header for the fasp4ns include of benchmarks/pnp_stokes/linear_pnp_ns.h:
 * input_ns_param / itsolver_ns_param / AMG_ns_param, fasp_ns_param_input / _init and
 * fasp_solver_bdcsr_krylov_pnp_stokes are declared in fasp_compat.h. INTEGRATION.md, section B'. */
#ifndef PNPB200_COMPAT_FASP4NS_H
#define PNPB200_COMPAT_FASP4NS_H
#include "../fasp_compat.h"
#endif
