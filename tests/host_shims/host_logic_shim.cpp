// Test shim (tests/test_host_logic.py): the host mirrors modular-pnp_b200/host/newton_status.hpp and host/domain.hpp
// behind the same C entry points oracle/ref_wrappers_host.inc gives the reference's src/newton_status.cpp and
// src/domain.cpp, so that one driver can run both.
#include "newton_status.hpp"
#include "domain.hpp"

extern "C" {

void* mir_newton_new(unsigned long max_it, double initial_residual, double rel_tol, double max_tol)
{ return new Newton_Status(max_it, initial_residual, rel_tol, max_tol); }
void mir_newton_delete(void* p) { delete (Newton_Status*)p; }
void mir_newton_update_iteration(void* p) { ((Newton_Status*)p)->update_iteration(); }
void mir_newton_update_residuals(void* p, double r, double m) { ((Newton_Status*)p)->update_residuals(r, m); }
void mir_newton_update_max_residual(void* p, double m) { ((Newton_Status*)p)->update_max_residual(m); }
void mir_newton_update_rel_residual(void* p, double r) { ((Newton_Status*)p)->update_rel_residual(r); }
int mir_newton_needs_to_iterate(void* p) { return ((Newton_Status*)p)->needs_to_iterate() ? 1 : 0; }
int mir_newton_converged(void* p) { return ((Newton_Status*)p)->converged() ? 1 : 0; }
void mir_newton_print_status(void* p) { ((Newton_Status*)p)->print_status(); fflush(stdout); }
void mir_newton_state(void* p, double out[3])
{ Newton_Status* s = (Newton_Status*)p; out[0] = (double)s->iteration; out[1] = s->relative_residual; out[2] = s->max_residual; }

int mir_domain_param_input(const char* filenm, double d[5], int g[4], char s[512])
{
  static domain_param p;
  memset(&p, 0, sizeof p);
  const int status = domain_param_input(filenm, &p);
  d[0] = p.ref_length; d[1] = p.length_x; d[2] = p.length_y; d[3] = p.length_z; d[4] = p.length_time;
  g[0] = p.grid_x; g[1] = p.grid_y; g[2] = p.grid_z; g[3] = p.grid_time;
  memcpy(s, p.mesh_output, 128); memcpy(s + 128, p.mesh_file, 128); memcpy(s + 256, p.subdomain_file, 128); memcpy(s + 384, p.surface_file, 128);
  return status;
}

int mir_domain_build(const double d[5], const int g[4], const char* mesh_file, double box[9])
{
  domain_param p;
  domain_param_input_init(&p);
  p.ref_length = d[0]; p.length_x = d[1]; p.length_y = d[2]; p.length_z = d[3]; p.length_time = d[4];
  p.grid_x = g[0]; p.grid_y = g[1]; p.grid_z = g[2]; p.grid_time = g[3];
  strncpy(p.mesh_file, mesh_file, 127); p.mesh_file[127] = 0;
  double p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0}; int n[3] = {0, 0, 0};
  const bool built = domain_box(p, p0, p1, n);
  fflush(stdout);
  for (int k = 0; k < 3; ++k) { box[k] = built ? p0[k] : 0.0; box[3 + k] = built ? p1[k] : 0.0; box[6 + k] = built ? n[k] : 0; }
  return built ? 1 : 0;
}

}
