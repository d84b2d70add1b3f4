// pnp-b200 host side — domain parameters and mesh construction with the interface of the
// reference's src/domain.cpp: domain_param_input_init (:18-39, the defaults), domain_param_check
// (:50-69), domain_param_input (:79-271: whitespace-separated `key = value` tokens; a token that
// starts with '%', '[' or '|' comments out the rest of its line; an unknown key warns and skips
// its line; a malformed entry STOPS the reading and what was read so far stands — the status of
// the syntax error is overwritten by the range check, :266), print_domain_param (:273-287) and
// domain_build (:290-311) = centred dolfin::BoxMesh(-L/2, +L/2, grid_x, grid_y, grid_z).
// Two deliberate differences, both at the edge of the process: where the reference ends the
// program through fasp_chkerr (file missing, range check failed) domain_param_input RETURNS the
// status; and the mesh is generated on the device (pnpb200_create_box), so domain_build returns
// the handle (nullptr where the reference returns the empty dolfin::Mesh).
// Pinned against the reference's own src/domain.cpp compiled where it lies
// (oracle/_ref/libpnpref_host.so, tests/test_host_logic.py, tests/golden/host_logic.json).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "pnpb200/pnpb200.h"

struct domain_param {
  REAL ref_length, length_x, length_y, length_z, length_time;
  INT grid_x, grid_y, grid_z, grid_time;
  char mesh_output[128], mesh_file[128], subdomain_file[128], surface_file[128];
};

inline SHORT domain_param_input_init(domain_param* p)
{
  p->ref_length = 1.0; p->length_x = p->length_y = p->length_z = 1.0; p->length_time = 0.0;
  p->grid_x = p->grid_y = p->grid_z = 10; p->grid_time = 0;
  strcpy(p->mesh_output, "./output/mesh.pvd");
  strcpy(p->mesh_file, "none"); strcpy(p->subdomain_file, "none"); strcpy(p->surface_file, "none");
  return FASP_SUCCESS;
}

inline SHORT domain_param_check(const domain_param* p)
{
  const bool bad = p->ref_length < 0.0 || p->length_x < 0.0 || p->length_y < 0.0 || p->length_z < 0.0 || p->length_time < 0.0
                || p->grid_x < 2 || p->grid_y < 2 || p->grid_z < 2 || p->grid_time < 0;
  return bad ? ERROR_INPUT_PAR : FASP_SUCCESS;
}

inline SHORT domain_param_input(const char* filenm, domain_param* p)
{
  domain_param_input_init(p);
  if (filenm == nullptr) return FASP_SUCCESS;          // no file: the defaults
  FILE* fp = fopen(filenm, "r");
  if (!fp) { printf("### ERROR: Could not open file %s...\n", filenm); return ERROR_OPEN_FILE; }
  struct Key { const char* name; char kind; void* dst; };       // kind: d = REAL, i = INT, s = char[128]
  const Key keys[] = {
    {"ref_length", 'd', &p->ref_length}, {"length_x", 'd', &p->length_x}, {"length_y", 'd', &p->length_y},
    {"length_z", 'd', &p->length_z}, {"length_time", 'd', &p->length_time},
    {"grid_x", 'i', &p->grid_x}, {"grid_y", 'i', &p->grid_y}, {"grid_z", 'i', &p->grid_z}, {"grid_time", 'i', &p->grid_time},
    {"mesh_output", 's', p->mesh_output}, {"mesh_file", 's', p->mesh_file},
    {"subdomain_file", 's', p->subdomain_file}, {"surface_file", 's', p->surface_file}};
  char tok[500], rest[500];
  auto skip_line = [&]() { if (!fgets(rest, sizeof rest, fp)) rest[0] = 0; };
  for (;;) {
    if (fscanf(fp, "%499s", tok) != 1) break;                               // end of file
    if (tok[0] == '[' || tok[0] == '%' || tok[0] == '|') { skip_line(); continue; }
    const Key* k = nullptr;
    for (const Key& q : keys) if (!strcmp(tok, q.name)) { k = &q; break; }
    if (!k) { printf("### WARNING: Unknown input keyword %s!\n", tok); skip_line(); continue; }
    if (fscanf(fp, "%499s", tok) != 1 || strcmp(tok, "=") != 0) break;      // malformed entry: reading stops here
    bool ok;
    if (k->kind == 'd') { double v; ok = fscanf(fp, "%lf", &v) == 1; if (ok) *(REAL*)k->dst = v; }
    else if (k->kind == 'i') { int v; ok = fscanf(fp, "%d", &v) == 1; if (ok) *(INT*)k->dst = v; }
    else { ok = fscanf(fp, "%499s", tok) == 1; if (ok) strncpy((char*)k->dst, tok, 128); }
    if (!ok) break;
    skip_line();
  }
  fclose(fp);
  return domain_param_check(p);
}

inline void print_domain_param(const domain_param& p)
{
  printf("\tSuccessfully read-in domain parameters\n");
  printf("\tThe reference length is %e meters\n", p.ref_length);
  if (strcmp(p.mesh_file, "none") == 0) {
    printf("\tDomain: %f x %f x %f\n", p.length_x, p.length_y, p.length_z);
    printf("\tGrid:   %d x %d x %d\n\n", p.grid_x, p.grid_y, p.grid_z);
  } else {
    printf("\tMesh file:      %s\n", p.mesh_file);
    printf("\tSubdomain file: %s\n", p.subdomain_file);
    printf("\tSurface file:   %s\n\n", p.surface_file);
  }
  fflush(stdout);
}

// The box domain_build asks for: corners -L/2, +L/2 and the grid. Returns false (with the reference's messages) where
// the reference returns the empty Mesh: a NaN length, or a mesh file (reading meshes is unsupported upstream too).
inline bool domain_box(const domain_param& domain, double p0[3], double p1[3], int n[3])
{
  if (std::isnan(domain.length_x) || std::isnan(domain.length_y) || std::isnan(domain.length_z)) {
    printf("### WARNING : invalid domain lengths!\n"); return false;
  }
  if (strcmp(domain.mesh_file, "none") != 0) {
    printf("### ERROR : Reading in meshes is currently unsupported: %s...\n\n", domain.mesh_file); return false;
  }
  p0[0] = -domain.length_x / 2; p0[1] = -domain.length_y / 2; p0[2] = -domain.length_z / 2;
  p1[0] = domain.length_x / 2; p1[1] = domain.length_y / 2; p1[2] = domain.length_z / 2;
  n[0] = domain.grid_x; n[1] = domain.grid_y; n[2] = domain.grid_z;
  return true;
}

// nullptr on invalid input (like the empty dolfin::Mesh) or when the device refuses the box (a grid below 1 can only get
// here past domain_param_check)
inline pnpb200_handle* domain_build(const domain_param& domain, int device = 0)
{
  double p0[3], p1[3]; int n[3];
  if (!domain_box(domain, p0, p1, n)) return nullptr;
  pnpb200_handle* h = nullptr;
  if (pnpb200_create_box(&h, n[0], n[1], n[2], domain.length_x, domain.length_y, domain.length_z, device) != FASP_SUCCESS) return nullptr;
  return h;
}
