// pnp-b200 host side — domain parameters and mesh construction with the interface of the
// reference's src/domain.cpp: `key = value` file (:79-271; '%', '[' and '|' start comments), and
// domain_build (:290-311) = centred dolfin::BoxMesh(-L/2, +L/2, grid_x, grid_y, grid_z). Here the
// mesh is generated on the device (pnpb200_create_box), so domain_build returns the handle.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "pnpb200/pnpb200.h"

struct domain_param {
  double ref_length = 1.0;
  double length_x = NAN, length_y = NAN, length_z = NAN, length_time = 0.0;
  int grid_x = 0, grid_y = 0, grid_z = 0, grid_time = 0;
  char mesh_output[256] = "none", mesh_file[256] = "none", subdomain_file[256] = "none", surface_file[256] = "none";
};

inline SHORT domain_param_input(const char* filenm, domain_param* p)
{
  FILE* fp = fopen(filenm, "r");
  if (!fp) { printf("### ERROR: %s file does not exist!\n", filenm); return ERROR_OPEN_FILE; }
  char line[1024];
  SHORT status = FASP_SUCCESS;
  while (fgets(line, sizeof line, fp)) {
    char key[512], eq[16], val[512];
    if (sscanf(line, "%511s", key) != 1) continue;
    if (key[0] == '%' || key[0] == '[' || key[0] == '|') continue;
    if (sscanf(line, "%511s %15s %511s", key, eq, val) != 3 || strcmp(eq, "=") != 0) { status = ERROR_INPUT_PAR; continue; }
    if (!strcmp(key, "ref_length")) p->ref_length = atof(val);
    else if (!strcmp(key, "length_x")) p->length_x = atof(val);
    else if (!strcmp(key, "length_y")) p->length_y = atof(val);
    else if (!strcmp(key, "length_z")) p->length_z = atof(val);
    else if (!strcmp(key, "length_time")) p->length_time = atof(val);
    else if (!strcmp(key, "grid_x")) p->grid_x = atoi(val);
    else if (!strcmp(key, "grid_y")) p->grid_y = atoi(val);
    else if (!strcmp(key, "grid_z")) p->grid_z = atoi(val);
    else if (!strcmp(key, "grid_time")) p->grid_time = atoi(val);
    else if (!strcmp(key, "mesh_output")) strncpy(p->mesh_output, val, 255);
    else if (!strcmp(key, "mesh_file")) strncpy(p->mesh_file, val, 255);
    else if (!strcmp(key, "subdomain_file")) strncpy(p->subdomain_file, val, 255);
    else if (!strcmp(key, "surface_file")) strncpy(p->surface_file, val, 255);
    else printf("### WARNING: Unknown input keyword %s!\n", key);
  }
  fclose(fp);
  return status;
}

// returns nullptr (and prints the reference's messages) on invalid input, like the empty dolfin::Mesh
inline pnpb200_handle* domain_build(const domain_param& domain, int device = 0)
{
  if (std::isnan(domain.length_x) || std::isnan(domain.length_y) || std::isnan(domain.length_z)) {
    printf("### WARNING : invalid domain lengths!\n"); return nullptr;
  }
  if (domain.grid_x < 1 || domain.grid_y < 1 || domain.grid_z < 1) { printf("### WARNING : invalid domain grid!\n"); return nullptr; }
  if (strcmp(domain.mesh_file, "none") != 0) {
    printf("### ERROR : Reading in meshes is currently unsupported: %s...\n\n", domain.mesh_file); return nullptr;
  }
  pnpb200_handle* h = nullptr;
  if (pnpb200_create_box(&h, domain.grid_x, domain.grid_y, domain.grid_z, domain.length_x, domain.length_y,
                         domain.length_z, device) != FASP_SUCCESS) return nullptr;
  return h;
}
