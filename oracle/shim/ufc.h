// TEST INFRASTRUCTURE ONLY (oracle/): declarative stand-in for the UFC 2016.1.0
// abstract interface (<ufc.h> of FEniCS, absent from this image). It exists so the
// UFC section of the reference's FFC-generated headers (include/EAFE.h:1-5513,
// benchmarks/pnp_exact_solution/vector_linear_pnp_forms.h:1-9237) can be compiled
// where they lie, unmodified, into oracle/_ref/libpnpref.so. The virtual signatures
// are the ones those headers override. Nothing here computes anything.
#ifndef PNPB200_ORACLE_SHIM_UFC_H
#define PNPB200_ORACLE_SHIM_UFC_H
#include <cstddef>
#include <vector>
#include <cmath>
#include "ufc_geometry.h"

// include/EAFE.h:4862 uses DOLFIN_EPS un-namespaced inside the UFC section; upstream
// DOLFIN defines it as 3.0e-16 (dolfin/common/constants.h) [EXT].
#ifndef DOLFIN_EPS
#define DOLFIN_EPS 3.0e-16
#endif

namespace ufc {

enum class shape { interval, triangle, quadrilateral, tetrahedron, hexahedron, vertex };

class cell {
public:
  shape cell_shape;
  std::size_t topological_dimension, geometric_dimension;
  std::vector<std::vector<std::size_t>> entity_indices;
  std::size_t index; int local_facet; int orientation; int mesh_identifier;
};

class function {
public:
  virtual ~function() {}
  virtual void evaluate(double* values, const double* coordinates, const cell& c) const = 0;
};

class finite_element {
public:
  virtual ~finite_element() {}
  virtual const char* signature() const = 0;
  virtual shape cell_shape() const = 0;
  virtual std::size_t topological_dimension() const = 0;
  virtual std::size_t geometric_dimension() const = 0;
  virtual std::size_t space_dimension() const = 0;
  virtual std::size_t value_rank() const = 0;
  virtual std::size_t value_dimension(std::size_t i) const = 0;
  virtual std::size_t value_size() const = 0;
  virtual std::size_t reference_value_rank() const = 0;
  virtual std::size_t reference_value_dimension(std::size_t i) const = 0;
  virtual std::size_t reference_value_size() const = 0;
  virtual std::size_t degree() const = 0;
  virtual const char* family() const = 0;
  virtual void evaluate_basis(std::size_t i, double* values, const double* x,
                              const double* coordinate_dofs, int cell_orientation) const = 0;
  virtual void evaluate_basis_all(double* values, const double* x,
                                  const double* coordinate_dofs, int cell_orientation) const = 0;
  virtual void evaluate_basis_derivatives(std::size_t i, std::size_t n, double* values, const double* x,
                                          const double* coordinate_dofs, int cell_orientation) const = 0;
  virtual void evaluate_basis_derivatives_all(std::size_t n, double* values, const double* x,
                                              const double* coordinate_dofs, int cell_orientation) const = 0;
  virtual double evaluate_dof(std::size_t i, const function& f, const double* coordinate_dofs,
                              int cell_orientation, const cell& c) const = 0;
  virtual void evaluate_dofs(double* values, const function& f, const double* coordinate_dofs,
                             int cell_orientation, const cell& c) const = 0;
  virtual void interpolate_vertex_values(double* vertex_values, const double* dof_values,
                                         const double* coordinate_dofs, int cell_orientation,
                                         const cell& c) const = 0;
  virtual void tabulate_dof_coordinates(double* dof_coordinates, const double* coordinate_dofs) const = 0;
  virtual std::size_t num_sub_elements() const = 0;
  virtual finite_element* create_sub_element(std::size_t i) const = 0;
  virtual finite_element* create() const = 0;
};

class dofmap {
public:
  virtual ~dofmap() {}
  virtual const char* signature() const = 0;
  virtual bool needs_mesh_entities(std::size_t d) const = 0;
  virtual std::size_t topological_dimension() const = 0;
  virtual std::size_t global_dimension(const std::vector<std::size_t>& num_global_entities) const = 0;
  virtual std::size_t num_element_dofs() const = 0;
  virtual std::size_t num_facet_dofs() const = 0;
  virtual std::size_t num_entity_dofs(std::size_t d) const = 0;
  virtual void tabulate_dofs(std::size_t* dofs, const std::vector<std::size_t>& num_global_entities,
                             const std::vector<std::vector<std::size_t>>& entity_indices) const = 0;
  virtual void tabulate_facet_dofs(std::size_t* dofs, std::size_t facet) const = 0;
  virtual void tabulate_entity_dofs(std::size_t* dofs, std::size_t d, std::size_t i) const = 0;
  virtual std::size_t num_sub_dofmaps() const = 0;
  virtual dofmap* create_sub_dofmap(std::size_t i) const = 0;
  virtual dofmap* create() const = 0;
};

class coordinate_mapping { public: virtual ~coordinate_mapping() {} };

class integral {
public:
  virtual ~integral() {}
  virtual const std::vector<bool>& enabled_coefficients() const = 0;
};

class cell_integral : public integral {
public:
  virtual void tabulate_tensor(double* A, const double* const* w, const double* coordinate_dofs,
                               int cell_orientation) const = 0;
};
class exterior_facet_integral : public integral {};
class interior_facet_integral : public integral {};
class vertex_integral : public integral {};
class custom_integral : public integral {};
class cutcell_integral : public integral {};
class interface_integral : public integral {};
class overlap_integral : public integral {};

class form {
public:
  virtual ~form() {}
  virtual const char* signature() const = 0;
  virtual std::size_t rank() const = 0;
  virtual std::size_t num_coefficients() const = 0;
  virtual std::size_t original_coefficient_position(std::size_t i) const = 0;
  virtual finite_element* create_coordinate_finite_element() const = 0;
  virtual dofmap* create_coordinate_dofmap() const = 0;
  virtual coordinate_mapping* create_coordinate_mapping() const = 0;
  virtual finite_element* create_finite_element(std::size_t i) const = 0;
  virtual dofmap* create_dofmap(std::size_t i) const = 0;
  virtual std::size_t max_cell_subdomain_id() const = 0;
  virtual std::size_t max_exterior_facet_subdomain_id() const = 0;
  virtual std::size_t max_interior_facet_subdomain_id() const = 0;
  virtual std::size_t max_vertex_subdomain_id() const = 0;
  virtual std::size_t max_custom_subdomain_id() const = 0;
  virtual std::size_t max_cutcell_subdomain_id() const = 0;
  virtual std::size_t max_interface_subdomain_id() const = 0;
  virtual std::size_t max_overlap_subdomain_id() const = 0;
  virtual bool has_cell_integrals() const = 0;
  virtual bool has_exterior_facet_integrals() const = 0;
  virtual bool has_interior_facet_integrals() const = 0;
  virtual bool has_vertex_integrals() const = 0;
  virtual bool has_custom_integrals() const = 0;
  virtual bool has_cutcell_integrals() const = 0;
  virtual bool has_interface_integrals() const = 0;
  virtual bool has_overlap_integrals() const = 0;
  virtual cell_integral* create_cell_integral(std::size_t subdomain_id) const = 0;
  virtual exterior_facet_integral* create_exterior_facet_integral(std::size_t subdomain_id) const = 0;
  virtual interior_facet_integral* create_interior_facet_integral(std::size_t subdomain_id) const = 0;
  virtual vertex_integral* create_vertex_integral(std::size_t subdomain_id) const = 0;
  virtual custom_integral* create_custom_integral(std::size_t subdomain_id) const = 0;
  virtual cutcell_integral* create_cutcell_integral(std::size_t subdomain_id) const = 0;
  virtual interface_integral* create_interface_integral(std::size_t subdomain_id) const = 0;
  virtual overlap_integral* create_overlap_integral(std::size_t subdomain_id) const = 0;
  virtual cell_integral* create_default_cell_integral() const = 0;
  virtual exterior_facet_integral* create_default_exterior_facet_integral() const = 0;
  virtual interior_facet_integral* create_default_interior_facet_integral() const = 0;
  virtual vertex_integral* create_default_vertex_integral() const = 0;
  virtual custom_integral* create_default_custom_integral() const = 0;
  virtual cutcell_integral* create_default_cutcell_integral() const = 0;
  virtual interface_integral* create_default_interface_integral() const = 0;
  virtual overlap_integral* create_default_overlap_integral() const = 0;
};

} // namespace ufc
#endif
