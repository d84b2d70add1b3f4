// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for <dolfin.h> so that the reference's plain-C++ host logic
// (src/newton_status.cpp, src/domain.cpp, src/dirichlet.cpp) compiles where it lies without DOLFIN: the few types
// those files name, with no behaviour. domain_build's BoxMesh records its arguments, nothing more.
#ifndef PNP_ORACLE_SHIM_DOLFIN_H
#define PNP_ORACLE_SHIM_DOLFIN_H
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstring>
#include <string>
#include <vector>
namespace dolfin {
// what src/dirichlet.cpp names: Array (a view), SubDomain::inside, Expression::eval
template <typename T> struct Array {
  std::size_t n; T* p;
  Array(std::size_t n_, T* p_) : n(n_), p(p_) {}
  std::size_t size() const { return n; }
  const T& operator[](std::size_t i) const { return p[i]; }
  T& operator[](std::size_t i) { return p[i]; }
};
struct SubDomain {
  virtual ~SubDomain() {}
  virtual bool inside(const Array<double>&, bool) const { return false; }
};
struct Expression {
  std::size_t value_dim;
  Expression() : value_dim(1) {}
  explicit Expression(std::size_t d) : value_dim(d) {}
  virtual ~Expression() {}
  virtual void eval(Array<double>&, const Array<double>&) const {}
};
struct Point {
  double c[3];
  Point(double x = 0.0, double y = 0.0, double z = 0.0) { c[0] = x; c[1] = y; c[2] = z; }
};
struct Mesh {
  bool built = false;
  double p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
  int n[3] = {0, 0, 0};
};
struct BoxMesh : Mesh {
  BoxMesh(const Point& a, const Point& b, int nx, int ny, int nz)
  {
    built = true;
    for (int k = 0; k < 3; ++k) { p0[k] = a.c[k]; p1[k] = b.c[k]; }
    n[0] = nx; n[1] = ny; n[2] = nz;
  }
};
}
#endif
