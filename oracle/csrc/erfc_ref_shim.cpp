// C wrapper around the reference's own erfc Green's-function expansion
// (/root/reference/libhelfem/src/erfc_expn.cpp), compiled from where it lies into
// oracle/_ref/liberfc_ref.so by oracle/Makefile.  No reference source is copied into this
// repository.  The reference's utils.h pulls in Eigen (absent here); erfc_expn.cpp needs only
// utils::pi<T>() from it, so the header is pre-empted through its include guard.
// Test infrastructure only.
#define UTILS_H
namespace helfem {
namespace utils {
template <typename T> inline T pi() { return T(3.14159265358979323846264338327950288419716939937510L); }
}  // namespace utils
}  // namespace helfem
#include "erfc_expn.cpp"

extern "C" {
// out[i] = Phi(n, Xi[i], xi[i])  (erfc_expn.cpp: Phi, Phi_short, Phi_general)
int ref_erfc_phi(double *out, unsigned int n, const double *Xi, const double *xi, long npts) {
  try {
    for (long i = 0; i < npts; i++) out[i] = helfem::erfc_expn::Phi<double>(n, Xi[i], xi[i]);
  } catch (...) {
    return -1;
  }
  return 0;
}
}
