// C wrapper around the reference's own header-only Legendre implementation,
// compiled from where it lies (/root/reference/src/legendre/Legendre.h) into
// oracle/_ref/liblegendre_ref.so by oracle/Makefile.  No reference source is
// copied into this repository.  Test infrastructure only.
#include "Legendre.h"
#include "Legendre.cpp"

extern "C" {
// out is column-major (lmax+1) x (mmax+1): out[m*(lmax+1)+l]
int ref_plm(double *out, int lmax, int mmax, double x) {
  try { helfem::legendre::plm<double>(out, lmax, mmax, x); } catch (...) { return -1; }
  return 0;
}
int ref_qlm(double *out, int lmax, int mmax, double x) {
  try { helfem::legendre::qlm<double>(out, lmax, mmax, x); } catch (...) { return -1; }
  return 0;
}
}
