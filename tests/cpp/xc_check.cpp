// Host-side check of helfem_b200/csrc/xc_builtin.cuh (the expressions the device kernel k_xc_builtin evaluates):
// reads "id n sigma tau" lines, prints exc, vrho, vsigma, vtau in libxc's conventions.  tests/test_host.py compares the
// output with the oracle's symbolic derivatives (oracle/xc.py).
#include <cstdio>

#include "../../helfem_b200/csrc/xc_builtin.cuh"

int main() {
  int id;
  double n, sigma, tau;
  while (std::scanf("%d %lf %lf %lf", &id, &n, &sigma, &tau) == 4) {
    const hfq::xc::D2 e = hfq::xc::energy(id, n, sigma, tau);
    std::printf("%.17e %.17e %.17e %.17e\n", e.v, e.v + n * e.n, n * e.s, n * e.t);
  }
  return 0;
}
