"""CPU oracle for the HelFEM Fock-build hot path (J, K, Vxc).

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy / C restatement of the
reference algorithm (susilehtola/HelFEM); every function cites the reference
file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it -- and
only as the checker, never as the thing measured or shipped.  The product
(``helfem_b200``) never imports, links or executes anything from here.

Pinning status (see DESIGN.md "Oracle"):
  * Gaunt / modified Gaunt coefficients: pinned against the reference's own
    table in src/general/gaunt_test.cpp (tests/golden/gaunt_ref.json).
  * Legendre P_l^m / Q_l^m (x>1): pinned against oracle/_ref (the reference's
    src/legendre/Legendre.h compiled as-is) and its Maple data.
  * Atomic J/K: pinned through the exact 2-node L=0 integrals of
    src/atomic/inttest.cpp and the recorded He/Be RHF energy components of
    tests/refs/ci.json.
  * Diatomic J/K: pinned through the recorded H2 RHF energy components of
    tests/refs/ci.json.
"""
