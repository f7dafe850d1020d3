"""Device-side compute_tei of the diatomic basis (csrc/tei_device.cu) against the host setup (which
tests/test_host.py pins on the oracle restatement of src/diatomic/basis.cpp:1382-1547): the reconstructed 2-channel
kernels W = B sigma B^T channel by channel, and J / K built from the device tables against the oracle."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _recon(tab, ilm, iel):
    sm, bg, B, sig = tab.block(ilm, iel)
    return (B * sig[None, :]) @ B.T, sm, bg, B.shape[1]


@pytest.mark.parametrize("lmax_per_m,nelem", [((3, 2), 2), ((8, 8, 7), 3)])
def test_device_tei_matches_host(hb, lmax_per_m, nelem):
    th = hb.Tables.diatomic(7, 7, 2.07, list(lmax_per_m), nelem)
    td = hb.Tables.diatomic(7, 7, 2.07, list(lmax_per_m), nelem, device=0)
    assert th.nlm == td.nlm and th.Nel == td.Nel and np.array_equal(th.lmL, td.lmL) and np.array_equal(th.lmM, td.lmM)
    worst = 0.0
    for ilm in range(th.nlm):
        for iel in range(th.Nel):
            Wh, smh, bgh, rh = _recon(th, ilm, iel)
            Wd, smd, bgd, rd = _recon(td, ilm, iel)
            assert np.array_equal(smh, smd) and np.array_equal(bgh, bgd)      # cross-element factors: same host code
            scale = np.abs(Wh).max()
            # both factorisations stop at 1e-12 of the largest diagonal element: each reconstructs W to that level
            worst = max(worst, np.abs(Wh - Wd).max() / scale)
            assert abs(rh - rd) <= 2, (ilm, iel, rh, rd)
    assert worst < 5e-12, worst


def test_device_tei_fock_matrices_match_oracle(hb):
    ob = cases.oracle_diatomic(7, 7, 2.07, (3, 2), 2)
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [3, 2], 2, tei_on_device=True).compute_tei()
    P = cases.random_density(ob.Nbf(), 3, 31, cases.m_blocks(ob.mval, ob.Nrad(), True))
    J, K = basis.coulomb(P), basis.exchange(P)
    assert cases.relerr(J, ob.coulomb(P)) < 1e-10 and cases.relerr(K, ob.exchange(P)) < 1e-10
