"""Host-side (setup-time) finite-element operators for the P2-P1 step.

What the reference obtains from UFL forms + FFC-generated element kernels +
``dolfin.SystemAssembler`` (/root/reference/src/flowcontrol/nsforms.py:238-305,
flowsolver.py:665-701) is restated here with exact reference-element tensors
contracted against the affine geometry.  Everything in this module runs once at
setup; the per-step arithmetic lives in ``csrc/`` (CUDA).

Blocks, all scalar P2 x P2 unless noted:
    M_ab  = int phi_a phi_b                K_ab = int grad phi_a . grad phi_b
    C_ab  = int (U0 . grad phi_b) phi_a    D^{ij}_ab = int phi_b (d_j U0_i) phi_a
    Bx_cb = int psi_c d_x phi_b  (P1 x P2) By_cb likewise
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .mesh import TaylorHoodTables

# --------------------------------------------------------------------------- #
# Radon's 7-point rule, exact to degree 5 on the reference triangle (area 1/2).
# The same table is compiled into the CUDA convection kernel (csrc/fcb200.cu).
# --------------------------------------------------------------------------- #
_S15 = np.sqrt(15.0)
_A1 = (6.0 - _S15) / 21.0
_A2 = (6.0 + _S15) / 21.0
_W1 = (155.0 - _S15) / 2400.0
_W2 = (155.0 + _S15) / 2400.0
RADON7_XI = np.array([1.0 / 3.0, _A1, 1 - 2 * _A1, _A1, _A2, 1 - 2 * _A2, _A2])
RADON7_ETA = np.array([1.0 / 3.0, _A1, _A1, 1 - 2 * _A1, _A2, _A2, 1 - 2 * _A2])
RADON7_W = np.array([9.0 / 80.0, _W1, _W1, _W1, _W2, _W2, _W2])


def p2_shape(xi, eta):
    """phi[q,6], dphi[q,6,2] (reference gradients) in the local order of mesh.py."""
    xi = np.atleast_1d(np.asarray(xi, dtype=np.float64))
    eta = np.atleast_1d(np.asarray(eta, dtype=np.float64))
    lam = np.stack([1.0 - xi - eta, xi, eta], axis=1)  # q,3
    dlam = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # 3,2
    nq = xi.shape[0]
    phi = np.empty((nq, 6))
    dphi = np.empty((nq, 6, 2))
    for i in range(3):
        phi[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dphi[:, i, :] = (4.0 * lam[:, i] - 1.0)[:, None] * dlam[i][None, :]
    for k, (i, j) in enumerate(((1, 2), (0, 2), (0, 1))):
        phi[:, 3 + k] = 4.0 * lam[:, i] * lam[:, j]
        dphi[:, 3 + k, :] = 4.0 * (lam[:, i, None] * dlam[j][None, :] + lam[:, j, None] * dlam[i][None, :])
    return phi, dphi


def p1_shape(xi, eta):
    xi = np.atleast_1d(np.asarray(xi, dtype=np.float64))
    eta = np.atleast_1d(np.asarray(eta, dtype=np.float64))
    return np.stack([1.0 - xi - eta, xi, eta], axis=1)


class ReferenceTensors:
    """Exact integrals over the reference triangle used by the assembly."""

    def __init__(self):
        w = RADON7_W
        phi, dphi = p2_shape(RADON7_XI, RADON7_ETA)
        psi = p1_shape(RADON7_XI, RADON7_ETA)
        self.mass = np.einsum("q,qa,qb->ab", w, phi, phi)  # 6,6
        self.stiff = np.einsum("q,qak,qbl->abkl", w, dphi, dphi)  # 6,6,2,2
        self.div = np.einsum("q,qc,qbk->cbk", w, psi, dphi)  # 3,6,2
        # adv[a,b,c,k] = int phi_a phi_c d_k phi_b
        self.adv = np.einsum("q,qa,qc,qbk->abck", w, phi, phi, dphi)


_REF = None


def reference_tensors() -> ReferenceTensors:
    global _REF
    if _REF is None:
        _REF = ReferenceTensors()
    return _REF


class ScalarBlocks:
    """Constant scalar blocks of one mesh, plus builders for the U0-dependent ones."""

    def __init__(self, tab: TaylorHoodTables):
        self.tab = tab
        ref = reference_tensors()
        nN, nV = tab.nN, tab.nV
        cn = tab.cell_nodes.astype(np.int64)
        self._r22 = np.repeat(cn, 6, axis=1).ravel()
        self._c22 = np.tile(cn, (1, 6)).ravel()
        tv = tab.tri.astype(np.int64)
        self._r12 = np.repeat(tv, 6, axis=1).ravel()
        self._c12 = np.tile(cn, (1, 3)).ravel()
        det, G = tab.detJ, tab.Jinv
        Me = det[:, None, None] * ref.mass[None]
        # grad phi_a . grad phi_b = sum_j (sum_k dref_a,k G[k,j]) (sum_l dref_b,l G[l,j])
        GG = np.einsum("ekj,elj->ekl", G, G)
        Ke = np.einsum("e,abkl,ekl->eab", det, ref.stiff, GG)
        Bxe = np.einsum("e,cbk,ek->ecb", det, ref.div, G[:, :, 0])
        Bye = np.einsum("e,cbk,ek->ecb", det, ref.div, G[:, :, 1])
        self.M = self._scatter22(Me)
        self.K = self._scatter22(Ke)
        self.Bx = sp.coo_matrix((Bxe.ravel(), (self._r12, self._c12)), shape=(nV, nN)).tocsr()
        self.By = sp.coo_matrix((Bye.ravel(), (self._r12, self._c12)), shape=(nV, nN)).tocsr()
        self.Mv = sp.block_diag([self.M, self.M], format="csr")

    def _scatter22(self, Ae) -> sp.csr_matrix:
        nN = self.tab.nN
        return sp.coo_matrix((Ae.ravel(), (self._r22, self._c22)), shape=(nN, nN)).tocsr()

    def advection(self, U0: np.ndarray):
        """Return C and {(i,j): D^{ij}} for a P2 velocity dof vector U0[2 nN]."""
        tab = self.tab
        ref = reference_tensors()
        cn = tab.cell_nodes
        nN = tab.nN
        Ue = np.stack([U0[:nN][cn], U0[nN:][cn]], axis=2)  # e,c,i
        G, det = tab.Jinv, tab.detJ
        # UG[e,c,k] = sum_j U_j[c] G[k,j]
        UG = np.einsum("ecj,ekj->eck", Ue, G)
        Ce = np.einsum("e,abck,eck->eab", det, ref.adv, UG, optimize=True)
        C = self._scatter22(Ce)
        D = {}
        for i in range(2):
            for j in range(2):
                # d_j U0_i = sum_c U_i[c] sum_k dref_c,k G[k,j];  int phi_a phi_b d_k phi_c = adv[a,c,b,k]
                Wck = np.einsum("ec,ek->eck", Ue[:, :, i], G[:, :, j])
                De = np.einsum("e,acbk,eck->eab", det, ref.adv, Wck, optimize=True)
                D[i, j] = self._scatter22(De)
        return C, D

    def convection(self, W: np.ndarray) -> np.ndarray:
        """Host restatement of the per-step convection vector (used at setup by
        the base-flow Newton residual only; the step uses the CUDA kernel)."""
        tab = self.tab
        cn = tab.cell_nodes
        nN = tab.nN
        phi, dref = p2_shape(RADON7_XI, RADON7_ETA)
        We = np.stack([W[:nN][cn], W[nN:][cn]], axis=2)  # e,a,i
        val = np.einsum("qa,eai->eqi", phi, We)
        gref = np.einsum("qak,eai->eqik", dref, We)
        grad = np.einsum("eqik,ekj->eqij", gref, tab.Jinv)
        conv = np.einsum("eqj,eqij->eqi", val, grad)
        Ne = np.einsum("e,q,qa,eqi->eai", tab.detJ, RADON7_W, phi, conv, optimize=True)
        out = np.zeros(2 * nN)
        out[:nN] = np.bincount(cn.ravel(), weights=Ne[:, :, 0].ravel(), minlength=nN)
        out[nN:] = np.bincount(cn.ravel(), weights=Ne[:, :, 1].ravel(), minlength=nN)
        return out

    def saddle_point(self, c_mass: float, Re: float, U0=None, shift: float = 0.0, linearised: bool = True):
        """A(c) = [[F+Dxx, Dxy, -Bx^T],[Dyx, F+Dyy, -By^T],[-Bx, -By, 0]], F = cM + C + K/Re - shift M.

        ``linearised=False`` drops the (u.grad)U0 blocks (Picard operator,
        nsforms.py:178-183); ``U0=None`` drops all advection."""
        F = (c_mass - shift) * self.M + self.K / Re
        if U0 is None:
            blocks = [[F, None, -self.Bx.T], [None, F, -self.By.T], [-self.Bx, -self.By, None]]
        else:
            C, D = self.advection(U0)
            F = F + C
            if linearised:
                blocks = [
                    [F + D[0, 0], D[0, 1], -self.Bx.T],
                    [D[1, 0], F + D[1, 1], -self.By.T],
                    [-self.Bx, -self.By, None],
                ]
            else:
                blocks = [[F, None, -self.Bx.T], [None, F, -self.By.T], [-self.Bx, -self.By, None]]
        A = sp.bmat(blocks, format="csr")
        A.sum_duplicates()
        A.sort_indices()
        return A
