"""Gate buffers as the kernels expect them.

Plays the role of ``CustomMatrices`` (/root/reference/src/qibojit/backends/matrices.py:7-70):
controlled gates collapse to their target-only matrix (matrices.py:13-57), ``U1``/``CU1`` to
the scalar phase (matrices.py:28-33) and ``fSim``/``GeneralizedfSim`` to a 5-vector
(matrices.py:62-70).  The uncontrolled matrices are the textbook definitions that qibo's
``NumpyMatrices`` provides upstream (qibo is a third-party dependency of the reference and is
not vendored, so they are restated here).
"""

import numpy as np


class CustomMatrices:
    def __init__(self, dtype="complex128"):
        self.dtype = np.dtype(dtype)

    def _cast(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=self.dtype))

    # ---- fixed one-qubit gates
    @property
    def I(self):
        return self._cast(np.eye(2))

    @property
    def H(self):
        return self._cast(np.array([[1, 1], [1, -1]]) / np.sqrt(2))

    @property
    def X(self):
        return self._cast([[0, 1], [1, 0]])

    @property
    def Y(self):
        return self._cast([[0, -1j], [1j, 0]])

    @property
    def Z(self):
        return self._cast([[1, 0], [0, -1]])

    @property
    def S(self):
        return self._cast([[1, 0], [0, 1j]])

    @property
    def SDG(self):
        return self._cast([[1, 0], [0, -1j]])

    @property
    def T(self):
        return self._cast([[1, 0], [0, np.exp(1j * np.pi / 4)]])

    @property
    def TDG(self):
        return self._cast([[1, 0], [0, np.exp(-1j * np.pi / 4)]])

    @property
    def SX(self):
        return self._cast(np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]) / 2)

    @property
    def SXDG(self):
        return self._cast(np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]]) / 2)

    # ---- parametrised one-qubit gates
    def RX(self, theta):
        c, s = np.cos(theta / 2), -1j * np.sin(theta / 2)
        return self._cast([[c, s], [s, c]])

    def RY(self, theta):
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        return self._cast([[c, -s], [s, c]])

    def RZ(self, theta):
        p = np.exp(0.5j * theta)
        return self._cast([[np.conj(p), 0], [0, p]])

    def GPI(self, phi):
        p = np.exp(1j * phi)
        return self._cast([[0, np.conj(p)], [p, 0]])

    def GPI2(self, phi):
        p = np.exp(1j * phi)
        return self._cast(np.array([[1, -1j * np.conj(p)], [-1j * p, 1]]) / np.sqrt(2))

    def U1(self, theta):
        # matrices.py:28-30: the kernel takes the scalar phase
        return np.asarray(np.exp(1j * theta), dtype=self.dtype)

    def U2(self, phi, lam):
        ep, em = np.exp(0.5j * (phi + lam)), np.exp(0.5j * (phi - lam))
        return self._cast(np.array([[np.conj(ep), -np.conj(em)], [em, ep]]) / np.sqrt(2))

    def U3(self, theta, phi, lam):
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        ep, em = np.exp(0.5j * (phi + lam)), np.exp(0.5j * (phi - lam))
        return self._cast([[np.conj(ep) * c, -np.conj(em) * s], [em * s, ep * c]])

    # ---- controlled gates -> target-only matrix (matrices.py:13-57)
    CNOT = X
    TOFFOLI = X
    CY = Y
    CZ = Z
    CCZ = Z
    CH = H
    CSX = SX
    CSXDG = SXDG

    def CRX(self, theta):
        return self.RX(theta)

    def CRY(self, theta):
        return self.RY(theta)

    def CRZ(self, theta):
        return self.RZ(theta)

    def CU1(self, theta):
        return self.U1(theta)

    def CU2(self, phi, lam):
        return self.U2(phi, lam)

    def CU3(self, theta, phi, lam):
        return self.U3(theta, phi, lam)

    def DEUTSCH(self, theta):
        return self._cast(1j * self.RX(2 * theta))

    # ---- two-qubit gates
    @property
    def SWAP(self):
        return self._cast([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])

    @property
    def iSWAP(self):
        return self._cast([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])

    @property
    def SiSWAP(self):
        a, b = 1 / np.sqrt(2), 1j / np.sqrt(2)
        return self._cast([[1, 0, 0, 0], [0, a, b, 0], [0, b, a, 0], [0, 0, 0, 1]])

    @property
    def FSWAP(self):
        return self._cast([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, -1]])

    def fSim(self, theta, phi):
        # matrices.py:62-66
        cost = np.cos(theta) + 0j
        isint = -1j * np.sin(theta)
        return self._cast([cost, isint, isint, cost, np.exp(-1j * phi)])

    def GeneralizedfSim(self, u, phi):
        # matrices.py:68-70
        u = np.asarray(u)
        return self._cast([u[0, 0], u[0, 1], u[1, 0], u[1, 1], np.exp(-1j * phi)])

    def RXX(self, theta):
        c, s = np.cos(theta / 2), -1j * np.sin(theta / 2)
        return self._cast([[c, 0, 0, s], [0, c, s, 0], [0, s, c, 0], [s, 0, 0, c]])

    def RYY(self, theta):
        c, s = np.cos(theta / 2), 1j * np.sin(theta / 2)
        return self._cast([[c, 0, 0, s], [0, c, -s, 0], [0, -s, c, 0], [s, 0, 0, c]])

    def RZZ(self, theta):
        p = np.exp(0.5j * theta)
        return self._cast(np.diag([np.conj(p), p, p, np.conj(p)]))

    def Unitary(self, u):
        return self._cast(u)

    def FusedGate(self, u):
        return self._cast(u)
