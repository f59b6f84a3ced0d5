"""Oracle (test infrastructure): the reference's Level-3 substep on the HOST CORES with torch - the CPU baseline of
bench.py (``cpu_baseline`` and ``--impl reference``), nothing in the product imports it.

The algorithm is the reference's *as written* - per-element gather of the 8 corner values, element-mean k and rho cp,
``(diag(Me) - Ke) @ T_e`` with the 8 x 8 element matrices of the hex8 / 2x2x2-Gauss discretisation, scatter-add of the
element vectors and lumped masses, ``(sum + F) / sum_M`` (solveMatrixFreeFE cF:582-642), computeStateProperties
cF:2567-2614 before it, the Gaussian source integrated at the 8 Gauss points of every element (computeSourcesL3
cF:2960-3012), the top-surface flux (computeConvRadBC cF:2207-2301) and the ``max(T_amb, .)`` of the substep - but
expressed as dense tensor operations on the structured grid (the 8 corner gathers / scatters are shifted views, the
8 x 8 apply is one GEMM), which is what a multi-threaded CPU array runtime - XLA-CPU for the reference's authors - makes
of it.  NumPy's ``np.add.at`` scatter (oracle/fem.py, the parity oracle) is one to two orders of magnitude slower than
that and single-threaded; this port is the fairer speed baseline.  tests/test_oracle_kats.py pins it to the NumPy
oracle (1e-5 relative: the summation order of the scatter differs).
"""
import numpy as np
import torch

from . import fem
from .setup import SetupProperties  # noqa: F401  (re-export for bench.py)

_CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]  # hex8 local order (dx, dy, dz)


class L3SubstepCPU:
    """One Level-3 window; ``substep(T, S1, laser_xyz, P, dt)`` -> (T_new, S1_new)."""

    def __init__(self, level, properties, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.p = fem.f32props(properties)
        self.nx, self.ny, self.nz = (int(v) for v in level["nodes"])
        coords = fem.getSampleCoords(level)
        N, dNdx, wq = fem.computeQuad3dFemShapeFunctions(coords)
        wq = float(wq[0][0])
        BTB = np.zeros((8, 8), np.float32)
        for d in range(3):
            BTB += dNdx[:, :, d].T @ dNdx[:, :, d]
        self.K0 = torch.from_numpy((BTB * np.float32(wq)).astype(np.float32))            # Ke = K0 * kbar
        self.m0 = torch.from_numpy(((N.T @ N) * np.float32(wq)).sum(axis=0).astype(np.float32))  # Me[a] = m0[a] * mbar
        self.N = torch.from_numpy(N.astype(np.float32))                                  # (8q, 8a)
        self.wq = np.float32(wq)
        x, y, z = (np.asarray(c, np.float32) for c in level["node_coords"])
        # Gauss-point coordinates per axis (getQuadratureCoords cF:3151-3166): element e, local point q in {0, 1}
        g = np.float32(0.57735026918962576)
        lo, hi = np.float32(0.5) * (1 + g), np.float32(0.5) * (1 - g)
        self.gq = [torch.from_numpy(np.stack([lo * c[:-1] + hi * c[1:], hi * c[:-1] + lo * c[1:]], axis=1).astype(np.float32))
                   for c in (x, y, z)]
        # 2-D shape functions of the top face (computeQuad2dFemShapeFunctions cF:778-855)
        N2, _, wq2 = fem.computeQuad2dFemShapeFunctions(np.stack([x[[0, 1, 1, 0, 0, 1, 1, 0]], y[[0, 0, 1, 1, 0, 0, 1, 1]]], axis=1))
        self.N2 = torch.from_numpy(np.asarray(N2, np.float32))
        self.wq2 = float(np.asarray(wq2).reshape(-1)[0])

    def _corner(self, A3, c):
        dx, dy, dz = c
        return A3[dz:self.nz - 1 + dz, dy:self.ny - 1 + dy, dx:self.nx - 1 + dx]

    def _gather(self, A):  # (nn,) -> (8, ne): the element's corner values (convert2XYZ cF:645-689)
        A3 = A.view(self.nz, self.ny, self.nx)
        return torch.stack([self._corner(A3, c).reshape(-1) for c in _CORNERS])

    def _scatter_add(self, E):  # (8, ne) -> (nn,): ordered sum of the 8 corner contributions
        out = torch.zeros(self.nz, self.ny, self.nx, dtype=torch.float32)
        shape = (self.nz - 1, self.ny - 1, self.nx - 1)
        for a, c in enumerate(_CORNERS):
            self._corner(out, c).add_(E[a].view(shape))
        return out.view(-1)

    def state_properties(self, T, S1, n_substrate=0):
        p = self.p
        S2 = T >= float(p["T_liquidus"])
        S3 = (T > float(p["T_solidus"])) & (T < float(p["T_liquidus"]))
        S1n = ((S1 > 0.499) | S2).float()
        if n_substrate:
            S1n[:n_substrate] = 1.0
        S2f, S3f = S2.float(), S3.float()
        k = ((1 - S1n) * (1 - S2f) * float(p["k_powder"]) + S1n * (1 - S2f) * (float(p["k_bulk_coeff_a1"]) * T + float(p["k_bulk_coeff_a0"]))
             + S2f * float(p["k_fluid_coeff_a0"])) / 1000.0
        rc = float(p["rho"]) * ((1 - S2f) * (1 - S3f) * (float(p["cp_solid_coeff_a1"]) * T + float(p["cp_solid_coeff_a0"]))
                                 + S3f * float(p["cp_mushy"]) + S2f * float(p["cp_fluid"]))
        return S1n, k, rc

    def source(self, v, P):
        """computeSourcesL3: Q at the 8 Gauss points of every element (separable Gaussian), assembled with N."""
        p = self.p
        r2, d2 = float(p["laser_radius"]) ** 2, float(p["laser_depth"]) ** 2
        rc_, dc_ = 1.0 / (float(p["laser_radius"]) * np.sqrt(np.pi)), 1.0 / (float(p["laser_depth"]) * np.sqrt(np.pi))
        qx = rc_ * torch.exp(-3.0 * (self.gq[0] - float(v[0])) ** 2 / r2)   # (nex, 2)
        qy = rc_ * torch.exp(-3.0 * (self.gq[1] - float(v[1])) ** 2 / r2)
        qz = dc_ * torch.exp(-3.0 * (self.gq[2] - float(v[2])) ** 2 / d2)
        pc = float(6.0 * np.sqrt(np.float32(3)) * np.float32(P) * p["laser_eta"])
        # Gauss point q = (qx_i, qy_i, qz_i) in the hex8 local order of the corners (ksi, eta, zeta signs)
        Q = torch.stack([(pc * qz[:, None, None, c[2]] * qy[None, :, None, c[1]] * qx[None, None, :, c[0]]).reshape(-1) for c in _CORNERS])
        return self._scatter_add((self.N.t() @ Q) * float(self.wq))       # (Nf @ Q) * wq, cF:3003

    def surface_flux(self, T):
        p = self.p
        top = T.view(self.nz, self.ny, self.nx)[-1]
        T4 = torch.stack([top[:-1, :-1].reshape(-1), top[:-1, 1:].reshape(-1), top[1:, 1:].reshape(-1), top[1:, :-1].reshape(-1)])
        Tq = torch.minimum(self.N2 @ T4, torch.tensor(float(p["T_boiling"]) + 1000.0))
        invT = 1.0 / Tq
        E = float(p["Lev"]) + float(p["cp_fluid"]) * (Tq - float(p["T_amb"]))
        S = float(p["evc"]) * float(p["CP_coeff"]) * torch.exp(-float(p["CT_coeff"]) * (invT - 1.0 / float(p["T_boiling"]))) \
            * torch.sqrt(float(p["CM_coeff"]) * invT) * E
        q = (float(p["h_conv"]) * (float(p["T_amb"]) - Tq) + float(p["sigma_sb"]) * float(p["vareps"]) * (float(p["T_amb"]) ** 4 - Tq ** 4) - S) * 1e-6
        aT = self.N2.t() @ (q * self.wq2)                                   # (4 corners, n_top)
        F = torch.zeros(self.ny, self.nx, dtype=torch.float32)
        n = (self.ny - 1, self.nx - 1)
        F[:-1, :-1] += aT[0].view(n); F[:-1, 1:] += aT[1].view(n); F[1:, 1:] += aT[2].view(n); F[1:, :-1] += aT[3].view(n)
        return F.view(-1)

    def substep(self, T, S1, v, P, dt, n_substrate=0):
        S1n, k, rc = self.state_properties(T, S1, n_substrate)
        F = self.source(v, P)
        F[-self.nx * self.ny:] += self.surface_flux(T)
        Te = self._gather(T)                                                # (8, ne)
        kbar = self._gather(k).mean(dim=0)
        mbar = self._gather(rc).mean(dim=0) / float(dt)
        Me = self.m0[:, None] * mbar[None, :]
        aT = Me * Te - (self.K0 @ Te) * kbar[None, :]                       # (diag(Me) - Ke) @ T_e
        Tn = (self._scatter_add(aT) + F) / self._scatter_add(Me)
        return torch.clamp_min(Tn, float(self.p["T_amb"])), S1n

    def dwell_step(self, T, S1, dt, bc5, n_substrate=0):
        """stepGOMELTDwellTime cF:2617-2664 on a part-scale level: surface flux -> properties -> solve (no source, no
        correction, no clamp) -> the five Dirichlet faces (assignBCs cF:1568-1595: y-, y+, x-, x+, z-; later wins)."""
        _, k, rc = self.state_properties(T, S1, n_substrate)
        F = torch.zeros_like(T)
        F[-self.nx * self.ny:] += self.surface_flux(T)
        Te = self._gather(T)
        kbar = self._gather(k).mean(dim=0)
        mbar = self._gather(rc).mean(dim=0) / float(dt)
        Me = self.m0[:, None] * mbar[None, :]
        aT = Me * Te - (self.K0 @ Te) * kbar[None, :]
        Tn = ((self._scatter_add(aT) + F) / self._scatter_add(Me)).view(self.nz, self.ny, self.nx)
        Tn[:, 0, :] = float(bc5[0]); Tn[:, -1, :] = float(bc5[1]); Tn[:, :, 0] = float(bc5[2]); Tn[:, :, -1] = float(bc5[3])
        Tn[0] = float(bc5[4])
        return Tn.view(-1)
