"""numpy restatement of the reference's HOST side of one time step for the finite-difference FULL SPHERE.  TEST INFRASTRUCTURE ONLY.

Same purpose as oracle/lmloop.py (read its header first): the reference pins the radial loop only end to end, through
samples/*/reference.out.  This module restates the host that `samples/full_sphere` runs (BASELINE config 4's geometry:
Marti et al. 2014 benchmark 1, Boussinesq hydro + internal heating, radratio = 0, `radial_scheme='FD'`, CNAB2, restart
from the shipped checkpoint) so that the radial loop -- handed in as a callable: the CPU oracle or the CUDA library through
the C ABI -- can be driven through the 100 steps of the reference's own autotest.  This is the one sample that exercises
the centre level (v_center_sphere), l_R(nR) < l_max (l_var_l), the double-curl form of get_dwdt and stress-free top.

It is NOT on the product path (magic_b200/ never imports oracle/); the product keeps the Fortran host.

Restated reference routines (file:line relative to /root/reference/src):
  FD grid and stencils                  finite_differences.f90:95-198 (get_FD_grid), :200-459 (get_FD_coeffs),
                                        :498-557 (get_der_mat), :559-606 (Fornberg 1988 weights)
  radial functions of the full sphere   radial.f90:268-300 (or1..or4 = 0 at r = 0, l_R), :628 (rgrav), preCalculations.f90:
                                        :170-176 (BuoFac), :304-310 (delxh2, delxr2), :329-330 (c_moi_oc), :364 (epsc)
  get_dr / get_ddr (FD)                 radial_derivatives.f90:415-443, :521-563  (== drMat, d2rMat applied to grid values)
  rInt_R (FD)                           integration.f90:97-155 (Simpson on the irregular grid)
  restart                               readCheckPoints.f90:840-1060, :1508-1601; startFields.f90:373-405
  time scheme                           multistep_schemes.f90:195-206, :430-470, :558-590 (as oracle/lmloop.py)
  finish_explicit_assembly              LMLoop.f90:390-453, updateS.f90:543-601, updateWP.f90:1002-1031
  updateS / get_sMat / rhs_imp          updateS.f90:156-342, :1065-1140 (centre row: ds/dr = 0 for l = 0, s = 0 else), :658-756
  updateZ / get_zMat / rhs_imp          updateZ.f90:191-488, :1820-1890 (centre row z = 0), :760-955 (l_correct_AMz)
  updateWP / get_wMat / rhs_imp         updateWP.f90:255-634, :2235-2358 (double curl; centre rows dw = 0 (l = 1) or
                                        ddw = 0), :1089-1260 (third and fourth derivative are D1.D2 w and D2.D2 w there,
                                        while the matrix uses the d3rMat / d4rMat stencils -- kept as in the reference)
  get_e_kin                             kinetic_energy.f90:95-230

The implicit matrices are banded in the reference; here they are solved dense (same matrix, rounding-level differences).
"""
import numpy as np


def fd_weights(x, m):
    """Weights c[i, k] of the k-th derivative (k = 0..m) at 0 on the nodes x (Fornberg 1988; finite_differences.f90:559-606
    `populate_fd_weights` with z = 0).  The weights are the derivatives of the Lagrange basis, so any correct evaluation
    agrees to rounding."""
    x = np.asarray(x, dtype=float)
    n = len(x) - 1
    c = np.zeros((n + 1, m + 1))
    c[0, 0] = 1.0
    c1 = 1.0
    c4 = x[0]
    for i in range(1, n + 1):
        mn = min(i, m)
        c2 = 1.0
        c5 = c4
        c4 = x[i]
        for j in range(i):
            c3 = x[i] - x[j]
            c2 *= c3
            if j == i - 1:
                for k in range(mn, 0, -1):
                    c[i, k] = c1 * (k * c[i - 1, k - 1] - c5 * c[i - 1, k]) / c2
                c[i, 0] = -c1 * c5 * c[i - 1, 0] / c2
            for k in range(mn, 0, -1):
                c[j, k] = (c4 * c[j, k] - k * c[j, k - 1]) / c3
            c[j, 0] = c4 * c[j, 0] / c3
        c1 = c2
    return c


class FDSphere:
    """Finite-difference radial grid of a full sphere (r_icb = 0, r_cmb = 1), nR = 1 at the surface, nR = n_r_max at r = 0."""

    def __init__(self, n_r_max, order=4, order_boundary=2, fd_stretch=0.3, fd_ratio=0.2):
        N = n_r_max
        self.n_r_max, self.order, self.order_boundary = N, order, order_boundary
        self.r_cmb, self.r_icb = 1.0, 0.0
        # ---- get_FD_grid, full-sphere branch (finite_differences.f90:130-185): geometric refinement towards the surface only
        r = np.zeros(N)
        r[0] = self.r_cmb
        if fd_ratio == 1.0:
            r = self.r_cmb - np.arange(N) / (N - 1.0)
        else:
            n_bound = int((N - 1.0) / (2.0 * (1.0 + fd_stretch)))
            n_bulk = N - 1 - n_bound
            q = np.exp(np.log(fd_ratio) / n_bound)
            dr = 1.0
            for _ in range(n_bound):
                dr *= q
            dr = 1.0 / (n_bulk + q * ((1.0 - dr) / (1.0 - q)))      # drMax
            for _ in range(n_bound):
                dr *= q                                              # drMin
            for n in range(1, n_bound + 1):
                r[n] = r[n - 1] - dr
                dr /= q
            for n in range(n_bulk):
                r[n + n_bound + 1] = r[n + n_bound] - dr
        assert abs(r[-1]) < 1e-12
        r[-1] = self.r_icb
        self.r = r
        # ---- get_FD_coeffs + get_der_mat: dense differentiation matrices (finite_differences.f90:200-557)
        h, ob = order // 2, order_boundary
        D = [np.zeros((N, N)) for _ in range(5)]

        def put(k, row, cols, npts_m):
            D[k][row, cols] = fd_weights(r[cols] - r[row], npts_m)[:, k]

        for n in range(h, N - h):                                   # bulk, first and second derivative
            cols = np.arange(n - h, n + h + 1)
            put(1, n, cols, order)
            put(2, n, cols, order)
        for n in range(h + 1, N - h - 1):                           # bulk, third and fourth derivative
            cols = np.arange(n - h - 1, n + h + 2)
            put(3, n, cols, order + 2)
            put(4, n, cols, order + 2)
        for n in range(h):                                          # one-sided rows at both ends
            put(1, n, np.arange(ob + 1), ob)
            put(1, N - 1 - n, N - 1 - np.arange(ob + 1), ob)
            put(2, n, np.arange(ob + 2), ob + 1)
            put(2, N - 1 - n, N - 1 - np.arange(ob + 2), ob + 1)
        for n in range(h + 1):
            put(3, n, np.arange(ob + 3), ob + 2)
            put(3, N - 1 - n, N - 1 - np.arange(ob + 3), ob + 2)
            put(4, n, np.arange(ob + 4), 4)
            put(4, N - 1 - n, N - 1 - np.arange(ob + 4), 4)
        self.D1, self.D2, self.D3, self.D4 = D[1], D[2], D[3], D[4]
        # ---- radial.f90:268-283: inverse powers of r vanish at the centre
        with np.errstate(divide="ignore"):
            self.or1 = np.where(r > 0.0, 1.0 / np.where(r > 0.0, r, 1.0), 0.0)
        self.or2 = self.or1 ** 2
        self.or4 = self.or2 ** 2

    def rInt_R(self, f):
        """simps (integration.f90:105-155); r decreases with the index, hence the final sign."""
        r, N = self.r, self.n_r_max

        def simpson(idx):       # idx: 0-based centres of the three-point panels
            h2 = r[idx + 1] - r[idx]
            h1 = r[idx] - r[idx - 1]
            return np.sum((h1 + h2) / 6.0 * (f[idx - 1] * (2.0 * h1 - h2) / h1 + f[idx] * (h1 + h2) ** 2 / (h1 * h2) +
                                              f[idx + 1] * (2.0 * h2 - h1) / h2), axis=0)

        if N % 2 == 1:
            return -simpson(np.arange(1, N - 1, 2))
        tot = 0.5 * (r[1] - r[0]) * (f[1] + f[0]) + simpson(np.arange(2, N - 1, 2))
        tot = tot + 0.5 * (r[N - 1] - r[N - 2]) * (f[N - 1] + f[N - 2]) + simpson(np.arange(1, N - 1, 2))
        return -0.5 * tot


def _cc2real(c, m):
    return np.where(m == 0, 1.0, 2.0) * (c.real ** 2 + c.imag ** 2)


class FullSphereHost:
    """LM side of MagIC for samples/full_sphere: Boussinesq (rho0 = 1, beta = 0, visc = kappa = 1), hydro + heat with a
    uniform heat source (it enters through get_dsdt of the radial loop, get_td.f90:480), fixed entropy and stress-free
    surface, regularity at the centre, double-curl poloidal equation, CNAB2, l_correct_AMz, l_R(nR) from l_var_l."""

    def __init__(self, lm2l, lm2m, radial_loop, ckpt, l_correct_AMz=True, l_var_l=True):
        self.lm2l, self.lm2m = np.asarray(lm2l), np.asarray(lm2m)
        self.lm_max = len(self.lm2l)
        self.l_max = int(self.lm2l.max())
        N = int(ckpt["n_r_max"])
        self.N = N
        self.g = g = FDSphere(N, int(ckpt["fd_order"]), int(ckpt["fd_order_bound"]), float(ckpt["fd_stretch"]), float(ckpt["fd_ratio"]))
        self.radial_loop = radial_loop
        ra, pr, ek = float(ckpt["ra"]), float(ckpt["pr"]), float(ckpt["ek"])
        self.opr = 1.0 / pr
        self.BuoFac = ra / pr
        self.CorFac = 1.0 / ek
        self.epsc = float(ckpt["epsc0"]) * np.sqrt(4.0 * np.pi)
        self.rgrav = g.r / g.r_cmb                                   # g0 = g2 = 0, g1 = 1
        self.alpha = float(ckpt["alpha"])
        self.dtmax = float(ckpt["dtmax"])
        self.l_correct_AMz = l_correct_AMz
        if l_var_l:   # radial.f90:286-293 (the second assignment is the one that counts)
            self.l_R = np.minimum((1.0 + self.l_max * np.sqrt(g.r / g.r_cmb / float(ckpt["rcut_l"]))).astype(int), self.l_max)
        else:
            self.l_R = np.full(N, self.l_max)
        self.c_moi_oc = 8.0 / 3.0 * np.pi * g.rInt_R(g.r ** 4)
        self.dL = (self.lm2l * (self.lm2l + 1)).astype(float)
        self.below_lR = self.lm2l[None, :] <= self.l_R[:, None]      # the `if ( l > l_R(n_r) ) cycle` of finish_exp_*
        # ---- restart (readCheckPoints.f90): fields and the explicit terms of the previous step
        self.time = float(ckpt["time"])
        self.dt = np.array(ckpt["dt"], dtype=float)
        self.w, self.z, self.s = (np.array(ckpt[k], dtype=np.complex128) for k in ("w", "z", "s"))
        assert self.w.shape == (N, self.lm_max)
        zero = lambda: np.zeros((N, self.lm_max), dtype=np.complex128)
        self.old, self.impl, self.expl = {}, {}, {}
        for nm in ("s", "w", "z"):
            self.old[nm], self.impl[nm] = zero(), zero()
            self.expl[nm] = [zero(), np.array(ckpt["d%sdt_expl2" % nm], dtype=np.complex128)]
        self._mats = None
        self.n_steps = 0
        # startFields.f90:373-405
        self._rhs_imp_s()
        self._rhs_imp_w()
        self._rhs_imp_z()

    def _lm(self, l, m):
        return int(np.nonzero((self.lm2l == l) & (self.lm2m == m))[0][0])

    # ---- implicit terms ------------------------------------------------------------------------------
    def _rhs_imp_s(self):
        """get_entropy_rhs_imp, updateS.f90:658-756."""
        g = self.g
        self.ds = g.D1 @ self.s
        dds = g.D2 @ self.s
        self.old["s"] = self.s.copy()
        self.impl["s"] = self.opr * (dds + 2.0 * g.or1[:, None] * self.ds - self.dL[None, :] * g.or2[:, None] * self.s)

    def _rhs_imp_z(self):
        """get_tor_rhs_imp, updateZ.f90:760-955 with rho0 = 1, beta = 0, non-rotating boundaries, AMstart = 0."""
        g = self.g
        r = g.r
        self.dz = g.D1 @ self.z
        ddz = g.D2 @ self.z
        if self.l_correct_AMz:
            lm = self._lm(1, 0)
            corr = (8.0 / 3.0 * np.pi * g.rInt_R(r * r * self.z[:, lm].real)) / self.c_moi_oc
            self.z[:, lm] -= r * r * corr
            self.dz[:, lm] -= 2.0 * r * corr
            ddz[:, lm] -= 2.0 * corr
        fac = self.dL[None, :] * g.or2[:, None]
        self.old["z"] = fac * self.z
        imp = fac * (ddz - fac * self.z)
        imp[0] = 0.0
        imp[-1] = 0.0
        self.impl["z"] = imp

    def _rhs_imp_w(self):
        """get_pol_rhs_imp, double-curl branch (updateWP.f90:1146-1151, :1169-1180, :1206-1245) with beta = dLvisc = 0."""
        g = self.g
        or1, or2 = g.or1[:, None], g.or2[:, None]
        dL = self.dL[None, :]
        self.dw = g.D1 @ self.w
        self.ddw = g.D2 @ self.w
        dddw = g.D1 @ self.ddw          # get_ddr(ddw, work_LMloc, ddddw): derivatives OF ddw, not the wide stencils
        ddddw = g.D2 @ self.ddw
        old = dL * or2 * (-(self.ddw - dL * or2 * self.w))
        Dif = -dL * or2 * (ddddw + 0.0 * dddw + (-2.0 * or2 * dL) * self.ddw + (2.0 * (2.0 * or1) * or2 * dL) * self.dw +
                           dL * or2 * (2.0 * or1 * (-3.0 * or1) + dL * or2) * self.w)
        Buo = self.BuoFac * dL * or2 * self.rgrav[:, None] * self.s
        imp = Dif + Buo
        l0 = self.lm2l == 0
        for a in (old, imp):
            a[0] = 0.0
            a[-1] = 0.0
            a[:, l0] = 0.0
        self.old["w"], self.impl["w"] = old, imp

    # ---- time scheme ---------------------------------------------------------------------------------
    def _weights(self):
        dt1, dt2 = self.dt
        return 1.0, self.alpha * dt1, (1.0 - self.alpha) * dt1, (1.0 + 0.5 * dt1 / dt2) * dt1, -0.5 * dt1 * dt1 / dt2

    def _imex_rhs(self, nm, wts):
        wimp, wl1, wl2, we1, we2 = wts
        return wimp * self.old[nm] + wl2 * self.impl[nm] + we1 * self.expl[nm][0] + we2 * self.expl[nm][1]

    def _build_mats(self, wl1):
        g, N = self.g, self.N
        I = np.eye(N)
        or1, or2 = g.or1[:, None], g.or2[:, None]
        mats = {"s": [], "z": [], "w": []}
        for l in range(self.l_max + 1):
            dL = float(l * (l + 1))
            # get_sMat (updateS.f90:1086-1140): ktops = 1; centre: ds/dr = 0 for l = 0, s = 0 otherwise
            M = I - wl1 * self.opr * (g.D2 + 2.0 * or1 * g.D1 - dL * or2 * I)
            M[0] = I[0]
            M[-1] = g.D1[-1] if l == 0 else I[-1]
            mats["s"].append(M)
            # get_zMat (updateZ.f90:1850-1890): stress-free surface, z = 0 at the centre
            M = dL * or2 * I - wl1 * dL * or2 * (g.D2 - dL * or2 * I)
            M[0] = g.D1[0] - 2.0 * g.or1[0] * I[0]
            M[-1] = I[-1]
            mats["z"].append(M)
            # get_wMat (updateWP.f90:2258-2325): w = 0 at both ends; stress-free surface in row 2; centre: dw = 0 for l = 1,
            # ddw = 0 otherwise in row N-1; bulk rows 3 .. N-2
            M = -dL * or2 * (g.D2 - dL * or2 * I) + wl1 * dL * or2 * (
                g.D4 + (-2.0 * dL * or2) * g.D2 + (2.0 * (2.0 * or1) * dL * or2) * g.D1 +
                dL * or2 * (2.0 * or1 * (-3.0 * or1) + dL * or2) * I)
            M[0] = I[0]
            M[1] = g.D2[0] - 2.0 * g.or1[0] * g.D1[0]
            M[-2] = g.D1[-1] if l == 1 else g.D2[-1]
            M[-1] = I[-1]
            mats["w"].append(M)
        self._mats = (wl1, mats)

    @staticmethod
    def _solve(M, rhs):
        f = 1.0 / np.max(np.abs(M), axis=1)     # row equilibration, as WITH_PRECOND_* / wMat_fac(:,1)
        return np.linalg.solve(M * f[:, None], rhs * f[:, None])

    # ---- one time step -------------------------------------------------------------------------------
    def fields_Rloc(self):
        return dict(w=self.w, dw=self.dw, ddw=self.ddw, z=self.z, dz=self.dz, s=self.s)

    def step(self):
        """One pass of the n_time_step loop of step_time.f90:480-763 (CNAB2: one stage)."""
        g, N = self.g, self.N
        out = self.radial_loop({k: np.ascontiguousarray(v) for k, v in self.fields_Rloc().items()})
        or2 = g.or2[:, None]
        # finish_explicit_assembly: finish_exp_entropy (dentropy0 = 0, orho1 = 1), finish_exp_pol
        self.expl["s"][0] = np.where(self.below_lR, out["dsdt"] - or2 * (g.D1 @ out["dVSrLM"]), out["dsdt"])
        self.expl["w"][0] = np.where(self.below_lR & (self.lm2l != 0)[None, :], out["dwdt"] + or2 * (g.D1 @ out["dVxVhLM"]),
                                     out["dwdt"])
        self.expl["z"][0] = np.array(out["dzdt"])
        # dt_courant (courant.f90:277-346); the centre level takes no part (rIter.f90:295)
        self.dtrkc_min, self.dthkc_min = float(np.min(out["dtrkc"][:-1])), float(np.min(out["dthkc"][:-1]))
        if self.dt[0] > min(self.dtrkc_min, self.dthkc_min):
            raise RuntimeError("Courant criterion asks for a smaller time step; not expected in this sample")
        self.dt = np.array([self.dt[0], self.dt[0]])
        wts = self._weights()
        wl1 = wts[1]
        if self._mats is None or self._mats[0] != wl1:
            self._build_mats(wl1)
        mats = self._mats[1]
        self.time += self.dt[0]
        m0 = self.lm2m == 0

        def rotate(nm):
            self.expl[nm][1] = self.expl[nm][0]

        # ---- updateS: tops = bots = 0 (ktops = 1, kbots = 2 without imposed s_top / s_bot)
        rhs = self._imex_rhs("s", wts)
        rhs[0], rhs[-1] = 0.0, 0.0
        for l in range(self.l_max + 1):
            idx = np.nonzero(self.lm2l == l)[0]
            self.s[:, idx] = self._solve(mats["s"][l], rhs[:, idx])
        self.s[:, m0] = self.s[:, m0].real
        rotate("s")
        self._rhs_imp_s()
        # ---- updateZ
        rhs = self._imex_rhs("z", wts)
        rhs[0], rhs[-1] = 0.0, 0.0
        for l in range(self.l_max + 1):
            idx = np.nonzero(self.lm2l == l)[0]
            self.z[:, idx] = 0.0 if l == 0 else self._solve(mats["z"][l], rhs[:, idx])
        self.z[:, m0] = self.z[:, m0].real
        rotate("z")
        self._rhs_imp_z()
        # ---- updateWP, double curl: rows 1, 2, N-1, N are boundary conditions; the buoyancy of the NEW entropy is implicit
        rhs = self._imex_rhs("w", wts) + wl1 * self.dL[None, :] * or2 * self.BuoFac * self.rgrav[:, None] * self.s
        rhs[:2] = 0.0
        rhs[-2:] = 0.0
        for l in range(self.l_max + 1):
            idx = np.nonzero(self.lm2l == l)[0]
            self.w[:, idx] = 0.0 if l == 0 else self._solve(mats["w"][l], rhs[:, idx])
        self.w[:, m0] = self.w[:, m0].real
        rotate("w")
        self._rhs_imp_w()
        self.n_steps += 1

    # ---- diagnostics ---------------------------------------------------------------------------------
    def e_kin(self):
        """Columns 2-9 of e_kin.TAG (kinetic_energy.f90:126-196)."""
        g = self.g
        m = self.lm2m[None, :]
        dL = self.dL[None, :]
        e_p = dL * (dL * g.or2[:, None] * _cc2real(self.w, m) + _cc2real(self.dw, m))
        e_t = dL * _cc2real(self.z, m)
        l, mm = self.lm2l, self.lm2m
        axi = mm == 0
        es = (l + mm) % 2 == 0
        eas_p = axi & (l % 2 == 0)
        eas_t = axi & (l % 2 != 0)
        cols = [e_p.sum(1), e_t.sum(1), e_p[:, axi].sum(1), e_t[:, axi].sum(1), e_p[:, es].sum(1), e_t[:, ~es].sum(1),
                e_p[:, eas_p].sum(1), e_t[:, eas_t].sum(1)]
        return np.array([0.5 * float(g.rInt_R(c)) for c in cols])
