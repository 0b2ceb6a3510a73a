"""numpy restatement of the reference's HOST side of one time step, for one configuration family.  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference pins the radial-loop hot path (SHT, get_nl, get_td, courant) only END TO END, through the
energy time series of its sample runs (samples/*/reference.out, compared at rtol 1e-8 by samples/*/unitTest.py).  The
Fortran host cannot be built in this image (no Fortran compiler, no MPI), so this module restates the part of the host
that surrounds the radial loop -- start fields, `finish_explicit_assembly`, the CN/AB2 right-hand sides, the implicit
solves of `LMLoop` and the energy diagnostics -- for the Boussinesq / Chebyshev / CNAB2 / "WP" path that
`samples/dynamo_benchmark` runs.  The radial loop itself is a callable handed in by the test: the CPU oracle or the CUDA
library through the C ABI.  With either, 100 steps must reproduce reference.out / referenceMag.out.

It is NOT on the product path (magic_b200/ never imports oracle/); the product keeps the Fortran host.

Restated reference routines (file:line relative to /root/reference/src):
  radial grid, Chebyshev matrices      radial.f90 (r, or1, or2, rgrav :628), chebyshev.f90 (rMat, drMat, d2rMat, d3rMat)
  start fields                         startFields.f90:281-330, init_fields.f90:373-560 (initS, init_s1=llmm),
                                       :1129-1188 (initB, init_b1=3), :2211-2497 (ps_cond, entropy diffusion branch)
  time scheme                          multistep_schemes.f90:195-206 (CNAB2 weights), :430-470 (set_imex_rhs),
                                       :558-590 (rotate_imex)
  finish_explicit_assembly             LMLoop.f90:390-453, updateS.f90:543-601, updateB.f90:1005-1041
  updateS / get_sMat / rhs_imp         updateS.f90:156-342, :1065-1212, :658-756
  updateZ / get_zMat / rhs_imp         updateZ.f90:191-488, :1820-1964, :760-1025
  updateWP / get_wpMat / rhs_imp       updateWP.f90:255-634, :1978-2171, :1089-1336
  updateB / get_bMat / rhs_imp         updateB.f90:226-694, :1785-2171, :1520-1652
  dt_courant                           courant.f90:277-346
  get_e_kin / get_e_mag                kinetic_energy.f90:95-230, magnetic_energy.f90:262-470
  step order                           step_time.f90:480-763

The reference solves for Chebyshev coefficients with collocation matrices rMat/drMat/... and transforms back
(costf1); with n_cheb_max = n_r_max this is algebraically the same as solving for the grid values with the
differentiation matrices D_k = d^kT . T^-1 used here (differences are rounding only).
"""
import numpy as np


class ChebShell:
    """Gauss-Lobatto radial grid of a spherical shell, nR=1 at the CMB (radial.f90, chebyshev.f90, no mapping)."""

    def __init__(self, n_r_max, radratio):
        N = n_r_max
        self.n_r_max = N
        self.r_cmb = 1.0 / (1.0 - radratio)
        self.r_icb = self.r_cmb - 1.0
        k = np.arange(N)
        x = np.cos(np.pi * k / (N - 1))
        self.x = x
        self.r = 0.5 * (self.r_cmb - self.r_icb) * x + 0.5 * (self.r_cmb + self.r_icb)
        self.r[0], self.r[-1] = self.r_cmb, self.r_icb
        drx = 2.0 / (self.r_cmb - self.r_icb)
        T = np.zeros((N, N))
        d1 = np.zeros((N, N))
        d2 = np.zeros((N, N))
        d3 = np.zeros((N, N))
        T[:, 0] = 1.0
        T[:, 1] = x
        d1[:, 1] = 1.0
        for n in range(1, N - 1):  # chebyshev.f90 get_chebs recurrences
            T[:, n + 1] = 2 * x * T[:, n] - T[:, n - 1]
            d1[:, n + 1] = 2 * T[:, n] + 2 * x * d1[:, n] - d1[:, n - 1]
            d2[:, n + 1] = 4 * d1[:, n] + 2 * x * d2[:, n] - d2[:, n - 1]
            d3[:, n + 1] = 6 * d2[:, n] + 2 * x * d3[:, n] - d3[:, n - 1]
        Tinv = np.linalg.inv(T)
        self.D1 = (d1 * drx) @ Tinv
        self.D2 = (d2 * drx ** 2) @ Tinv
        self.D3 = (d3 * drx ** 3) @ Tinv
        self.T, self.Tinv = T, Tinv
        self.or1 = 1.0 / self.r
        self.or2 = self.or1 ** 2
        # rInt_R (integration.f90): exact integral of the Chebyshev interpolant
        n = np.arange(N)
        with np.errstate(divide="ignore"):
            wn = np.where(n % 2 == 0, 2.0 / (1.0 - n.astype(float) ** 2), 0.0)
        self.w_int = 0.5 * (self.r_cmb - self.r_icb) * (wn @ Tinv)

    def rInt_R(self, f):
        return float(self.w_int @ f)


def _cc2real(c, m):
    """useful.f90 cc2real: |c|^2 (m=0) or 2|c|^2."""
    return np.where(m == 0, 1.0, 2.0) * (c.real ** 2 + c.imag ** 2)


class BoussinesqDynamoHost:
    """The LM-side of MagIC for a Boussinesq MHD shell with rigid walls, fixed entropy, insulating boundaries,
    Chebyshev collocation and CN/AB2 -- the setup of samples/dynamo_benchmark/input.nml."""

    def __init__(self, lm2l, lm2m, radial_loop, n_r_max=33, radratio=0.35, ra=1e5, ek=1e-3, pr=1.0, prmag=5.0,
                 dtmax=1e-4, alpha=0.6, init_s1=404, amp_s1=0.1, init_b1=3, amp_b1=5.0):
        self.lm2l = np.asarray(lm2l)
        self.lm2m = np.asarray(lm2m)
        self.lm_max = len(self.lm2l)
        self.l_max = int(self.lm2l.max())
        self.g = g = ChebShell(n_r_max, radratio)
        self.N = n_r_max
        self.radial_loop = radial_loop
        self.opr, self.opm = 1.0 / pr, 1.0 / prmag
        self.BuoFac = ra / pr                 # preCalculations.f90:170
        self.LFfac = 1.0 / (ek * prmag)       # preCalculations.f90:158
        self.rgrav = g.r / g.r_cmb            # radial.f90:628 with g1=1
        self.alpha = alpha
        self.dtmax = dtmax
        self.dt = np.array([dtmax, dtmax])    # startFields.f90:301
        self.time = 0.0
        self.n_steps = 0
        self.dL = (self.lm2l * (self.lm2l + 1)).astype(float)
        sq4pi = np.sqrt(4.0 * np.pi)
        N, lm_max = self.N, self.lm_max
        z = lambda: np.zeros((N, lm_max), dtype=np.complex128)
        self.w, self.z, self.p, self.s, self.b, self.aj = z(), z(), z(), z(), z(), z()
        self.tops = np.zeros(lm_max, dtype=np.complex128)
        self.bots = np.zeros(lm_max, dtype=np.complex128)
        lm00 = self._lm(0, 0)
        self.bots[lm00] = sq4pi               # preCalculations.f90:415-418
        # ---- initS: conductive state (ps_cond, entropy diffusion, epsc=0) + one mode (init_fields.f90:428-541)
        M = g.D2 + 2.0 * g.or1[:, None] * g.D1
        M[0] = 0.0
        M[0, 0] = 1.0
        M[-1] = 0.0
        M[-1, -1] = 1.0
        rhs = np.zeros(N)
        rhs[0], rhs[-1] = self.tops[lm00].real, self.bots[lm00].real
        self.s[:, lm00] = np.linalg.solve(M, rhs)
        if init_s1 >= 100:
            l, m = init_s1 // 100, init_s1 % 100
            x = 2.0 * g.r - g.r_cmb - g.r_icb
            s1 = 1.0 - 3.0 * x ** 2 + 3.0 * x ** 4 - x ** 6
            self.s[:, self._lm(l, m)] += amp_s1 * s1
        # ---- initB, init_b1=3, insulating inner core (init_fields.f90:1129-1188)
        if init_b1 == 3:
            b_pol = amp_b1 * np.sqrt(3.0 * np.pi) / 4.0
            b_tor = -4.0 / 3.0 * amp_b1 * np.sqrt(np.pi / 5.0)
            self.b[:, self._lm(1, 0)] += b_pol * (g.r ** 3 - 4.0 / 3.0 * g.r_cmb * g.r ** 2 + g.r_icb ** 4 / 3.0 * g.or1)
            self.aj[:, self._lm(2, 0)] += b_tor * g.r * np.sin(np.pi * (g.r - g.r_icb))
        # ---- time arrays (time_array.f90): old, impl (one level each for CNAB2), expl (two levels)
        self.old, self.impl, self.expl = {}, {}, {}
        for nm in ("s", "w", "p", "z", "b", "j"):
            self.old[nm], self.impl[nm] = z(), z()
            self.expl[nm] = [z(), z()]
        self._mats = None
        # startFields.f90:373-432: derivatives and old/implicit terms of the start fields
        self._rhs_imp_s()
        self._rhs_imp_wp()
        self._rhs_imp_z()
        self._rhs_imp_b()

    # ------------------------------------------------------------------------------------------------
    def _lm(self, l, m):
        return int(np.nonzero((self.lm2l == l) & (self.lm2m == m))[0][0])

    def _d(self, D, f):
        return D @ f

    def _rhs_imp_s(self):
        """get_entropy_rhs_imp, updateS.f90:658-756 (Boussinesq: beta=dLtemp0=dLkappa=0, kappa=1)."""
        g = self.g
        self.ds = g.D1 @ self.s
        dds = g.D2 @ self.s
        self.old["s"] = self.s.copy()
        self.impl["s"] = self.opr * (dds + 2.0 * g.or1[:, None] * self.ds - self.dL[None, :] * g.or2[:, None] * self.s)

    def _rhs_imp_z(self):
        """get_tor_rhs_imp, updateZ.f90:760-1025 (visc=1, beta=0, no rotating walls)."""
        g = self.g
        self.dz = g.D1 @ self.z
        ddz = g.D2 @ self.z
        fac = self.dL[None, :] * g.or2[:, None]
        self.old["z"] = fac * self.z
        imp = fac * (ddz - fac * self.z)
        imp[0] = 0.0
        imp[-1] = 0.0      # n_r_top=n_r_cmb+1 .. n_r_bot=n_r_icb-1
        self.impl["z"] = imp

    def _rhs_imp_wp(self):
        """get_pol_rhs_imp, updateWP.f90:1089-1336, non double-curl branch."""
        g = self.g
        self.dw = g.D1 @ self.w
        self.ddw = g.D2 @ self.w
        dddw = g.D3 @ self.w
        self.dp = g.D1 @ self.p
        fac = self.dL[None, :] * g.or2[:, None]
        old_w = fac * self.w
        old_p = -fac * self.dw
        Dif = fac * (self.ddw - fac * self.w)
        Pre = -self.dp
        Buo = self.BuoFac * self.rgrav[:, None] * self.s
        imp_w = Pre + Dif + Buo
        imp_p = fac * self.p + fac * (-dddw + fac * self.dw - fac * 2.0 * g.or1[:, None] * self.w)
        l0 = self.lm2l == 0
        for a in (old_w, old_p, imp_w, imp_p):
            a[0] = 0.0
            a[-1] = 0.0
            a[:, l0] = 0.0   # lmStart_00
        self.old["w"], self.old["p"], self.impl["w"], self.impl["p"] = old_w, old_p, imp_w, imp_p

    def _rhs_imp_b(self):
        """get_mag_rhs_imp, updateB.f90:1520-1652 (lambda=1, dLlambda=0)."""
        g = self.g
        self.db = g.D1 @ self.b
        self.ddb = g.D2 @ self.b
        self.dj = g.D1 @ self.aj
        ddj = g.D2 @ self.aj
        fac = self.dL[None, :] * g.or2[:, None]
        self.old["b"] = fac * self.b
        self.old["j"] = fac * self.aj
        ib = self.opm * fac * (self.ddb - fac * self.b)
        ij = self.opm * fac * (ddj - fac * self.aj)
        for a in (ib, ij):
            a[0] = 0.0
            a[-1] = 0.0
        self.impl["b"], self.impl["j"] = ib, ij

    # ------------------------------------------------------------------------------------------------
    def _weights(self):
        """multistep_schemes.f90:195-206."""
        dt1, dt2 = self.dt
        wimp = 1.0
        wl1 = self.alpha * dt1
        wl2 = (1.0 - self.alpha) * dt1
        we1 = (1.0 + 0.5 * dt1 / dt2) * dt1
        we2 = -0.5 * dt1 * dt1 / dt2
        return wimp, wl1, wl2, we1, we2

    def _imex_rhs(self, nm, wts):
        wimp, wl1, wl2, we1, we2 = wts
        return wimp * self.old[nm] + wl2 * self.impl[nm] + we1 * self.expl[nm][0] + we2 * self.expl[nm][1]

    def _build_mats(self, wl1):
        """get_sMat / get_zMat / get_wpMat / get_bMat for every degree (LU by numpy at solve time)."""
        g, N = self.g, self.N
        I = np.eye(N)
        mats = {"s": [], "z": [], "wp": [], "b": [], "j": []}
        for l in range(self.l_max + 1):
            dL = float(l * (l + 1))
            or1, or2 = g.or1[:, None], g.or2[:, None]
            # sMat (updateS.f90:1086-1140), ktops=kbots=1
            M = I - wl1 * self.opr * (g.D2 + 2.0 * or1 * g.D1 - dL * or2 * I)
            M[0], M[-1] = I[0], I[-1]
            mats["s"].append(M)
            # zMat (updateZ.f90:1850-1890), no slip
            M = dL * or2 * I - wl1 * dL * or2 * (g.D2 - dL * or2 * I)
            M[0], M[-1] = I[0], I[-1]
            mats["z"].append(M)
            # wpMat (updateWP.f90:1999-2090), no slip
            W = np.zeros((2 * N, 2 * N))
            W[:N, :N] = dL * or2 * I - wl1 * dL * or2 * (g.D2 - dL * or2 * I)
            W[:N, N:] = wl1 * g.D1
            W[N:, :N] = -dL * or2 * g.D1 - wl1 * dL * or2 * (-g.D3 + dL * or2 * g.D1 - dL * or2 * 2.0 * or1 * I)
            W[N:, N:] = -wl1 * dL * or2 * I
            W[0] = 0.0
            W[0, :N] = I[0]
            W[N - 1] = 0.0
            W[N - 1, :N] = I[-1]
            W[N] = 0.0
            W[N, :N] = g.D1[0]
            W[2 * N - 1] = 0.0
            W[2 * N - 1, :N] = g.D1[-1]
            mats["wp"].append(W)
            # bMat / jMat (updateB.f90:1824-1905), ktopb=kbotb=1, conductance_ma=0
            B = dL * or2 * I - wl1 * self.opm * dL * or2 * (g.D2 - dL * or2 * I)
            J = B.copy()
            B[0] = g.D1[0] + l * g.or1[0] * I[0]
            B[-1] = g.D1[-1] - (l + 1.0) * g.or1[-1] * I[-1]
            J[0], J[-1] = I[0], I[-1]
            mats["b"].append(B)
            mats["j"].append(J)
        self._mats = (wl1, mats)

    @staticmethod
    def _solve(M, rhs):
        # row equilibration like WITH_PRECOND_* (conditioning only)
        f = 1.0 / np.max(np.abs(M), axis=1)
        return np.linalg.solve(M * f[:, None], rhs * f[:, None])

    # ------------------------------------------------------------------------------------------------
    def fields_Rloc(self):
        """What transp_LMloc_to_Rloc hands to the radial loop (step_time.f90:1005-1132)."""
        return dict(w=self.w, dw=self.dw, ddw=self.ddw, z=self.z, dz=self.dz, s=self.s, b=self.b, db=self.db,
                    ddb=self.ddb, aj=self.aj, dj=self.dj)

    def step(self):
        """One pass of the n_time_step loop of step_time.f90:480-763 (CNAB2: one stage)."""
        g, N = self.g, self.N
        out = self.radial_loop({k: np.ascontiguousarray(v) for k, v in self.fields_Rloc().items()})
        or2 = g.or2[:, None]
        l0 = (self.lm2l == 0)[None, :]
        # finish_explicit_assembly (LMLoop.f90:390-453): orho1=1, dentropy0=0
        self.expl["s"][0] = out["dsdt"] - or2 * (g.D1 @ out["dVSrLM"])                 # updateS.f90:587-597
        self.expl["w"][0] = np.array(out["dwdt"])
        self.expl["p"][0] = np.array(out["dpdt"])
        self.expl["z"][0] = np.array(out["dzdt"])
        self.expl["b"][0] = np.array(out["dbdt"])
        self.expl["j"][0] = out["djdt"] + np.where(l0, 0.0, or2 * (g.D1 @ out["dVxBhLM"]))  # updateB.f90:1030-1037
        # dt_courant (courant.f90:277-346)
        self.dtrkc_min, self.dthkc_min = float(np.min(out["dtrkc"])), float(np.min(out["dthkc"]))
        dt_min = min(self.dtrkc_min, self.dthkc_min, 1000.0 * self.dtmax)
        if self.dt[0] > dt_min:
            raise RuntimeError("Courant criterion asks for a smaller time step; not expected in this sample")
        self.dt = np.array([self.dt[0], self.dt[0]])                                    # set_dt_array
        wts = self._weights()
        wimp, wl1, wl2, we1, we2 = wts
        if self._mats is None or self._mats[0] != wl1:
            self._build_mats(wl1)
        mats = self._mats[1]
        self.time += self.dt[0]
        m0 = self.lm2m == 0

        def per_degree(fn):
            for l in range(self.l_max + 1):
                idx = np.nonzero(self.lm2l == l)[0]
                fn(l, idx)

        def rotate(nm):
            self.expl[nm][1] = self.expl[nm][0]

        # ---- updateS (updateS.f90:156-342)
        rhs = self._imex_rhs("s", wts)
        rhs[0], rhs[-1] = self.tops, self.bots

        def up_s(l, idx):
            self.s[:, idx] = self._solve(mats["s"][l], rhs[:, idx])
        per_degree(up_s)
        self.s[:, m0] = self.s[:, m0].real
        rotate("s")
        self._rhs_imp_s()
        # ---- updateZ (updateZ.f90:191-488)
        rhs = self._imex_rhs("z", wts)
        rhs[0], rhs[-1] = 0.0, 0.0

        def up_z(l, idx):
            if l == 0:
                self.z[:, idx] = 0.0
            else:
                self.z[:, idx] = self._solve(mats["z"][l], rhs[:, idx])
        per_degree(up_z)
        self.z[:, m0] = self.z[:, m0].real
        rotate("z")
        self._rhs_imp_z()
        # ---- updateWP (updateWP.f90:255-634), buoyancy of the NEW entropy is implicit (:514-524)
        rw = self._imex_rhs("w", wts) + wl1 * self.BuoFac * self.rgrav[:, None] * self.s
        rp = self._imex_rhs("p", wts)
        for a in (rw, rp):
            a[0], a[-1] = 0.0, 0.0

        def up_wp(l, idx):
            if l == 0:
                self.w[:, idx] = 0.0     # p(l=0) (get_p0Mat) does not feed back into the flow; left untouched
                return
            sol = self._solve(mats["wp"][l], np.concatenate([rw[:, idx], rp[:, idx]], axis=0))
            self.w[:, idx] = sol[:N]
            self.p[:, idx] = sol[N:]
        per_degree(up_wp)
        self.w[:, m0] = self.w[:, m0].real
        self.p[:, m0] = self.p[:, m0].real
        rotate("w")
        rotate("p")
        self._rhs_imp_wp()
        # ---- updateB (updateB.f90:226-694)
        rb = self._imex_rhs("b", wts)
        rj = self._imex_rhs("j", wts)
        for a in (rb, rj):
            a[0], a[-1] = 0.0, 0.0

        def up_b(l, idx):
            if l == 0:
                self.b[:, idx] = 0.0
                self.aj[:, idx] = 0.0
                return
            self.b[:, idx] = self._solve(mats["b"][l], rb[:, idx])
            self.aj[:, idx] = self._solve(mats["j"][l], rj[:, idx])
        per_degree(up_b)
        self.b[:, m0] = self.b[:, m0].real
        self.aj[:, m0] = self.aj[:, m0].real
        rotate("b")
        rotate("j")
        self._rhs_imp_b()
        self.n_steps += 1

    # ------------------------------------------------------------------------------------------------
    def e_kin(self):
        """Columns 2-9 of e_kin.TAG (kinetic_energy.f90:126-196): e_p, e_t, e_p_as, e_t_as, e_p_es, e_t_es, e_p_eas,
        e_t_eas."""
        g = self.g
        l, m = self.lm2l[None, :], self.lm2m[None, :]
        dL = self.dL[None, :]
        e_p = dL * (dL * g.or2[:, None] * _cc2real(self.w, m) + _cc2real(self.dw, m))
        e_t = dL * _cc2real(self.z, m)
        return self._energy_columns(e_p, e_t, es_parity=0, eas_parity=0, fac=0.5)

    def e_mag_oc(self):
        """Columns 2-13 of e_mag_oc.TAG (magnetic_energy.f90:262-300, 447-470, 570-600): e_p, e_t, e_p_as, e_t_as,
        e_p_os, e_p_as_os (potential field outside the CMB), e_p_es, e_t_es, e_p_eas, e_t_eas (note the reference's
        parity convention for the magnetic field), e_p_e, e_p_as_e (external field, zero without n_imp)."""
        g = self.g
        m = self.lm2m[None, :]
        dL = self.dL[None, :]
        e_p = dL * (dL * g.or2[:, None] * _cc2real(self.b, m) + _cc2real(self.db, m))
        e_t = dL * _cc2real(self.aj, m)
        c = self._energy_columns(e_p, e_t, es_parity=1, eas_parity=1, fac=0.5 * self.LFfac)
        l = self.lm2l
        os_lm = (l * l * (l + 1.0)) * _cc2real(self.b[0], self.lm2m)
        fac = 0.5 * self.LFfac / g.r_cmb
        e_p_os, e_p_as_os = fac * os_lm.sum(), fac * os_lm[self.lm2m == 0].sum()
        return np.concatenate([c[:4], [e_p_os, e_p_as_os], c[4:], [0.0, 0.0]])

    def _energy_columns(self, e_p, e_t, es_parity, eas_parity, fac):
        l, m = self.lm2l, self.lm2m
        axi = m == 0
        es_p = (l + m) % 2 == es_parity
        eas_p = axi & (l % 2 == eas_parity)
        eas_t = axi & (l % 2 != eas_parity)
        I = self.g.rInt_R
        cols = [e_p.sum(1), e_t.sum(1), e_p[:, axi].sum(1), e_t[:, axi].sum(1), e_p[:, es_p].sum(1),
                e_t[:, ~es_p].sum(1), e_p[:, eas_p].sum(1), e_t[:, eas_t].sum(1)]
        return np.array([fac * I(c) for c in cols])
