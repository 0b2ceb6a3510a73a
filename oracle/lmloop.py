"""numpy restatement of the reference's HOST side of one time step, for one configuration family.  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference pins the radial-loop hot path (SHT, get_nl, get_td, courant) only END TO END, through the
energy time series of its sample runs (samples/*/reference.out, compared at rtol 1e-8 by samples/*/unitTest.py).  The
Fortran host cannot be built in this image (no Fortran compiler, no MPI), so this module restates the part of the host
that surrounds the radial loop -- start fields, `finish_explicit_assembly`, the CN/AB2 right-hand sides, the implicit
solves of `LMLoop` and the energy diagnostics -- for the Chebyshev / CNAB2 / "WP" (pressure formulation, separate
matrices) path that `samples/dynamo_benchmark` (Boussinesq MHD, rigid insulating walls) and `samples/hydro_bench_anel`
(anelastic polytropic hydro, stress-free walls, angular-momentum correction, n_cheb_max < n_r_max) run.  The radial loop
itself is a callable handed in by the test: the CPU oracle or the CUDA library through the C ABI.  With either, the
energy series of reference.out / referenceMag.out must be reproduced at the autotest tolerance.

Grown since, one reference sample at a time, and always the same way (a switch of ShellHost, the routines it restates cited
below): samples/precession (l_heat off, Poincare force), samples/dynamo_benchmark_condICrotIC (conducting inner core on an
even-Chebyshev grid, coupled outer/inner-core matrices, freely rotating inner core, restart with other boundary conditions,
nonlinear magnetic boundary condition at the ICB), samples/varCond (anelastic MHD, variable conductivity),
samples/doubleDiffusion and samples/boussBenchSat (composition equation, start from a checkpoint, IMEX Runge-Kutta scheme
BPR353 in DirkShellHost).  The finite-difference full sphere of samples/full_sphere lives in oracle/lmloop_fd.py.

It is NOT on the product path (magic_b200/ never imports oracle/); the product keeps the Fortran host.

Restated reference routines (file:line relative to /root/reference/src):
  radial grid, Chebyshev matrices      radial.f90 (r, or1, or2, rgrav :628), chebyshev.f90 (rMat, drMat, d2rMat, d3rMat)
  polytropic reference state           radial.f90:625-729 (adiabatic branch), :759-767 (ViscHeatFac), preCalculations.f90:313-330
  start fields                         startFields.f90:281-330, init_fields.f90:373-560 (initS, init_s1=llmm),
                                       :1129-1188 (initB, init_b1=3), :2211-2497 (ps_cond, entropy diffusion branch)
  time scheme                          multistep_schemes.f90:195-206 (CNAB2 weights), :430-470 (set_imex_rhs),
                                       :558-590 (rotate_imex)
  finish_explicit_assembly             LMLoop.f90:390-453, updateS.f90:543-601, updateB.f90:1005-1041
  updateS / get_sMat / rhs_imp         updateS.f90:156-342, :1065-1212, :658-756
  updateZ / get_zMat / rhs_imp         updateZ.f90:191-488, :1820-1964, :760-1025 (incl. l_correct_AMz / l_correct_AMe :840-915,
                                       get_angular_moment outRot.f90:485-543)
  updateWP / get_wpMat / rhs_imp       updateWP.f90:255-634, :1978-2171, :1089-1336
  updateB / get_bMat / rhs_imp         updateB.f90:226-694, :1785-2171, :1520-1652
  dt_courant                           courant.f90:277-346
  get_e_kin / get_e_mag                kinetic_energy.f90:95-230, magnetic_energy.f90:262-470
  step order                           step_time.f90:480-763
  precession                           updateZ.f90:231-235, :385-392, :955-958; preCalculations.f90:164-166
  conducting inner core                radial.f90:798-868, chebyshev_polynoms.f90 get_chebs_even, radial_derivatives_even.f90:16-70,
                                       init_fields.f90:1140-1176, updateB.f90:529-548 (rhs), :1955-2060 (get_bMat), :1077-1187
                                       (get_mag_ic_rhs_imp), :955-1003 (finish_exp_mag_ic)
  rotating inner core                  updateZ.f90:300-356 (z10 rhs), :1719-1800 (get_z10Mat), :996-1008, :1563-1619
                                       (update_rot_rates), :1659-1688 (finish_exp_tor); preCalculations.f90:313-345
  nonlinear magnetic BC at the ICB     Namelists.f90:713-720, nonlinear_bcs.f90:71-118 (get_b_nl_bcs), updateB.f90:534-539
  variable conductivity                radial.f90:903-916 (nVarCond = 2), updateB.f90:1633-1637, :1826-1837
  composition                          updateXI.f90:495-511, :579-650, :924-965; preCalculations.f90:181-185, :613-617
  restart                              readCheckPoints.f90:767-1468, startFields.f90:257-279, :373-432
  IMEX Runge-Kutta (BPR353)            dirk_schemes.f90:723-742, :764-788, :856-895, :1075-1085; step_time.f90:396-763
  get_dr of raw explicit terms         radial_derivatives.f90 get_dcheb (modes below n_cheb_max only)

The reference solves for Chebyshev coefficients with collocation matrices rMat/drMat/... and transforms back (costf1).
Here the operators are written for grid values with the differentiation matrices D_k = d^kT . T^-1 and mapped to
coefficient space by one product with the basis matrix (M_coef = M_grid . B0), which is algebraically the same matrix;
the reference's dealiasing -- boundary rows ignore the modes >= n_cheb_max and those modes are zeroed after the solve --
is applied in coefficient space exactly as in get_sMat / updateS (differences are rounding only).
"""
import numpy as np


class ChebShell:
    """Gauss-Lobatto radial grid of a spherical shell, nR=1 at the CMB (radial.f90, chebyshev.f90, no mapping)."""

    def __init__(self, n_r_max, radratio, n_cheb_max=None):
        N = n_r_max
        self.n_r_max = N
        self.n_cheb_max = N if n_cheb_max is None else n_cheb_max
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
        self._lu = {}
        # get_dr on raw (not dealiased) data, e.g. the explicit terms of finish_exp_*: the reference differentiates the modes
        # below n_cheb_max only (radial_derivatives.f90 get_dcheb with r_scheme%n_max)
        keep = (np.arange(N) < self.n_cheb_max).astype(float)
        self.D1t = ((d1 * drx) * keep[None, :]) @ Tinv
        self.D2t = ((d2 * drx ** 2) * keep[None, :]) @ Tinv   # get_ddr: second derivative of the same truncated series
        # basis of the reference's coefficient space: f = B0 c with rnorm = sqrt(2/(N-1)) and boundary_fac = 1/2 on the
        # first and last mode (chebyshev.f90; the same convention as costf1)
        wts = np.ones(N)
        wts[0] = wts[-1] = 0.5
        self.B0 = T * wts[None, :] * np.sqrt(2.0 / (N - 1))
        self.or1 = 1.0 / self.r
        self.or2 = self.or1 ** 2
        # rInt_R (integration.f90): exact integral of the Chebyshev interpolant
        n = np.arange(N)
        with np.errstate(divide="ignore"):
            wn = np.where(n % 2 == 0, 2.0 / (1.0 - n.astype(float) ** 2), 0.0)
        self.w_int = 0.5 * (self.r_cmb - self.r_icb) * (wn @ Tinv)

    def rInt_R(self, f):
        return self.w_int @ f

    def solve(self, M, rhs, bc_rows, key=None):
        """Solve like the reference: unknowns = Chebyshev coefficients, boundary rows blind to modes >= n_cheb_max, those
        modes zeroed after the solve (e.g. get_sMat updateS.f90:1107-1112 and updateS :311-318); returns grid values.
        M acts on grid values of nblk stacked fields ([nblk*N, nblk*N])."""
        N, nc = self.n_r_max, self.n_cheb_max
        nblk = M.shape[0] // N
        fac = self._lu.get(key) if key is not None else None      # coefficient-space matrix, built once per matrix
        if fac is None:
            Mc = np.empty_like(M)
            for b in range(nblk):
                Mc[:, b * N:(b + 1) * N] = M[:, b * N:(b + 1) * N] @ self.B0
            if nc < N:
                for row in bc_rows:
                    for b in range(nblk):
                        Mc[row, b * N + nc:(b + 1) * N] = 0.0
            f = 1.0 / np.max(np.abs(Mc), axis=1)  # row equilibration like WITH_PRECOND_* (conditioning only)
            fac = (Mc * f[:, None], f)
            if key is not None:
                self._lu[key] = fac
        c = np.linalg.solve(fac[0], rhs * fac[1][:, None])
        out = np.empty_like(c)
        for b in range(nblk):
            cb = c[b * N:(b + 1) * N].copy()
            cb[nc:] = 0.0
            out[b * N:(b + 1) * N] = self.B0 @ cb
        return out


class ChebEvenIC:
    """Inner-core radial grid r_icb .. 0 and the even Chebyshev basis on it (radial.f90:798-822, chebyshev_polynoms.f90
    `get_chebs_even`): the n_r_ic_max first extrema of T_(2 n_r_ic_max - 2) mapped to [-r_icb, r_icb], basis functions
    T_0, T_2, ..., derivatives with respect to r."""

    def __init__(self, n_r_ic_max, n_cheb_ic_max, r_icb):
        Ni = n_r_ic_max
        self.n_r_ic_max, self.n_cheb_ic_max = Ni, n_cheb_ic_max
        y = np.cos(np.pi * np.arange(Ni) / (2.0 * Ni - 2.0))
        self.r = r_icb * y
        self.r[-1] = 0.0
        with np.errstate(divide="ignore"):
            self.O_r = np.where(self.r > 0.0, 1.0 / np.where(self.r > 0.0, self.r, 1.0), 0.0)
        map_fac = 1.0 / r_icb
        cheb, dcheb, d2cheb = (np.zeros((Ni, Ni)) for _ in range(3))     # [mode, point]
        cheb[0] = 1.0
        last, dlast, d2last = y.copy(), np.full(Ni, map_fac), np.zeros(Ni)   # the odd polynomial in between
        for n in range(1, Ni):
            cheb[n] = 2 * y * last - cheb[n - 1]
            dcheb[n] = 2 * map_fac * last + 2 * y * dlast - dcheb[n - 1]
            d2cheb[n] = 4 * map_fac * dlast + 2 * y * d2last - d2cheb[n - 1]
            last, dlast, d2last = (2 * y * cheb[n] - last, 2 * map_fac * cheb[n] + 2 * y * dcheb[n] - dlast,
                                   4 * map_fac * dcheb[n] + 2 * y * d2cheb[n] - d2last)
        self.cheb, self.dcheb, self.d2cheb = cheb, dcheb, d2cheb
        self.cheb_norm = np.sqrt(2.0 / (Ni - 1))
        w = np.ones(Ni)
        w[0] = w[-1] = 0.5          # costf1 convention: first and last mode carry 1/2
        self.w = w
        self.B0 = (cheb * w[:, None]).T * self.cheb_norm                  # grid values = B0 . coefficients
        keep = (np.arange(Ni) < n_cheb_ic_max).astype(float)
        B0inv = np.linalg.inv(self.B0)
        # get_ddr_even (radial_derivatives_even.f90:16-70): derivatives from the modes below n_cheb_ic_max
        self.D1 = ((dcheb * (w * keep)[:, None]).T * self.cheb_norm) @ B0inv
        self.D2 = ((d2cheb * (w * keep)[:, None]).T * self.cheb_norm) @ B0inv


def _cc2real(c, m):
    """useful.f90 cc2real: |c|^2 (m=0) or 2|c|^2."""
    return np.where(m == 0, 1.0, 2.0) * (c.real ** 2 + c.imag ** 2)


class ShellHost:
    """The LM side of MagIC for a spherical shell with fixed-entropy boundaries, Chebyshev collocation, CN/AB2 and the
    pressure ("WP") formulation with separate matrices.  Options: polytropic anelastic reference state (strat, polind,
    g0/g1/g2), rigid (2) or stress-free (1) walls, magnetic field with insulating boundaries, n_cheb_max < n_r_max,
    l_correct_AMz / l_correct_AMe."""

    def __init__(self, lm2l, lm2m, radial_loop, n_r_max=33, n_cheb_max=None, radratio=0.35, ra=1e5, ek=1e-3, pr=1.0,
                 prmag=5.0, dtmax=1e-4, alpha=0.6, init_s1=404, amp_s1=0.1, init_b1=3, amp_b1=5.0, l_mag=True, ktopv=2,
                 kbotv=2, strat=0.0, polind=2.0, g0=0.0, g1=1.0, g2=0.0, l_correct_AMz=False, l_correct_AMe=False, l_heat=True,
                 po=0.0, prec_angle=23.5, l_cond_ic=False, l_rot_ic=False, sigma_ratio=1.0, n_r_ic_max=17, n_cheb_ic_max=15, var_cond=None, raxi=0.0, sc=1.0, dif_exp=None,
                 phase=None):
        self.lm2l = np.asarray(lm2l)
        self.lm2m = np.asarray(lm2m)
        self.lm_max = len(self.lm2l)
        self.l_max = int(self.lm2l.max())
        self.g = g = ChebShell(n_r_max, radratio, n_cheb_max)
        self.N = n_r_max
        self.radial_loop = radial_loop
        self.l_mag, self.ktopv, self.kbotv = l_mag, ktopv, kbotv
        self.l_correct_AMz, self.l_correct_AMe = l_correct_AMz, l_correct_AMe
        self.opr, self.opm = 1.0 / pr, 1.0 / prmag
        self.l_heat = l_heat and ra != 0.0    # Namelists.f90:429
        self.BuoFac = ra / pr if self.l_heat else 0.0   # preCalculations.f90:170-179 (lScale=1)
        # precession (Namelists.f90:581-588, preCalculations.f90:164-166, updateZ.f90:231-235): the Poincare force on (1,1)
        self.oek = 1.0 / ek
        self.prec_fac = np.sqrt(8.0 * np.pi / 3.0) * po * self.oek ** 2 * np.sin(np.deg2rad(prec_angle)) if po != 0.0 else 0.0
        self.LFfac = 1.0 / (ek * prmag)       # preCalculations.f90:158
        self.l_chem = raxi != 0.0             # Namelists.f90:421-427: chemical convection (double diffusion)
        self.ChemFac, self.osc = (raxi / sc if self.l_chem else 0.0), 1.0 / sc           # preCalculations.f90:181-185
        r, or1, or2 = g.r, g.or1, g.or2
        self.rgrav = g0 + g1 * r / g.r_cmb + g2 * g.r_cmb ** 2 * or2            # radial.f90:628
        self.l_anel = strat > 0.0
        if self.l_anel:   # adiabatic polytropic reference state, radial.f90:673-729
            self.DissNb = (np.exp(strat / polind) - 1.0) / (g.r_cmb - g.r_icb) / (g0 + 0.5 * g1 * (1.0 + radratio) + g2 / radratio)
            temp0 = -self.DissNb * (g0 * r + 0.5 * g1 * r ** 2 / g.r_cmb - g2 * g.r_cmb ** 2 * or1) + 1.0 + \
                self.DissNb * g.r_cmb * (g0 + 0.5 * g1 - g2)
            dg = g1 / g.r_cmb - 2.0 * g2 * g.r_cmb ** 2 * or1 ** 3
            self.temp0 = temp0
            self.rho0 = temp0 ** polind
            self.beta = -polind * self.DissNb * self.rgrav / temp0
            self.dbeta = -polind * self.DissNb / temp0 ** 2 * (dg * temp0 + self.DissNb * self.rgrav ** 2)
            self.dLtemp0 = -self.DissNb * self.rgrav / temp0
            self.ViscHeatFac = self.DissNb * pr / ra                            # radial.f90:762
        else:
            one = np.ones(n_r_max)
            self.DissNb, self.temp0, self.rho0, self.beta, self.dbeta, self.dLtemp0 = 0.0, one, one, 0 * one, 0 * one, 0 * one
            self.ViscHeatFac = 0.0
        self.orho1 = 1.0 / self.rho0
        # transport properties (radial.f90 transportProperties): nVarVisc = nVarDiff = 2, visc = kappa = (rho0 / rho0(icb))^difExp,
        # logarithmic derivatives through get_dr as in the reference
        one_r = np.ones(n_r_max)
        self.visc, self.dLvisc, self.ddLvisc, self.kappa, self.dLkappa = one_r, 0 * one_r, 0 * one_r, one_r, 0 * one_r
        if dif_exp is not None:
            self.visc = (self.rho0 / self.rho0[-1]) ** dif_exp
            self.dLvisc = (g.D1t @ self.visc) / self.visc
            self.ddLvisc = g.D1t @ self.dLvisc
            self.kappa = self.visc.copy()
            self.dLkappa = self.dLvisc.copy()
        # heating prefactors (radial.f90:758-766) and magnetic diffusivity profile (radial.f90:903-916, nVarCond = 2: the
        # two-branch conductivity of Gomez-Perez et al.; var_cond = dict(con_DecRate, con_RadRatio, con_LambdaMatch))
        self.OhmLossFac = self.ViscHeatFac / (ek * prmag ** 2) if (self.l_anel and l_mag) else 0.0
        self.lam, self.dLlam = np.ones(n_r_max), np.zeros(n_r_max)
        if var_cond is not None and l_mag:
            r0 = var_cond["con_RadRatio"] * g.r_cmb
            r0 = r[np.argmin(np.abs(r - r0))]
            LM, DR = var_cond["con_LambdaMatch"], var_cond["con_DecRate"]
            ds0 = (LM - 1.0) * DR / (r0 - g.r_icb)
            x = (r - g.r_icb) / (r0 - g.r_icb)
            inner = r < r0
            sigma = np.where(inner, 1.0 + (LM - 1.0) * x ** DR, LM * np.exp(ds0 / LM * (r - r0)))
            dsigma = np.where(inner, ds0 * x ** (DR - 1.0), ds0 * np.exp(ds0 / LM * (r - r0)))
            self.lam, self.dLlam = 1.0 / sigma, -dsigma / sigma
        self.c_moi_oc = 8.0 / 3.0 * np.pi * g.rInt_R(r ** 4 * self.rho0)         # preCalculations.f90:329-330
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
        self.xi = z()
        self.topxi = np.zeros(lm_max, dtype=np.complex128)
        self.botxi = np.zeros(lm_max, dtype=np.complex128)
        if self.l_chem:
            self.botxi[self._lm(0, 0)] = sq4pi      # ktopxi = kbotxi = 1 (preCalculations.f90:613-617)
        # conducting and / or freely rotating inner core (kbotb=3, nRotIC=1; Namelists.f90:398-407, :737-739)
        self.l_cond_ic, self.l_rot_ic, self.O_sr = l_cond_ic and l_mag, l_rot_ic, 1.0 / sigma_ratio
        self.omega_ic = 0.0
        self.lorentz_torque_ic = 0.0
        self.c_z10_omega_ic = 0.5 * np.sqrt(3.0 / np.pi) * or2[-1] / self.rho0[-1]      # preCalculations.f90:313-318
        self.c_dt_z10_ic = 0.2 * g.r_icb * self.rho0[-1]                                # :336 (rho_ratio_ic = 1)
        self.c_lorentz_ic = 0.25 * np.sqrt(3.0 / np.pi) * or2[-1]                       # :343
        self.c_moi_ic = 8.0 * np.pi / 15.0 * g.r_icb ** 5 * self.rho0[-1]               # :326
        self.sigma_ratio, self.prmag = sigma_ratio, prmag
        self.aj_nl_icb = None
        self.dom_ic = dict(old=0.0, impl=0.0, expl=[0.0, 0.0])
        if self.l_cond_ic:
            self.ic = ChebEvenIC(n_r_ic_max, n_cheb_ic_max, g.r_icb)
            zi = lambda: np.zeros((n_r_ic_max, lm_max), dtype=np.complex128)
            self.b_ic, self.aj_ic = zi(), zi()
        self.tops = np.zeros(lm_max, dtype=np.complex128)
        self.bots = np.zeros(lm_max, dtype=np.complex128)
        lm00 = self._lm(0, 0)
        if self.l_heat:
            self.bots[lm00] = sq4pi           # preCalculations.f90:415-418
        # ---- initS: conductive state (ps_cond, entropy diffusion, epsc=0; init_fields.f90:2270-2291) + one mode (:428-541)
        M = self.opr * self.kappa[:, None] * (g.D2 + (self.beta + self.dLtemp0 + 2.0 * or1 + self.dLkappa)[:, None] * g.D1)
        M[0] = 0.0
        M[0, 0] = 1.0
        M[-1] = 0.0
        M[-1, -1] = 1.0
        rhs = np.zeros((N, 1))
        rhs[0, 0], rhs[-1, 0] = self.tops[lm00].real, self.bots[lm00].real
        if self.l_heat:
            self.s[:, lm00] = g.solve(M, rhs, (0, N - 1))[:, 0]
        if self.l_heat and init_s1 >= 100:
            l, m = init_s1 // 100, init_s1 % 100
            x = 2.0 * r - g.r_cmb - g.r_icb
            s1 = 1.0 - 3.0 * x ** 2 + 3.0 * x ** 4 - x ** 6
            self.s[:, self._lm(l, m)] += amp_s1 * s1
        # ---- phase field (l_phase_field; Namelists.f90:565-573, init_fields.f90:1431-1481): phase = dict(stef, tmelt, phaseDiffFac,
        #      penaltyFac, epsPhase, ktopphi, kbotphi); init_phi = 1: a tanh front of width epsPhase at the melting radius of the
        #      conductive state, no entropy perturbation in the solid
        self.l_phase = phase is not None
        self.phase = phase or {}
        self.phi = z()
        self.pr = pr
        if self.l_phase:
            temp00 = self.s[:, lm00].real / sq4pi
            n_melt = None
            for n in range(1, N):
                if temp00[n - 1] < phase["tmelt"] <= temp00[n]:
                    n_melt = n
            phi0 = 0.5 * (1.0 + np.tanh((r - r[n_melt]) / 2.0 / np.sqrt(2.0) / phase["epsPhase"]))
            self.phi[:, lm00] = sq4pi * phi0
            self.s[:, self.lm2l != 0] *= (1.0 - phi0)[:, None]
            self.phi_top = 0.0 if phase.get("ktopphi", 1) != 1 else sq4pi     # Namelists.f90:567-572
            self.phi_bot = 0.0
        # ---- initB, init_b1=3, insulating inner core (init_fields.f90:1129-1188)
        if l_mag and init_b1 == 3 and self.l_cond_ic:   # init_fields.f90:1140-1176
            b_pol = amp_b1 * np.sqrt(3.0 * np.pi) / (3.0 + g.r_cmb)
            b_tor = -4.0 / 3.0 * amp_b1 * np.sqrt(np.pi / 5.0)
            ri = self.ic.r
            self.b[:, self._lm(1, 0)] += b_pol * (r ** 3 - 4.0 / 3.0 * g.r_cmb * r ** 2)
            self.b_ic[:, self._lm(1, 0)] += b_pol * g.r_icb ** 2 * (0.5 * ri ** 2 / g.r_icb + 0.5 * g.r_icb - 4.0 / 3.0 * g.r_cmb)
            self.aj[:, self._lm(2, 0)] += b_tor * r * np.sin(np.pi * r / g.r_cmb)
            arg = np.pi * g.r_icb / g.r_cmb
            aj_ic1 = (arg - 2.0 * np.sin(arg) * np.cos(arg)) / (arg + np.sin(arg) * np.cos(arg))
            aj_ic2 = (1.0 - aj_ic1) * g.r_icb * np.sin(arg) / np.cos(arg)
            self.aj_ic[:, self._lm(2, 0)] += b_tor * (aj_ic1 * ri * np.sin(np.pi * ri / g.r_cmb) + aj_ic2 * np.cos(np.pi * ri / g.r_cmb))
        elif l_mag and init_b1 == 3:
            b_pol = amp_b1 * np.sqrt(3.0 * np.pi) / 4.0
            b_tor = -4.0 / 3.0 * amp_b1 * np.sqrt(np.pi / 5.0)
            self.b[:, self._lm(1, 0)] += b_pol * (r ** 3 - 4.0 / 3.0 * g.r_cmb * r ** 2 + g.r_icb ** 4 / 3.0 * or1)
            self.aj[:, self._lm(2, 0)] += b_tor * r * np.sin(np.pi * (r - g.r_icb))
        # ---- time arrays (time_array.f90): old, impl (one level each for CNAB2), expl (two levels)
        self.old, self.impl, self.expl = {}, {}, {}
        for nm in ("s", "xi", "phi", "w", "p", "z", "b", "j"):
            self.old[nm], self.impl[nm] = z(), z()
            self.expl[nm] = [z(), z()]
        if self.l_cond_ic:
            for nm in ("b_ic", "j_ic"):
                self.old[nm], self.impl[nm] = zi(), zi()
                self.expl[nm] = [zi(), zi()]
        self._mats = None
        # startFields.f90:373-432: derivatives and old/implicit terms of the start fields
        self._rhs_imp_phi()
        self._rhs_imp_s()
        self._rhs_imp_xi()
        self._rhs_imp_wp()
        self._rhs_imp_z()
        if l_mag:
            self._rhs_imp_b()

    def load_checkpoint(self, ck):
        """Start from a MagIC checkpoint (magic_b200.checkpoint.Checkpoint, same grid and truncation -- no remapping):
        fields, time, rotation rate of the inner core and, for a multistep file, the explicit terms of the previous step
        (readCheckPoints.f90:767-1468); then startFields.f90:373-432 rebuilds the old / implicit terms."""
        assert np.abs(ck.r - self.g.r).max() < 1e-13 and ck.lm_max == self.lm_max
        names = {"w": "w", "z": "z", "p": "p", "s": "s", "xi": "xi", "b": "b", "aj": "aj", "b_ic": "b_ic", "aj_ic": "aj_ic"}
        tarr = {"w": "w", "z": "z", "p": "p", "s": "s", "xi": "xi", "b": "b", "aj": "j", "b_ic": "b_ic", "aj_ic": "j_ic"}
        for nm, attr in names.items():
            if nm in ck.fields and hasattr(self, attr):
                setattr(self, attr, np.array(ck.fields[nm], dtype=np.complex128))
                past = ck.past.get(nm, {}).get("expl", [])
                if past and tarr[nm] in self.expl:
                    self.expl[tarr[nm]][1] = np.array(past[0], dtype=np.complex128)
        # startFields.f90:257-279: a file written with the legacy boundary values (-ri^2, ro^2) / (ri^2 + ro^2) * sqrt(4 pi)
        # of s(0,0) / xi(0,0) is translated to the present convention (0, sqrt(4 pi))
        g = self.g
        lm00 = self._lm(0, 0)
        topval = -g.r_icb ** 2 / (g.r_icb ** 2 + g.r_cmb ** 2) * np.sqrt(4.0 * np.pi)
        botval = g.r_cmb ** 2 / (g.r_icb ** 2 + g.r_cmb ** 2) * np.sqrt(4.0 * np.pi)
        for f in (self.s, self.xi):
            if abs(f[0, lm00] - topval) <= 1e4 * np.finfo(float).eps and abs(f[-1, lm00] - botval) <= 1e4 * np.finfo(float).eps:
                f[:, lm00] -= topval
        self.time = float(ck.time)
        self.omega_ic = float(ck.rotation["omega_ic1"])
        past = ck.scalars_past.get("domega_ic_dt", {}).get("expl", [])
        if len(past):
            self.dom_ic["expl"][1] = float(past[0])
        self.restart()

    _STATE_FIELDS = ("w", "z", "p", "s", "xi", "b", "aj", "b_ic", "aj_ic")

    def state_dict(self):
        """What a checkpoint of a multistep run holds (storeCheckPoints.f90:45-277): fields, the explicit terms of the
        previous step, rotation rate and its explicit term, time and time step -- as a flat dict of arrays (np.savez)."""
        d = {"time": self.time, "dt": self.dt, "omega_ic": self.omega_ic, "dom_ic_expl2": self.dom_ic["expl"][1]}
        for nm in self._STATE_FIELDS:
            if hasattr(self, nm):
                d["field_" + nm] = getattr(self, nm)
        for nm, e in self.expl.items():
            d["expl2_" + nm] = e[1]
        return d

    def load_state_dict(self, d):
        for nm in self._STATE_FIELDS:
            if "field_" + nm in d and hasattr(self, nm):
                setattr(self, nm, np.array(d["field_" + nm], dtype=np.complex128))
        for nm in self.expl:
            if "expl2_" + nm in d:
                self.expl[nm][1] = np.array(d["expl2_" + nm], dtype=np.complex128)
        self.time, self.dt = float(d["time"]), np.array(d["dt"], dtype=float)
        self.omega_ic = float(d["omega_ic"])
        self.dom_ic["expl"][1] = float(d["dom_ic_expl2"])
        self.restart()

    def restart(self, **flags):
        """A run restarted from its own checkpoint with other switches (samples/*/input_restart.nml): the checkpoint carries
        the fields, the explicit terms of the previous step and the rotation rates (storeCheckPoints.f90:45-277), all of
        which this object still holds; startFields.f90:373-432 then rebuilds the old / implicit terms under the new switches
        (this is where l_correct_AMz / AMe first act on the restart state)."""
        for k, v in flags.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)
        self._mats = None
        self._rhs_imp_phi()
        self._rhs_imp_s()
        self._rhs_imp_xi()
        self._rhs_imp_wp()
        self._rhs_imp_z()
        if self.l_mag:
            self._rhs_imp_b()

    # ------------------------------------------------------------------------------------------------
    def _lm(self, l, m):
        return int(np.nonzero((self.lm2l == l) & (self.lm2m == m))[0][0])

    def _rhs_imp_s(self):
        """get_entropy_rhs_imp, updateS.f90:658-756 (entropy diffusion, kappa=1)."""
        g = self.g
        # get_ddr differentiates the modes below n_cheb_max only: the same as D1, D2 on solved fields (those hold no higher modes),
        # not on start fields with a sharp front (initPhi scales the entropy perturbation by 1 - phi0)
        D1, D2 = (g.D1t, g.D2t) if self.l_phase else (g.D1, g.D2)
        self.ds = D1 @ self.s
        dds = D2 @ self.s
        self.old["s"] = self.s - self.phase["stef"] * self.phi if self.l_phase else self.s.copy()     # updateS.f90:706-716
        self.impl["s"] = self.opr * self.kappa[:, None] * (
            dds + (self.beta + self.dLtemp0 + 2.0 * g.or1 + self.dLkappa)[:, None] * self.ds -
            self.dL[None, :] * g.or2[:, None] * self.s)

    def _rhs_imp_phi(self):
        """get_phase_rhs_imp, updatePHI.f90:470-542."""
        if not self.l_phase:
            return
        g, ph = self.g, self.phase
        dphi = g.D1t @ self.phi
        ddphi = g.D2t @ self.phi
        self.old["phi"] = 5.0 / 6.0 * ph["stef"] * self.pr * self.phi
        self.impl["phi"] = ph["phaseDiffFac"] * (ddphi + 2.0 * g.or1[:, None] * dphi - self.dL[None, :] * g.or2[:, None] * self.phi)

    def _rhs_imp_xi(self):
        """get_comp_rhs_imp, updateXI.f90:579-650."""
        g = self.g
        self.dxi = g.D1 @ self.xi
        ddxi = g.D2 @ self.xi
        self.old["xi"] = self.xi.copy()
        self.impl["xi"] = self.osc * (ddxi + (self.beta + 2.0 * g.or1)[:, None] * self.dxi -
                                      self.dL[None, :] * g.or2[:, None] * self.xi)

    def _rhs_imp_z(self):
        """get_tor_rhs_imp, updateZ.f90:760-1025 (visc=1, no rotating walls), with the angular-momentum corrections."""
        g = self.g
        r, beta, dbeta, rho0 = g.r, self.beta, self.dbeta, self.rho0
        self.dz = g.D1 @ self.z
        ddz = g.D2 @ self.z
        fac3 = 8.0 / 3.0 * np.pi
        for (l, m, on) in ((1, 0, self.l_correct_AMz), (1, 1, self.l_correct_AMe)):
            if not on or m > self.lm2m.max():
                continue
            lm = self._lm(l, m)
            # updateZ.f90:840-915 + outRot.f90:485-543 with AMstart=0 and non-rotating walls: remove the rigid rotation
            # rho0 r^2 corr that carries the angular momentum of z(1,m)
            corr = fac3 * g.rInt_R(r * r * self.z[:, lm]) / self.c_moi_oc
            if m == 0:
                corr = corr.real
                if self.l_rot_ic:   # angular_moment_ic(3) = c_moi_ic * omega_ic (outRot.f90:538), nomi = c_moi_oc * y10_norm
                    corr += self.c_moi_ic * self.omega_ic / (self.c_moi_oc * 0.5 * np.sqrt(3.0 / np.pi))
            self.z[:, lm] -= rho0 * r * r * corr
            self.dz[:, lm] -= rho0 * (2.0 * r + r * r * beta) * corr
            ddz[:, lm] -= rho0 * (2.0 + 4.0 * beta * r + dbeta * r * r + beta * beta * r * r) * corr
        fac = self.dL[None, :] * g.or2[:, None]
        self.old["z"] = fac * self.z
        dLv = self.dLvisc
        imp = fac * self.visc[:, None] * (ddz + (dLv - beta)[:, None] * self.dz -
                                          (fac + (dLv * beta + 2.0 * dLv * g.or1 + dbeta + 2.0 * beta * g.or1)[:, None]) * self.z)
        if self.prec_fac != 0.0:   # updateZ.f90:955-958, evaluated at the time the fields belong to
            imp[:, self._lm(1, 1)] += self.prec_fac * (np.sin(self.oek * self.time) - 1j * np.cos(self.oek * self.time))
        imp[0] = 0.0
        imp[-1] = 0.0      # n_r_top=n_r_cmb+1 .. n_r_bot=n_r_icb-1
        self.impl["z"] = imp
        if self.l_rot_ic and self.kbotv == 2:   # updateZ.f90:1001-1008: viscous torque on the inner core (visc = 1)
            z10, dz10 = self.z[-1, self._lm(1, 0)].real, self.dz[-1, self._lm(1, 0)].real
            self.dom_ic["old"] = self.c_dt_z10_ic * z10
            self.dom_ic["impl"] = -((2.0 * g.or1[-1] + beta[-1]) * z10 - dz10)
        elif self.l_rot_ic:                     # updateZ.f90:997-999: stress-free, only the Lorentz torque acts
            self.dom_ic["old"] = self.c_moi_ic * self.c_lorentz_ic * self.omega_ic
            self.dom_ic["impl"] = 0.0

    def _rhs_imp_wp(self):
        """get_pol_rhs_imp, updateWP.f90:1089-1336, non double-curl branch (visc=1, dLvisc=0)."""
        g = self.g
        beta, dbeta, or1 = self.beta[:, None], self.dbeta[:, None], g.or1[:, None]
        self.dw = g.D1 @ self.w
        self.ddw = g.D2 @ self.w
        dddw = g.D3 @ self.w
        self.dp = g.D1 @ self.p
        fac = self.dL[None, :] * g.or2[:, None]
        old_w = fac * self.w
        old_p = -fac * self.dw
        visc, dLv = self.visc[:, None], self.dLvisc[:, None]
        Dif = fac * visc * (self.ddw + (2.0 * dLv - beta / 3.0) * self.dw -
                            (fac + 4.0 / 3.0 * (dbeta + dLv * beta + (3.0 * dLv + beta) * or1)) * self.w)
        Pre = -self.dp + beta * self.p
        Buo = (self.rho0 * self.rgrav)[:, None] * (self.BuoFac * self.s + self.ChemFac * self.xi)
        imp_w = Pre + Dif + Buo
        imp_p = fac * self.p + fac * visc * (-dddw + (beta - dLv) * self.ddw +
                                             (fac + dLv * beta + dbeta + 2.0 * (dLv + beta) * or1) * self.dw -
                                             fac * (2.0 * or1 + 2.0 / 3.0 * beta + dLv) * self.w)
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
        lam, dLlam = self.lam[:, None], self.dLlam[:, None]
        ib = self.opm * lam * fac * (self.ddb - fac * self.b)
        ij = self.opm * lam * fac * (ddj + dLlam * self.dj - fac * self.aj)
        for a in (ib, ij):
            a[0] = 0.0
            a[-1] = 0.0
        self.impl["b"], self.impl["j"] = ib, ij
        if self.l_cond_ic:   # get_mag_ic_rhs_imp, updateB.f90:1077-1187
            ic = self.ic
            dLN = self.dL[None, :] * g.or2[-1]
            lp1 = (self.lm2l + 1.0)[None, :]
            for nm, f in (("b_ic", self.b_ic), ("j_ic", self.aj_ic)):
                df, ddf = ic.D1 @ f, ic.D2 @ f
                self.old[nm] = dLN * f
                imp = self.opm * self.O_sr * dLN * (ddf + 2.0 * lp1 * ic.O_r[:, None] * df)
                imp[-1] = (self.opm * self.O_sr * dLN * (1.0 + 2.0 * lp1) * ddf)[-1]     # r = 0: 1/r d/dr -> d2/dr2
                imp[0] = 0.0
                imp[:, self.lm2l == 0] = 0.0
                self.old[nm][:, self.lm2l == 0] = 0.0
                self.impl[nm] = imp

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
        """get_sMat / get_zMat / get_wpMat / get_bMat for every degree, acting on grid values (ChebShell.solve maps them
        to coefficient space and applies the dealiasing)."""
        g, N = self.g, self.N
        g._lu = {}
        I = np.eye(N)
        beta, dbeta = self.beta[:, None], self.dbeta[:, None]
        or1, or2 = g.or1[:, None], g.or2[:, None]
        b0, bN = self.beta[0], self.beta[-1]
        mats = {"s": [], "xi": [], "z": [], "wp": [], "b": [], "j": []}
        for l in range(self.l_max + 1):
            dL = float(l * (l + 1))
            # sMat (updateS.f90:1086-1140), ktops=kbots=1
            M = I - wl1 * self.opr * self.kappa[:, None] * (
                g.D2 + (beta + self.dLtemp0[:, None] + 2.0 * or1 + self.dLkappa[:, None]) * g.D1 - dL * or2 * I)
            M[0], M[-1] = I[0], I[-1]
            mats["s"].append(M)
            if self.l_phase:   # phiMat (updatePHI.f90:805-895): Dirichlet (1) or Neumann boundary rows
                ph = self.phase
                M = 5.0 / 6.0 * ph["stef"] * self.pr * I - wl1 * ph["phaseDiffFac"] * (g.D2 + 2.0 * or1 * g.D1 - dL * or2 * I)
                M[0] = I[0] if ph.get("ktopphi", 1) == 1 else g.D1[0]
                M[-1] = I[-1] if ph.get("kbotphi", 1) == 1 else g.D1[-1]
                mats.setdefault("phi", []).append(M)
            # xiMat (updateXI.f90:924-965), ktopxi = kbotxi = 1
            M = I - wl1 * self.osc * (g.D2 + (beta + 2.0 * or1) * g.D1 - dL * or2 * I)
            M[0], M[-1] = I[0], I[-1]
            mats["xi"].append(M)
            # zMat (updateZ.f90:1850-1890)
            visc, dLv = self.visc[:, None], self.dLvisc[:, None]
            M = dL * or2 * I - wl1 * dL * or2 * visc * (
                g.D2 + (dLv - beta) * g.D1 - (dLv * beta + 2.0 * dLv * or1 + dL * or2 + dbeta + 2.0 * beta * or1) * I)
            M[0] = I[0] if self.ktopv == 2 else g.D1[0] - (2.0 * g.or1[0] + b0) * I[0]
            M[-1] = I[-1] if self.kbotv == 2 else g.D1[-1] - (2.0 * g.or1[-1] + bN) * I[-1]
            mats["z"].append(M)
            # wpMat (updateWP.f90:1999-2090)
            W = np.zeros((2 * N, 2 * N))
            W[:N, :N] = dL * or2 * I - wl1 * dL * or2 * visc * (
                g.D2 + (2.0 * dLv - beta / 3.0) * g.D1 -
                (dL * or2 + 4.0 / 3.0 * (dLv * beta + (3.0 * dLv + beta) * or1 + dbeta)) * I)
            W[:N, N:] = wl1 * (g.D1 - beta * I)
            W[N:, :N] = -dL * or2 * g.D1 - wl1 * dL * or2 * visc * (
                -g.D3 + (beta - dLv) * g.D2 + (dL * or2 + dbeta + dLv * beta + 2.0 * (dLv + beta) * or1) * g.D1 -
                dL * or2 * (2.0 * or1 + dLv + 2.0 / 3.0 * beta) * I)
            W[N:, N:] = -wl1 * dL * or2 * I
            W[0] = 0.0
            W[0, :N] = I[0]
            W[N - 1] = 0.0
            W[N - 1, :N] = I[-1]
            W[N] = 0.0
            W[N, :N] = g.D1[0] if self.ktopv == 2 else g.D2[0] - (2.0 * g.or1[0] + b0) * g.D1[0]
            W[2 * N - 1] = 0.0
            W[2 * N - 1, :N] = g.D1[-1] if self.kbotv == 2 else g.D2[-1] - (2.0 * g.or1[-1] + bN) * g.D1[-1]
            mats["wp"].append(W)
            if self.l_mag:
                # bMat / jMat (updateB.f90:1824-1905), ktopb=kbotb=1, conductance_ma=0
                lam, dLlam = self.lam[:, None], self.dLlam[:, None]
                B = dL * or2 * I - wl1 * self.opm * lam * dL * or2 * (g.D2 - dL * or2 * I)
                J = dL * or2 * I - wl1 * self.opm * lam * dL * or2 * (g.D2 + dLlam * g.D1 - dL * or2 * I)
                B[0] = g.D1[0] + l * g.or1[0] * I[0]
                B[-1] = g.D1[-1] - (l + 1.0) * g.or1[-1] * I[-1]
                J[0], J[-1] = I[0], I[-1]
                mats["b"].append(B)
                mats["j"].append(J)
        if self.l_rot_ic and self.kbotv == 2:
            # get_z10Mat (updateZ.f90:1719-1800): zMat(l=1) with the torque balance of the inner core in the ICB row
            M = mats["z"][1].copy()
            M[-1] = self.c_dt_z10_ic * I[-1] + wl1 * ((2.0 * g.or1[-1] + bN) * I[-1] - g.D1[-1])
            mats["z10"] = M
        if self.l_mag and self.l_cond_ic:
            mats["bic"], mats["jic"] = [None], [None]
            for l in range(1, self.l_max + 1):
                for nm, sr in (("bic", 1.0), ("jic", 1.0 / self.O_sr)):
                    mats[nm].append(self._coupled_mag_matrix(l, wl1, mats["b" if nm == "bic" else "j"][l], sr))
        self._mats = (wl1, mats)

    def _coupled_mag_matrix(self, l, wl1, M_oc, sr):
        """get_bMat with kbotb = 3 (updateB.f90:1813-2060): outer-core rows as for the insulating case, then continuity of the
        potential and of its radial derivative (times sigma_ratio for the toroidal part) at the ICB, then the diffusion
        equation of the inner core for g(r) with potential = (r / r_icb)^(l+1) g(r).  Acts on COEFFICIENTS (outer-core
        Chebyshev modes, then inner-core even modes); boundary rows are blind to the dealiased modes."""
        g, ic, N = self.g, self.ic, self.N
        Ni, nc, nci = ic.n_r_ic_max, g.n_cheb_max, ic.n_cheb_ic_max
        dL, lp1 = float(l * (l + 1)), l + 1.0
        A = np.zeros((N + Ni, N + Ni))
        oc = M_oc.copy()
        oc[-1] = np.eye(N)[-1]                               # row N: b_oc(r_icb) ...
        A[:N, :N] = oc @ g.B0
        A[N, :N] = (sr * g.D1[-1]) @ g.B0                     # row N+1: d/dr of the outer-core potential ...
        for row in (0, N - 1, N):
            A[row, nc:N] = 0.0
        cn = ic.cheb_norm
        cheb, dcheb, d2cheb = ic.cheb, ic.dcheb, ic.d2cheb    # [mode, point]
        fac = dL * g.or2[-1]
        blk = np.zeros((Ni + 1, Ni))                          # rows: N-1 (continuity), N (derivative), N+1 .. (IC points 2 ..)
        blk[0, :nci] = -cn * cheb[:nci, 0]
        blk[1, :nci] = -cn * (dcheb[:nci, 0] + lp1 * g.or1[-1] * cheb[:nci, 0])
        for k in range(1, Ni - 1):
            blk[1 + k] = cn * fac * (cheb[:, k] - wl1 * self.opm * self.O_sr * (d2cheb[:, k] + 2.0 * lp1 * ic.O_r[k] * dcheb[:, k]))
        blk[Ni] = cn * fac * (cheb[:, -1] - wl1 * self.opm * self.O_sr * (1.0 + 2.0 * lp1) * d2cheb[:, -1])
        blk[:, 0] *= 0.5
        blk[:, -1] *= 0.5
        A[N - 1:, N:] = blk
        return A

    def _solve_coupled(self, A, rhs_oc, rhs_ic, icb_derivative_bc=None, key=None):
        """Right-hand side rows as updateB.f90:500-548; returns grid values (outer core, inner core)."""
        g, ic, N = self.g, self.ic, self.N
        rhs = np.concatenate([rhs_oc, rhs_ic], axis=0)
        rhs[0] = 0.0
        rhs[N - 1] = 0.0
        rhs[N] = 0.0 if icb_derivative_bc is None else icb_derivative_bc
        fac = g._lu.get(key) if key is not None else None
        if fac is None:
            f = 1.0 / np.max(np.abs(A), axis=1)
            fac = (A * f[:, None], f)
            if key is not None:
                g._lu[key] = fac
        c = np.linalg.solve(fac[0], rhs * fac[1][:, None])
        co, ci = c[:N].copy(), c[N:].copy()
        co[g.n_cheb_max:] = 0.0
        ci[ic.n_cheb_ic_max:] = 0.0
        return g.B0 @ co, ic.B0 @ ci

    # ------------------------------------------------------------------------------------------------
    def fields_Rloc(self):
        """What transp_LMloc_to_Rloc hands to the radial loop (step_time.f90:1005-1132)."""
        f = dict(w=self.w, dw=self.dw, ddw=self.ddw, z=self.z, dz=self.dz)
        if self.l_heat:
            f["s"] = self.s
        if self.l_chem:
            f["xi"] = self.xi
        if self.l_phase:
            f["phi"] = self.phi
        if self.l_mag:
            f.update(b=self.b, db=self.db, ddb=self.ddb, aj=self.aj, dj=self.dj)
        return f

    def step(self):
        """One pass of the n_time_step loop of step_time.f90:480-763 (CNAB2: one stage)."""
        out = self.radial_loop({k: np.ascontiguousarray(v) for k, v in self.fields_Rloc().items()})
        self._explicit(out)
        # dt_courant (courant.f90:277-346)
        dt_min = min(self.dtrkc_min, self.dthkc_min, 1000.0 * self.dtmax)
        if self.dt[0] > dt_min:
            raise RuntimeError("Courant criterion asks for a smaller time step; not expected in this sample")
        self.dt = np.array([self.dt[0], self.dt[0]])                                    # set_dt_array
        wts = self._weights()
        wimp, wl1, wl2, we1, we2 = wts
        if self._mats is None or self._mats[0] != wl1:
            self._build_mats(wl1)
        self.time += self.dt[0]
        d = self.dom_ic

        def rotate(nm):
            if nm == "dom_ic":
                d["expl"][1] = d["expl"][0]
            else:
                self.expl[nm][1] = self.expl[nm][0]

        self._expl_w_stage = self.expl["w"][0]     # dwdt%expl(:,:,1) of a multistep scheme: the newest explicit term
        self._lm_loop(lambda nm: self._imex_rhs(nm, wts),
                      lambda: wimp * d["old"] + wl2 * d["impl"] + we1 * d["expl"][0] + we2 * d["expl"][1], wl1, rotate)
        self.n_steps += 1

    def _explicit(self, out):
        """finish_explicit_assembly (LMLoop.f90:390-453, dentropy0 = 0) on the output of one radial loop: the explicit
        terms of the current stage go to self.expl[*][0] (and self.dom_ic['expl'][0])."""
        g = self.g
        or2 = g.or2[:, None]
        l0 = (self.lm2l == 0)[None, :]
        if self.l_heat:
            self.expl["s"][0] = self.orho1[:, None] * (out["dsdt"] - or2 * (g.D1t @ out["dVSrLM"]))   # updateS.f90:587-597
        if self.l_chem:
            self.expl["xi"][0] = self.orho1[:, None] * (out["dxidt"] - or2 * (g.D1t @ out["dVXirLM"]))   # updateXI.f90:495-511
        if self.l_phase:
            self.expl["phi"][0] = np.array(out["dphidt"])          # rIter.f90:698, no finish step (LMLoop.f90:390-453)
        self.expl["w"][0] = np.array(out["dwdt"])
        self.expl["p"][0] = np.array(out["dpdt"])
        self.expl["z"][0] = np.array(out["dzdt"])
        if self.l_mag:
            self.expl["b"][0] = np.array(out["dbdt"])
            self.expl["j"][0] = out["djdt"] + np.where(l0, 0.0, or2 * (g.D1t @ out["dVxBhLM"]))  # updateB.f90:1030-1037
        if self.l_rot_ic:     # finish_exp_tor (updateZ.f90:1659-1688), gammatau_gravi = 0
            self.lorentz_torque_ic = float(out["lorentz_torque_ic"])
            self.dom_ic["expl"][0] = self.c_lorentz_ic * self.lorentz_torque_ic
        if self.l_cond_ic and self.kbotv == 1:   # l_b_nl_icb (Namelists.f90:713-720); get_b_nl_bcs('ICB') nonlinear_bcs.f90:104-111
            self.aj_nl_icb = -self.sigma_ratio * self.prmag * np.asarray(out["br_vp_lm_icb"])
            self.aj_nl_icb[0] = 0.0
        else:
            self.aj_nl_icb = None
        if self.l_cond_ic:    # finish_exp_mag_ic (updateB.f90:955-1003): advection of the inner-core field by its rotation
            fac = -self.omega_ic * g.or2[-1] * 1j * self.lm2m[None, :] * self.dL[None, :] if self.l_rot_ic else 0.0
            for nm, f in (("b_ic", self.b_ic), ("j_ic", self.aj_ic)):
                e = fac * f
                e[0] = 0.0
                self.expl[nm][0] = e
        self.dtrkc_min, self.dthkc_min = float(np.min(out["dtrkc"])), float(np.min(out["dthkc"]))

    def _lm_loop(self, rhs_of, dom_ic_rhs, wl1, rotate):
        """LMLoop (LMLoop.f90:150-330): updateS, updateZ, updateWP, updateB in the reference's order.  rhs_of(name) is the
        time scheme's set_imex_rhs for the time array `name`, dom_ic_rhs() its scalar twin for the inner-core rotation,
        wl1 = wimp_lin(1) the implicit weight the matrices were built with, rotate(name) the scheme's rotate_imex."""
        g, N = self.g, self.N
        mats = self._mats[1]
        m0 = self.lm2m == 0

        def per_degree(fn):
            for l in range(self.l_max + 1):
                idx = np.nonzero(self.lm2l == l)[0]
                fn(l, idx)

        # ---- updatePhi (updatePHI.f90:137-297), before updateS (LMLoop.f90:221-229)
        if self.l_phase:
            rhs = rhs_of("phi")
            rhs[0], rhs[-1] = 0.0, 0.0
            rhs[0, self._lm(0, 0)], rhs[-1, self._lm(0, 0)] = self.phi_top, self.phi_bot

            def up_phi(l, idx):
                self.phi[:, idx] = g.solve(mats["phi"][l], rhs[:, idx], (0, N - 1), ("phi", l))
            per_degree(up_phi)
            self.phi[:, m0] = self.phi[:, m0].real
            rotate("phi")
            self._rhs_imp_phi()
        # ---- updateS (updateS.f90:156-342)
        rhs = rhs_of("s")
        if self.l_phase:
            rhs = rhs + self.phase["stef"] * self.phi      # updateS.f90:202-209: St dphi/dt with the new phase field
        rhs[0], rhs[-1] = self.tops, self.bots

        def up_s(l, idx):
            self.s[:, idx] = g.solve(mats["s"][l], rhs[:, idx], (0, N - 1), ("s", l))
        if self.l_heat:
            per_degree(up_s)
            self.s[:, m0] = self.s[:, m0].real
            rotate("s")
            self._rhs_imp_s()
        # ---- updateXi (updateXI.f90:150-330)
        if self.l_chem:
            rhs = rhs_of("xi")
            rhs[0], rhs[-1] = self.topxi, self.botxi

            def up_xi(l, idx):
                self.xi[:, idx] = g.solve(mats["xi"][l], rhs[:, idx], (0, N - 1), ("xi", l))
            per_degree(up_xi)
            self.xi[:, m0] = self.xi[:, m0].real
            rotate("xi")
            self._rhs_imp_xi()
        # ---- updateZ (updateZ.f90:191-488)
        rhs = rhs_of("z")
        if self.prec_fac != 0.0:   # updateZ.f90:385-392: the implicit half of the Poincare force, at the new time
            rhs[:, self._lm(1, 1)] += wl1 * self.prec_fac * (np.sin(self.oek * self.time) - 1j * np.cos(self.oek * self.time))
        rhs[0], rhs[-1] = 0.0, 0.0

        def up_z(l, idx):
            if l == 0:
                self.z[:, idx] = 0.0
            else:
                self.z[:, idx] = g.solve(mats["z"][l], rhs[:, idx], (0, N - 1), ("z", l))
        per_degree(up_z)
        if self.l_rot_ic:     # updateZ.f90:300-356, :416-420: z(1,0) with the inner-core torque balance, then update_rot_rates
            lm10 = self._lm(1, 0)
            dom = dom_ic_rhs()
            rotate("dom_ic")
            if self.kbotv == 2:
                r10 = rhs[:, [lm10]].copy()
                r10[-1] = dom
                self.z[:, [lm10]] = g.solve(mats["z10"], r10, (0, N - 1), ("z10",)).real
                self.omega_ic = self.c_z10_omega_ic * self.z[-1, lm10].real
            else:             # free slip: explicit time stepping of omega_ic (updateZ.f90:1606-1608, gammatau_gravi = 0)
                self.omega_ic = dom / (self.c_lorentz_ic * self.c_moi_ic)
        self.z[:, m0] = self.z[:, m0].real
        rotate("z")
        self._rhs_imp_z()
        # ---- updateWP (updateWP.f90:255-634), buoyancy of the NEW entropy is implicit (:514-524)
        rw = rhs_of("w") + wl1 * (self.rho0 * self.rgrav)[:, None] * (self.BuoFac * self.s + self.ChemFac * self.xi)
        rp = rhs_of("p")
        for a in (rw, rp):
            a[0], a[-1] = 0.0, 0.0

        def up_wp(l, idx):
            if l == 0:
                # updateWP.f90:358-394 with get_p0Mat (:2117-2190; Boussinesq / ThExpNb ViscHeatFac = 0 branch): (d/dr - beta) p00 =
                # rho0 (BuoFac rgrav s00 + ChemFac rgrav xi00) + the explicit term of THIS stage, p00(r_cmb) = 0.  It does not feed
                # back into the flow; the r.m.s. force balance reads it (pressure gradient and buoyancy of the l = 0 mode)
                self.w[:, idx] = 0.0
                ex = self._expl_w_stage
                rhs0 = (self.rho0 * self.rgrav)[:, None] * (self.BuoFac * self.s[:, idx].real + self.ChemFac * self.xi[:, idx].real)
                if ex is not None:
                    rhs0 = rhs0 + ex[:, idx].real
                rhs0[0] = 0.0
                M0 = g.D1 - np.diag(self.beta)
                M0[0] = np.eye(N)[0]
                self.p[:, idx] = g.solve(M0, rhs0.astype(complex), (0,), ("p0",))
                return
            sol = g.solve(mats["wp"][l], np.concatenate([rw[:, idx], rp[:, idx]], axis=0), (0, N - 1, N, 2 * N - 1), ("wp", l))
            self.w[:, idx] = sol[:N]
            self.p[:, idx] = sol[N:]
        per_degree(up_wp)
        self.w[:, m0] = self.w[:, m0].real
        self.p[:, m0] = self.p[:, m0].real
        rotate("w")
        rotate("p")
        self._rhs_imp_wp()
        # ---- updateB (updateB.f90:226-694)
        if self.l_mag:
            rb = rhs_of("b")
            rj = rhs_of("j")
            for a in (rb, rj):
                a[0], a[-1] = 0.0, 0.0

            def up_b(l, idx):
                if l == 0:
                    self.b[:, idx] = 0.0
                    self.aj[:, idx] = 0.0
                    return
                if self.l_cond_ic:
                    self.b[:, idx], self.b_ic[:, idx] = self._solve_coupled(mats["bic"][l], rb[:, idx], rbi[:, idx], key=("bic", l))
                    bc = None if self.aj_nl_icb is None else self.aj_nl_icb[idx]      # updateB.f90:534-539
                    self.aj[:, idx], self.aj_ic[:, idx] = self._solve_coupled(mats["jic"][l], rj[:, idx], rji[:, idx], bc, key=("jic", l))
                    return
                self.b[:, idx] = g.solve(mats["b"][l], rb[:, idx], (0, N - 1), ("b", l))
                self.aj[:, idx] = g.solve(mats["j"][l], rj[:, idx], (0, N - 1), ("j", l))
            if self.l_cond_ic:
                rbi, rji = rhs_of("b_ic"), rhs_of("j_ic")
            per_degree(up_b)
            self.b[:, m0] = self.b[:, m0].real
            self.aj[:, m0] = self.aj[:, m0].real
            rotate("b")
            rotate("j")
            if self.l_cond_ic:
                self.b_ic[:, self.lm2l == 0] = 0.0
                self.aj_ic[:, self.lm2l == 0] = 0.0
                self.b_ic[:, m0] = self.b_ic[:, m0].real
                self.aj_ic[:, m0] = self.aj_ic[:, m0].real
                rotate("b_ic")
                rotate("j_ic")
            self._rhs_imp_b()

    # ------------------------------------------------------------------------------------------------
    def e_kin(self):
        """Columns 2-9 of e_kin.TAG (kinetic_energy.f90:126-196): e_p, e_t, e_p_as, e_t_as, e_p_es, e_t_es, e_p_eas,
        e_t_eas."""
        g = self.g
        m = self.lm2m[None, :]
        dL = self.dL[None, :]
        e_p = self.orho1[:, None] * dL * (dL * g.or2[:, None] * _cc2real(self.w, m) + _cc2real(self.dw, m))
        e_t = self.orho1[:, None] * dL * _cc2real(self.z, m)
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
        return np.array([fac * float(I(c)) for c in cols])


class DirkShellHost(ShellHost):
    """ShellHost advanced by a diagonally implicit IMEX Runge-Kutta scheme without assembly stage (dirk_schemes.f90).  Per
    stage: radial loop on the current fields where the scheme needs an explicit term (l_exp_calc), set_imex_rhs =
    old(1) + dt sum_j a_exp(i+1,j) expl(j) + dt sum_j a_imp(i+1,j) impl(j) (dirk_schemes.f90:856-895), one LM loop with the
    constant diagonal weight, then the implicit term of the new stage state (step_time.f90:396-763)."""

    SCHEMES = {
        # dirk_schemes.f90:723-742
        "BPR353": dict(imp=[[0, 0, 0, 0, 0], [0.5, 0.5, 0, 0, 0], [5.0 / 18.0, -1.0 / 9.0, 0.5, 0, 0], [0.5, 0, 0, 0.5, 0],
                            [0.25, 0, 0.75, -0.5, 0.5]],
                       exp=[[0, 0, 0, 0, 0], [1.0, 0, 0, 0, 0], [4.0 / 9.0, 2.0 / 9.0, 0, 0, 0], [0.25, 0, 0.75, 0, 0],
                            [0.25, 0, 0.75, 0, 0]],
                       l_exp_calc=[True, True, True, False], c=[1.0, 2.0 / 3.0, 1.0, 1.0], wimp=0.5),
    }

    def __init__(self, *a, time_scheme="BPR353", **kw):
        super().__init__(*a, **kw)
        sc = self.SCHEMES[time_scheme]
        self.a_imp, self.a_exp = np.array(sc["imp"], dtype=float), np.array(sc["exp"], dtype=float)
        self.l_exp_calc, self.c_stage, self.wimp0 = sc["l_exp_calc"], sc["c"], sc["wimp"]
        self.nstages = len(self.l_exp_calc)
        self.time_stage = self.time

    def step(self):
        dt = self.dt[0]
        names = [nm for nm in self.old]
        old1 = {nm: self.old[nm].copy() for nm in names}
        impl = {nm: [self.impl[nm].copy()] for nm in names}
        expl = {nm: [] for nm in names}
        d = self.dom_ic
        dom_old1, dom_impl, dom_expl = d["old"], [d["impl"]], []
        wl1 = dt * self.wimp0
        if self._mats is None or self._mats[0] != wl1:
            self._build_mats(wl1)
        t_last = self.time
        for ist in range(1, self.nstages + 1):
            if self.l_exp_calc[ist - 1]:
                out = self.radial_loop({k: np.ascontiguousarray(v) for k, v in self.fields_Rloc().items()})
                self._explicit(out)
                if self.dt[0] > min(self.dtrkc_min, self.dthkc_min):
                    raise RuntimeError("Courant criterion asks for a smaller time step; not expected in this sample")
                for nm in names:
                    expl[nm].append(np.array(self.expl[nm][0]))
                dom_expl.append(d["expl"][0])
            else:
                for nm in names:
                    expl[nm].append(None)
                dom_expl.append(0.0)
            if ist == 1:
                self.time = t_last + dt                                   # step_time.f90:721-722
            self.time_stage = t_last + dt * self.c_stage[ist - 1]         # get_time_stage, dirk_schemes.f90:1075-1085

            def rhs_of(nm, ist=ist):
                r = old1[nm].copy()
                for j in range(ist):
                    if self.a_exp[ist, j] != 0.0:
                        r += dt * self.a_exp[ist, j] * expl[nm][j]
                    if self.a_imp[ist, j] != 0.0:
                        r += dt * self.a_imp[ist, j] * impl[nm][j]
                return r

            def dom_rhs(ist=ist):
                return dom_old1 + dt * sum(self.a_exp[ist, j] * dom_expl[j] + self.a_imp[ist, j] * dom_impl[j] for j in range(ist))

            self._expl_w_stage = expl["w"][ist - 1]   # dwdt%expl(:,:,istage); a stage without explicit evaluation reads zeros
            self._lm_loop(rhs_of, dom_rhs, wl1, lambda nm: None)
            for nm in names:                   # get_*_rhs_imp(..., istage+1): implicit term of the new stage state
                impl[nm].append(self.impl[nm].copy())
            dom_impl.append(d["impl"])
        self.n_steps += 1


class BoussinesqDynamoHost(ShellHost):
    """samples/dynamo_benchmark/input.nml: Boussinesq MHD, rigid insulating walls (the defaults of ShellHost)."""
