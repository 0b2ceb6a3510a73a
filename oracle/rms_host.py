"""numpy restatement of the host side of MagIC's r.m.s. force balance -- TEST INFRASTRUCTURE ONLY.

What the reference does with the spectra of the radial loop's r.m.s. batch (transform_to_lm_RMS, RMS.f90:576-610) on an
lRmsCalc step:

  compute_lm_forces   RMS.f90:612-863   poloidal parts (Coriolis, buoyancy, pressure gradient, Lorentz, advection, inertia) from
                                         the fields, the force balances Geo / Mag / Arc / ArcMag / CLF / PLF / CIA, and the
                                         horizontal integrals per degree (hIntRms, RMS_helpers.f90:154-186)
  init_rNB            RMS.f90:348-467   the radial grid cut back by rCut at either boundary, with its own Chebyshev scheme
  get_force           RMS.f90:865-926   radial integral over the cut grid (rInt_R, integration.f90) per degree, summed, / volume
  dtVrms              RMS.f90:928-1237  the row of dtVrms.TAG (:1178-1186)

The viscous force (DifRms, column 5) does not involve the radial loop: it is formed from w and z by the implicit solver on the last
stage of the preceding step (updateWP.f90:1266-1276 / :1318-1321, updateZ.f90:948-965 with lRmsNext) and completed in dtVrms
(RMS.f90:976-988); restated here so that the whole row can be compared.
Field arrays are [n_r, lm_max] in st_map order; rq is the batch result [14, n_r, lm_max] in the order of include/magic_sht.h.
"""
import numpy as np

from oracle.lmloop import _cc2real


class RmsHost:
    def __init__(self, h, rCut=1e-2, rDea=0.0, l_adv_curl=True):
        self.h, self.l_adv_curl = h, l_adv_curl
        g = h.g
        r, N = g.r, len(g.r)
        l, m = h.lm2l.astype(int), h.lm2m.astype(int)
        self.l, self.m, self.l_max = l, m, int(l.max())
        clm = lambda ll, mm: np.sqrt(((ll + mm) * (ll - mm)) / ((2.0 * ll - 1.0) * (2.0 * ll + 1.0)))   # horizontal.f90:203
        self.dTheta2S = (l - 1.0) * clm(l.astype(float), m.astype(float))                                   # horizontal.f90:222-223
        self.dTheta2A = (l + 2.0) * clm(l + 1.0, m.astype(float))
        self.dPhi = 1j * m
        lm = np.arange(len(l))
        self.lmA = np.where(l < self.l_max, lm + 1, -1)      # (l+1, m): the next entry of the same order in st_map
        self.lmS = np.where(l > m, lm - 1, -1)
        # ---- init_rNB: cut-back grid and its Chebyshev scheme
        nS = None
        for n in range(1, (N - 1) // 2 + 1):
            if r[0] - r[n - 1] >= rCut:
                nS = n
                break
        assert nS is not None
        nS -= 1
        n2 = N - 2 * nS
        allowed = [25, 33, 37, 41, 49, 61, 65, 73, 81, 97, 101, 109, 121, 129, 145, 161, 181, 193, 201, 217, 241, 257, 289, 301, 321, 325,
                   361, 385, 401, 433, 481, 501, 513, 541, 577, 601, 641, 649, 721, 769, 801, 865, 901, 961, 973, 1001, 1025, 1081, 1153,
                   1201, 1281, 1297, 1441, 1501, 1537, 1601, 1621, 1729, 1801, 1921, 1945, 2001, 2049]
        n2 = max(a for a in allowed if a <= n2)
        self.nCut = (N - n2) // 2
        self.n2 = n2
        r2 = r[self.nCut:self.nCut + n2]
        k = np.arange(n2)
        x = np.cos(np.pi * k / (n2 - 1))
        T = np.zeros((n2, n2))
        d1 = np.zeros((n2, n2))
        T[:, 0], T[:, 1], d1[:, 1] = 1.0, x, 1.0
        for n in range(1, n2 - 1):
            T[:, n + 1] = 2 * x * T[:, n] - T[:, n - 1]
            d1[:, n + 1] = 2 * T[:, n] + 2 * x * d1[:, n] - d1[:, n - 1]
        Tinv = np.linalg.inv(T)
        n_cheb2 = min(int((1.0 - rDea) * n2), g.n_cheb_max)
        keep = (np.arange(n2) < n_cheb2).astype(float)
        dr2 = ((d1 * keep[None, :]) @ Tinv) @ r2             # get_dr with drx = 1: dr/dx on the cut grid (drx = 1 / dr2)
        nn = np.arange(n2)
        with np.errstate(divide="ignore"):
            wn = np.where(nn % 2 == 0, 2.0 / (1.0 - nn.astype(float) ** 2), 0.0)
        self.w_cut = (wn @ Tinv) * dr2                       # rInt_R(f, rC, rscheme_RMS) = w_cut @ f
        self.volC = 4.0 / 3.0 * np.pi * (r[self.nCut] ** 3 - r[N - 1 - self.nCut] ** 3)

    # hIntRms, RMS_helpers.f90:154-186: [n_r, lm] -> [n_r, l_max + 1]
    def _hint(self, f, sphertor):
        r2 = self.h.g.r[:, None] ** 2
        help_ = r2 * _cc2real(f, self.m[None, :])
        if sphertor:
            help_ = help_ * (self.l * (self.l + 1.0))[None, :]
        out = np.zeros((f.shape[0], self.l_max + 1))
        for ll in range(self.l_max + 1):
            out[:, ll] = help_[:, self.l == ll].sum(axis=1)
        return out

    def _force(self, h2):          # get_force
        nC, n2 = self.nCut, self.n2
        per_l = self.w_cut @ h2[nC:nC + n2, :]
        return np.sqrt(per_l.sum() / self.volC)

    def row(self, rq, CorFac):
        """One row of dtVrms.TAG from the batch result and the host's present fields."""
        h, g = self.h, self.h.g
        Advr, LFr, dtVr, dpkin, Advt2, Advp2, LFt2, LFp2, CFt2, CFp2, PFt2, PFp2, dtVt, dtVp = rq
        or1, or2, beta = g.or1[:, None], g.or2[:, None], h.beta[:, None]
        l, m = self.l[None, :], self.m[None, :]
        zA = np.where(self.lmA[None, :] >= 0, h.z[:, np.maximum(self.lmA, 0)], 0.0)
        zS = np.where(self.lmS[None, :] >= 0, h.z[:, np.maximum(self.lmS, 0)], 0.0)
        # Coriolis, RMS.f90:640-644 (l = m = 0) and :710-724
        cor_mid = 2.0 * CorFac * or1 * (self.dPhi[None, :] * h.dw + self.dTheta2A[None, :] * zA - self.dTheta2S[None, :] * zS)
        cor_eq = 2.0 * CorFac * or1 * (self.dPhi[None, :] * h.dw + self.dTheta2A[None, :] * zA)
        CorPol = np.where(l == self.l_max, 0.0, np.where(l == m, cor_eq, cor_mid))
        CorPol[:, 0] = 2.0 * CorFac * g.or1 * self.dTheta2A[0] * h.z[:, self.lmA[0]]
        Buo = h.BuoFac * (h.rho0 * h.rgrav)[:, None] * h.s
        dpdr = h.dp
        AdvPol = or2 * Advr
        if self.l_adv_curl:
            AdvPol = AdvPol - dpkin
        LFPol = or2 * LFr
        AdvPol = AdvPol - LFPol
        H, S = (lambda f: self._hint(f, False)), (lambda f: self._hint(f, True))
        pk = dpkin if self.l_adv_curl else 0.0
        Pgrad = dpdr - beta * h.p - pk
        Adv2 = H(AdvPol) + S(Advt2) + S(Advp2)
        Iner2 = H(AdvPol - dtVr) + S(Advt2 - dtVt) + S(Advp2 - dtVp)
        Pre2 = H(Pgrad) + S(PFt2) + S(PFp2)
        Buo2 = H(Buo)
        Cor2 = H(CorPol) + S(CFt2) + S(CFp2)
        LF2 = H(LFPol) + S(LFt2) + S(LFp2)
        Geo = CorPol - dpdr + beta * h.p + pk
        PLF = LFPol - dpdr + beta * h.p + pk
        CLF, Mag = CorPol + LFPol, Geo + LFPol
        Arc, ArcMag = Geo + Buo, Mag + Buo
        CIA = ArcMag + AdvPol - dtVr
        Geo2, CLF2, PLF2, Mag2, Arc2, ArcMag2, CIA2 = (H(x) for x in (Geo, CLF, PLF, Mag, Arc, ArcMag, CIA))
        for CF, PF, LFh, Ad, dV in ((CFt2, PFt2, LFt2, Advt2, dtVt), (CFp2, PFp2, LFp2, Advp2, dtVp)):
            Geo_h = -CF - PF
            Geo2 += S(Geo_h)
            CLF2 += S(-CF + LFh)
            PLF2 += S(LFh - PF)
            Mag2 += S(Geo_h + LFh)
            Arc2 += S(Geo_h)
            ArcMag2 += S(Geo_h + LFh)
            CIA2 += S(Geo_h + LFh + Ad - dV)
        # ---- viscous force (host only)
        dLh = (self.l * (self.l + 1.0))[None, :]
        visc, dLv, dbeta = h.visc[:, None], h.dLvisc[:, None], h.dbeta[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            DifW = dLh * or2 * visc * (g.D2 @ h.w + (2.0 * dLv - beta / 3.0) * h.dw -
                                       (dLh * or2 + 4.0 / 3.0 * (dbeta + dLv * beta + (3.0 * dLv + beta) * or1)) * h.w)
            DifW[:, self.l == 0] = 0.0
            DifPolLMr = np.where(dLh > 0, g.r[:, None] ** 2 / np.where(dLh > 0, dLh, 1.0) * DifW, 0.0)
            DifZ = dLh * or2 * visc * (g.D2 @ h.z + (dLv - beta) * h.dz -
                                       (dLv * beta + 2.0 * dLv * or1 + dLh * or2 + dbeta + 2.0 * beta * or1) * h.z)
            DifZ[:, self.l == 0] = 0.0
            tor = np.where(dLh > 0, g.r[:, None] ** 4 / np.where(dLh > 0, dLh, 1.0) * _cc2real(DifZ, self.m[None, :]), 0.0)
        Dif2 = H(DifW)
        dpol = dLh * _cc2real(g.D1t @ DifPolLMr, self.m[None, :])          # hInt2dPol of d/dr DifPolLMr, RMS.f90:984-987
        for ll in range(self.l_max + 1):
            Dif2[:, ll] += dpol[:, self.l == ll].sum(axis=1) + tor[:, self.l == ll].sum(axis=1)
        F = self._force
        Iner, Cor, LFR, Adv, Buoy, Pre = F(Iner2), F(Cor2), F(LF2), F(Adv2), F(Buo2), F(Pre2)
        Geo_, Mag_, Arc_, ArcMag_, CLF_, PLF_, CIA_ = F(Geo2), F(Mag2), F(Arc2), F(ArcMag2), F(CLF2), F(PLF2), F(CIA2)
        chem = 0.0
        return np.array([h.time, Iner, Cor, LFR, Adv, F(Dif2), Buoy, chem, Pre, Geo_ / (Cor + Pre), Mag_ / (Cor + Pre + LFR),
                         Arc_ / (Cor + Pre + Buoy + chem), ArcMag_ / (Cor + Pre + LFR + Buoy + chem), CLF_ / (Cor + LFR), PLF_ / (Pre + LFR),
                         CIA_ / (Cor + Pre + Buoy + chem + Iner + LFR)])


class DtbHost:
    """What the reference does with the eleven spectra of get_dtBLM (dtB.f90:144-223) on the way to dtBrms.TAG:

      get_dH_dtBLM      dtB.f90:225-337    horizontal-derivative terms: poloidal / toroidal stretching, advection, omega effect
      get_dtBLMfinish   dtB.f90:339-451    + the radial-derivative parts, and the diffusion terms from b, aj
      get_PolTorRms     RMS_helpers.f90:30-100
      dtBrms            RMS.f90:1239-1417  the row of dtBrms.TAG; its first two columns (the time derivative of B, updateB.f90:1638-1643
                                           with lRmsNext) need the field at the start of the step that has just ended
    dtb is the batch result [11, n_r, lm_max] in the order of include/magic_sht.h."""

    def __init__(self, h):
        self.h = h
        g = h.g
        l, m = h.lm2l.astype(int), h.lm2m.astype(int)
        self.l, self.m, self.l_max = l, m, int(l.max())
        clm = lambda ll, mm: np.sqrt(((ll + mm) * (ll - mm)) / ((2.0 * ll - 1.0) * (2.0 * ll + 1.0)))
        lf, mf = l.astype(float), m.astype(float)
        self.dTheta1S = (lf + 1.0) * clm(lf, mf)              # horizontal.f90:219-220
        self.dTheta1A = lf * clm(lf + 1.0, mf)
        lm = np.arange(len(l))
        self.lmA = np.where(l < self.l_max, lm + 1, -1)
        self.lmS = np.where(l > m, lm - 1, -1)
        self.vol_oc = 4.0 / 3.0 * np.pi * (g.r_cmb ** 3 - g.r_icb ** 3)

    def _nb(self, f, idx):          # f at the neighbouring degree (0 where there is none: its coefficient vanishes there)
        return np.where(idx[None, :] >= 0, f[:, np.maximum(idx, 0)], 0.0)

    def _poltor(self, Pol, drPol, Tor):
        """get_PolTorRms: (PolRms, TorRms, PolAsRms, TorAsRms)."""
        g = self.h.g
        dLh = (self.l * (self.l + 1.0))[None, :]
        m = self.m[None, :]
        P = dLh * (dLh * g.or2[:, None] * _cc2real(Pol, m) + _cc2real(drPol, m))
        T = dLh * _cc2real(Tor, m)
        ax = (self.m == 0)
        out = []
        for q in (P.sum(axis=1), T.sum(axis=1), P[:, ax].sum(axis=1), T[:, ax].sum(axis=1)):
            out.append(np.sqrt(g.rInt_R(q) / self.vol_oc))
        return out

    def row(self, dtb, b_start, aj_start, dt):
        h, g = self.h, self.h.g
        BtVr, BpVr, BrVt, BrVp, BtVp, BpVt, Cot, Sn2, BrVZ, BtVZ, _ = dtb
        or1, or2 = g.or1[:, None], g.or2[:, None]
        or3 = or1 * or2
        dLh = (self.l * (self.l + 1.0))[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            fac = np.where(dLh > 0, or2 / np.where(dLh > 0, dLh, 1.0), 0.0)
        not0 = (self.l > 0)[None, :]
        dPhi2 = -(self.m.astype(float) ** 2)[None, :]                       # dPhi(lm)**2 = (i m)^2
        d1S, d1A = self.dTheta1S[None, :], self.dTheta1A[None, :]
        cotS, cotA = self._nb(Cot, self.lmS), self._nb(Cot, self.lmA)
        vzS, vzA = self._nb(BrVZ, self.lmS), self._nb(BrVZ, self.lmA)
        # get_dH_dtBLM
        Pstr, Padv = not0 * or2 * BtVr, not0 * or2 * BrVt
        TstrR, TadvR = not0 * or1 * BrVp, not0 * or2 * BpVr
        hor = fac * (d1S * cotS - d1A * cotA) - fac * dPhi2 * Sn2
        Tstr = not0 * (-or2 * BtVp + or3 * BpVr + hor)
        Tadv = not0 * (-or2 * BpVt + or3 * BrVp + hor)
        TomeR = not0 * fac * (d1S * vzS - d1A * vzA)
        Tome = not0 * (-or2 * BtVZ) - or1 * TomeR
        # get_dtBLMfinish
        D1 = g.D1t
        Tome, Tstr, Tadv = Tome + or1 * (D1 @ TomeR), Tstr + or1 * (D1 @ TstrR), Tadv + or1 * (D1 @ TadvR)
        lam = h.lam[:, None] if hasattr(h, "lam") else 1.0
        dLlam = h.dLlam[:, None] if hasattr(h, "dLlam") else 0.0
        Pdif = h.opm * lam * (h.ddb - dLh * or2 * h.b)
        Tdif = h.opm * lam * (g.D2 @ h.aj + dLlam * h.dj - dLh * or2 * h.aj)
        # dtBrms
        Pdyn, drPdyn, Tdyn = Pstr - Padv, D1 @ Pstr - D1 @ Padv, Tstr - Tadv
        PdynRms, TdynRms, _, _ = self._poltor(Pdyn, drPdyn, Tdyn)
        drPdif = D1 @ Pdif
        PdifRms, TdifRms, _, _ = self._poltor(Pdif, drPdif, Tdif)
        _, TomeRms, _, TomeAsRms = self._poltor(Pdif, drPdif, Tome)
        dip = ((self.l == 1) & (self.m <= 1))[None, :]
        DdynRms, _, DdynAsRms, _ = self._poltor(np.where(dip, Pdyn, 0.0), np.where(dip, drPdyn, 0.0), Tdyn)
        # time derivative of the field over the step that has just ended (updateB.f90:1638-1643)
        dtP = dLh * or2 / dt * (h.b - b_start)
        dtT = dLh * or2 / dt * (h.aj - aj_start)
        with np.errstate(divide="ignore", invalid="ignore"):
            dtBPolLMr = np.where(dLh > 0, g.r[:, None] ** 2 / np.where(dLh > 0, dLh, 1.0) * dtP, 0.0)
            tor = np.where(dLh > 0, g.r[:, None] ** 4 / np.where(dLh > 0, dLh, 1.0) * _cc2real(dtT, self.m[None, :]), 0.0)
        pol = g.r[:, None] ** 2 * _cc2real(dtP, self.m[None, :]) + dLh * _cc2real(D1 @ dtBPolLMr, self.m[None, :])
        dtBPolRms = np.sqrt(g.rInt_R((not0 * pol).sum(axis=1)) / self.vol_oc)
        dtBTorRms = np.sqrt(g.rInt_R((not0 * tor).sum(axis=1)) / self.vol_oc)
        return np.array([h.time, dtBPolRms, dtBTorRms, PdynRms, TdynRms, PdifRms, TdifRms, TomeRms / TdynRms, TomeAsRms / TdynRms, DdynRms,
                         DdynAsRms])


def simps(f, r):
    """integration.f90 simps: Simpson's rule on a non-uniform, DEcreasing grid."""
    f, r = np.asarray(f, dtype=float), np.asarray(r, dtype=float)
    n = len(f)

    def seg(i):   # 0-based centre i: points i-1, i, i+1
        h2, h1 = r[i + 1] - r[i], r[i] - r[i - 1]
        return (h1 + h2) / 6.0 * (f[i - 1] * (2.0 * h1 - h2) / h1 + f[i] * (h1 + h2) * (h1 + h2) / (h1 * h2) + f[i + 1] * (2.0 * h2 - h1) / h2)
    if n % 2 == 1:
        return -sum(seg(i) for i in range(1, n - 1, 2))
    tot = 0.5 * (r[1] - r[0]) * (f[1] + f[0])
    tot += sum(seg(i) for i in range(2, n - 1, 2))
    tot += 0.5 * (r[n - 1] - r[n - 2]) * (f[n - 1] + f[n - 2])
    tot += sum(seg(i) for i in range(1, n - 1, 2))
    return -0.5 * tot


class ToHost:
    """What the reference does with getTO's (r, theta) arrays on the way to Tay.TAG (out_TO.f90:213-556): the z-averages on a
    cylindrical grid (cylmean_otc / cylmean_itc, integration.f90:157-533: fourth-order Lagrange interpolation in r and theta,
    Simpson in z), the Taylorisation measures and the geostrophic flow energy; and getTOfinish's axisymmetric viscous stress
    (TO.f90:354-390, get_PAS :392-425).  to is the batch result [n_r, 15, n_theta] in the order of include/magic_sht.h."""

    def __init__(self, h, theta_ord, toraxi_to_spat, sDens=1.0, zDens=1.0):
        self.h, self.theta, self.toraxi_to_spat, self.zDens = h, np.asarray(theta_ord, dtype=float), toraxi_to_spat, zDens
        g = h.g
        N = len(g.r)
        self.n_s_max = int(sDens * (N + int(g.r_icb * N)))                      # preCalculations.f90:768-769
        smin, smax = g.r_cmb * np.sin(self.theta[0]), g.r_cmb
        ds = (smax - smin) / (self.n_s_max - 1)
        self.cyl = g.r_cmb - np.arange(self.n_s_max) * ds                       # out_TO.f90:98-103
        self.hh = np.zeros(self.n_s_max)
        for i, s in enumerate(self.cyl):
            if s >= g.r_icb:
                self.hh[i] = 2.0 * np.sqrt(g.r_cmb ** 2 - s ** 2)
                self.n_s_otc = i + 1                                            # last index (1-based) outside the tangent cylinder
            else:
                self.hh[i] = np.sqrt(g.r_cmb ** 2 - s ** 2) - np.sqrt(g.r_icb ** 2 - s ** 2)
        k = self.n_s_otc
        self.volcyl_oc = 2.0 * np.pi * (simps(self.hh * self.cyl, self.cyl) + simps(self.hh[k:] * self.cyl[k:], self.cyl[k:]))
        self.vol_oc = 4.0 / 3.0 * np.pi * (g.r_cmb ** 3 - g.r_icb ** 3)

    def _interp(self, a, rc, thet, south):
        """The Lagrange interpolation shared by cylmean_otc and cylmean_itc; a is [n_theta, n_r] (ordered colatitudes)."""
        r, theta = self.h.g.r, self.theta
        n_r_max, n_theta_max = len(r), len(theta)
        n_r2 = n_r_max - 1                       # 1-based indices as in the reference
        for n_r in range(n_r_max - 1, 0, -1):
            if r[n_r - 1] >= rc:
                n_r2 = n_r
                break
        if n_r2 == n_r_max - 1:
            n_r2 = n_r_max - 2
        if n_r2 == 1:
            n_r2 = 2
        n_r3, n_r1, n_r0 = n_r2 - 1, n_r2 + 1, n_r2 + 2
        if thet < theta[0]:
            n_th1 = 1
        else:
            n_th1 = n_theta_max
            for n_th in range(n_theta_max, 0, -1):
                if theta[n_th - 1] <= thet:
                    n_th1 = n_th
                    break
        if n_th1 == n_theta_max:
            n_th1 = n_theta_max - 2
        if n_th1 == n_theta_max - 1:
            n_th1 = n_theta_max - 2
        if n_th1 == 1:
            n_th1 = 2
        n_th2, n_th3, n_th0 = n_th1 + 1, n_th1 + 2, n_th1 - 1
        R = lambda i: r[i - 1]
        T = lambda i: theta[i - 1]
        rr0, rr1, rr2, rr3 = rc - R(n_r0), rc - R(n_r1), rc - R(n_r2), rc - R(n_r3)
        r10, r20, r30 = 1.0 / (R(n_r1) - R(n_r0)), 1.0 / (R(n_r2) - R(n_r0)), 1.0 / (R(n_r3) - R(n_r0))
        r21, r31, r32 = 1.0 / (R(n_r2) - R(n_r1)), 1.0 / (R(n_r3) - R(n_r1)), 1.0 / (R(n_r3) - R(n_r2))
        tt0, tt1, tt2, tt3 = thet - T(n_th0), thet - T(n_th1), thet - T(n_th2), thet - T(n_th3)
        t10, t20, t30 = 1.0 / (T(n_th1) - T(n_th0)), 1.0 / (T(n_th2) - T(n_th0)), 1.0 / (T(n_th3) - T(n_th0))
        t21, t31, t32 = 1.0 / (T(n_th2) - T(n_th1)), 1.0 / (T(n_th3) - T(n_th1)), 1.0 / (T(n_th3) - T(n_th2))
        ait = [0.0] * 4
        for itr in range(4):
            n_th = n_th0 + itr
            if south:
                n_th = n_theta_max + 1 - n_th
            A = lambda nr: a[n_th - 1, nr - 1]
            a01 = (rr0 * A(n_r1) - rr1 * A(n_r0)) * r10
            a12 = (rr1 * A(n_r2) - rr2 * A(n_r1)) * r21
            a23 = (rr2 * A(n_r3) - rr3 * A(n_r2)) * r32
            a012 = (rr0 * a12 - rr2 * a01) * r20
            a123 = (rr1 * a23 - rr3 * a12) * r31
            ait[itr] = (rr0 * a123 - rr3 * a012) * r30
        a01 = (tt0 * ait[1] - tt1 * ait[0]) * t10
        a12 = (tt1 * ait[2] - tt2 * ait[1]) * t21
        a23 = (tt2 * ait[3] - tt3 * ait[2]) * t32
        a012 = (tt0 * a12 - tt2 * a01) * t20
        a123 = (tt1 * a23 - tt3 * a12) * t31
        return (tt0 * a123 - tt3 * a012) * t30

    def cylmean(self, a):
        """out_TO.f90 cylmean: z-averages (north, south) of a[n_theta, n_r] on the cylindrical grid."""
        g = self.h.g
        r_cmb, r_icb = g.r[0], g.r[-1]
        eps = 10.0 * np.finfo(float).eps
        n_s_max, n_theta_max = self.n_s_max, len(self.theta)
        vN, vS = np.zeros(n_s_max), np.zeros(n_s_max)

        def angle(s, z, rc):
            if s == 0.0 or z > rc:
                return 0.0
            return np.arccos(z / rc)
        for n_s in range(1, self.n_s_otc + 1):                      # cylmean_otc
            s = self.cyl[n_s - 1]
            zmax, zmin = np.sqrt(r_cmb * r_cmb - s * s), 0.0
            nz = 2 * int(n_s_max * (zmax - zmin) / (2.0 * r_cmb))
            nz = max(int(self.zDens * nz), 4)
            dz = (zmax - zmin) / nz
            ac = {}
            for n_z in range(-nz, nz + 1):
                z = zmin + dz * n_z
                rc = np.sqrt(s * s + z * z)
                if rc >= r_cmb:
                    rc = r_cmb - eps
                ac[n_z] = self._interp(a, rc, angle(s, z, rc), False)
            tot = ac[-nz] + ac[nz]
            tot += sum(4.0 * ac[n_z] for n_z in range(-nz + 1, nz, 2))
            tot += sum(2.0 * ac[n_z] for n_z in range(-nz + 2, nz - 1, 2))
            vN[n_s - 1] = vS[n_s - 1] = tot / (6.0 * nz)
        vN[0] = vS[0] = 0.5 * (a[n_theta_max // 2 - 1, 0] + a[n_theta_max // 2, 0])
        for n_s in range(self.n_s_otc + 1, n_s_max + 1):            # cylmean_itc
            s = self.cyl[n_s - 1]
            zmax, zmin = np.sqrt(r_cmb * r_cmb - s * s), np.sqrt(r_icb * r_icb - s * s)
            nz = 2 * int(n_s_max * (zmax - zmin) / (2.0 * r_cmb))
            nz = max(int(self.zDens * nz), 4)
            dz = (zmax - zmin) / nz
            acn, acs = [], []
            for n_z in range(0, nz + 1):
                z = zmin + dz * n_z
                rc = np.sqrt(s * s + z * z)
                if rc >= r_cmb:
                    rc = r_cmb - eps
                if rc <= r_icb:
                    rc = r_icb + eps
                th = angle(s, z, rc)
                acn.append(self._interp(a, rc, th, False))
                acs.append(self._interp(a, rc, th, True))
            for ac, v in ((acn, vN), (acs, vS)):
                tot = ac[0] + ac[nz]
                tot += sum(4.0 * ac[n_z] for n_z in range(1, nz, 2))
                tot += sum(2.0 * ac[n_z] for n_z in range(2, nz - 1, 2))
                v[n_s - 1] = tot / (3.0 * nz)
        return vN, vS

    def dzStrAS(self):
        """getTOfinish's viscous stress of the axisymmetric toroidal flow, TO.f90:372-386 + get_PAS: [n_theta (ordered), n_r]."""
        h, g = self.h, self.h.g
        m0 = np.nonzero(h.lm2m == 0)[0]
        l = h.lm2l[m0].astype(float)
        z, dz, ddz = h.z[:, m0].real, h.dz[:, m0].real, (g.D2 @ h.z)[:, m0].real
        dLh = l * (l + 1.0)
        out = np.zeros((len(self.theta), len(g.r)))
        n = len(self.theta)
        ordered = np.array([t // 2 if t % 2 == 0 else n - 1 - t // 2 for t in range(n)])
        for i in range(len(g.r)):
            strl = ddz[i] - h.beta[i] * dz[i] - (dLh * g.or2[i] + h.dbeta[i] + 2.0 * h.beta[i] * g.or1[i]) * z[i]
            strl[l == 0] = 0.0
            _, tmpp = self.toraxi_to_spat(strl.astype(complex), int(l.max()))
            sin_scr = np.zeros(n)
            sin_scr[:] = np.sin(self.theta[ordered])
            out[ordered, i] = tmpp / sin_scr * g.or1[i]
        return out

    def cyl_means(self, to, with_hemi_files=False):
        """The z-averages outTO forms (out_TO.f90:278-350): name -> (north, south) on the cylindrical grid.  with_hemi_files adds
        what only TOnhs / TOshs.TAG hold (:477-507): the axisymmetric stress and the relative geostrophic flow VpRInt."""
        VAS, dzRstr, dzLF = (np.ascontiguousarray(to[:, q, :].T) for q in (1, 3, 5))
        dzStr = self.dzStrAS()
        c = {"Vp": self.cylmean(VAS), "LF": self.cylmean(dzLF), "Rstr": self.cylmean(dzRstr), "Str": self.cylmean(dzStr)}

        def ratio(num, den):      # out_TO.f90:338-349 (the test is on the northern value for both hemispheres)
            return np.where(np.abs(den[0]) > 0.0, num[0] / np.where(den[0] != 0, den[0], 1.0), den[0]), \
                   np.where(np.abs(den[0]) > 0.0, num[1] / np.where(den[1] != 0, den[1], 1.0), den[1])
        c["Tay"] = ratio(c["LF"], self.cylmean(np.abs(dzLF)))
        c["TayR"] = ratio(c["Rstr"], self.cylmean(np.abs(dzRstr)))
        c["TayV"] = ratio(c["Str"], self.cylmean(np.abs(dzStr)))
        if with_hemi_files:
            c["Astr"] = self.cylmean(np.ascontiguousarray(to[:, 4, :].T))
            V2 = self.cylmean(np.ascontiguousarray(to[:, 0, :].T))
            vpr = []
            for vp, v2 in zip(c["Vp"], V2):          # out_TO.f90:318-335
                with np.errstate(divide="ignore", invalid="ignore"):
                    q = np.where(v2 < 0.0, 1.0, np.where(np.abs(v2) <= 10.0 * np.finfo(float).eps, 0.0, np.abs(vp) / np.sqrt(np.abs(v2))))
                vpr.append(np.minimum(1.0, q))
            c["VpR"] = tuple(vpr)
        return c

    def row(self, to, e_kin_cols, means=None):
        """One row of Tay.TAG (out_TO.f90:515-553)."""
        h = self.h
        c = means if means is not None else self.cyl_means(to)
        (VpN, VpS), (TayN, TayS), (TayRN, TayRS), (TayVN, TayVS) = c["Vp"], c["Tay"], c["TayR"], c["TayV"]
        cyl, hh, k = self.cyl, self.hh, self.n_s_otc

        def integ(fn, fs):
            return simps(fn * cyl * hh, cyl) + simps(fs[k:] * cyl[k:] * hh[k:], cyl[k:])
        VgRMS = np.sqrt(2.0 * np.pi * integ(VpN * VpN, VpS * VpS) / self.volcyl_oc)
        TayRMS, TayRRMS, TayVRMS = (2.0 * np.pi * integ(np.abs(a), np.abs(b)) / self.volcyl_oc
                                    for a, b in ((TayN, TayS), (TayRN, TayRS), (TayVN, TayVS)))
        eKin, eKinTAS = e_kin_cols[0] + e_kin_cols[1], e_kin_cols[3]
        VRMS, VpRMS = np.sqrt(2.0 * eKin / self.vol_oc), np.sqrt(2.0 * eKinTAS / self.vol_oc)
        if VRMS != 0.0:
            VpRMS, VgRMS = VpRMS / VRMS, VgRMS / VRMS
        return np.array([h.time, VpRMS ** 2, VgRMS ** 2, TayRMS, TayRRMS, TayVRMS, eKin])
