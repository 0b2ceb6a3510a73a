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
