"""Beta-decay rates from FAM strengths: lepton phase space, phase-space weighted shape
factor, contour integration -> rates and half-lives (the content of pynfam's beta.out).

Host-side mirror of the reference's Python for this step -- same class / method names, settings keys and results:

    phaseSpace    pynfam/strength/phase_space.py:26-556   (beta-minus / beta-plus at zero temperature; RATINT continuation)
    shapeFactor   pynfam/strength/shape_factor.py:27-1031 (prepConstants, calcPhaseSpace, prepStrengths, calcSfByJ,
                                                           calcSfContributions, calcShapeFactor, calcBetaRates)
    Fermi, lambda_ke, L0, thieleInterpolator ...          phase_space.py:557-982

after Mustonen et al., Phys. Rev. C 90, 024308 (2014).  Inputs are `strength.famStrength` objects (filled by the batched
GPU solve or read from OP.out.ctr files) and the two HFB numbers pynfam takes from the HFB log (HFB_Qval, EQRPA_max).
Not built (raises): electron capture ('c'), finite temperature, the POLYFIT continuation, zeroed negative strength.
Pinned by tests/test_rates.py against the reference's own phasespace / shapefactor / beta.out files.
"""
import numpy as np
from scipy.special import factorial2, loggamma

from .strength import KAPPA, complex_quadrature, famStrength

# pynfam/config.py:219-233
ALPHA = 7.2973525698e-03    # fine-structure constant
HBAR_MEC = 386.15926800     # (hbar c)/(me c^2) [fm]
MEC2 = 0.510998928          # me c^2 [MeV]
R0 = 1.2                    # nuclear radius parameter [fm]
MN = 939.0                  # nucleon mass [MeV]
EPSILON = np.finfo(np.float64).eps

_PSI_DEFAULTS = {"psi_approx": "RATINT", "psi_glpts": 15, "screening": None, "Q_eff": None, "Q_eff_mode": 0,
                 "GA": -1.27, "GV": 1.0, "ratint_pts": 20}      # config.py:199-212


# ---- numerical helpers (phase_space.py:557-708) -----------------------------------------------------------------------
def transform_interval(x, wts, xmin, xmax):
    """Gauss nodes / weights on [-1, 1] -> [xmin, xmax]."""
    half = 0.5 * (xmax - xmin)
    return x * half + 0.5 * (xmin + xmax), wts * half


def thieleInterpolator(x, y):
    """Rational interpolant through (x, y) as a Thiele continued fraction (reciprocal differences); the returned
    function accepts complex arguments -- this is how the phase-space integrals reach the complex contour."""
    n = len(y)
    rho = [[y[i]] * (n - i) for i in range(n)]
    for i in range(n - 1):
        rho[i][1] = (x[i] - x[i + 1]) / (rho[i][0] - rho[i + 1][0] + 1e-15)
    for order in range(2, n):
        for j in range(n - order):
            rho[j][order] = (x[j] - x[j + order]) / (rho[j][order - 1] - rho[j + 1][order - 1]) + rho[j + 1][order - 2]
    c = rho[0]

    def t(xin):
        a = 0
        for i in range(n - 1, 1, -1):
            a = (xin - x[i - 1]) / (c[i] - c[i - 2] + a)
        return y[0] + (xin - x[0]) / (c[1] + a)
    return t


# ---- Coulomb functions (phase_space.py:711-982) ----------------------------------------------------------------------
def V0_shift(Zd):
    """Screening shift of the electron energy, N(Z) alpha^2 |Z-1|^(4/3) with N = 1.43."""
    return 1.43 * ALPHA ** 2 * abs(Zd - 1) ** (4.0 / 3.0)


def w_screen(Zd, w):
    return np.asarray(w) - V0_shift(Zd)


def mu_ke(ke):
    return 1.0


def gam_ke(ke, Zd):
    return np.sqrt(float(ke) ** 2 - (ALPHA * Zd) ** 2)


def L0(Zd):
    return 0.5 * (1.0 + gam_ke(1, Zd))


def lambda_ke_F0Fk(ke, Zd):
    return (ke + gam_ke(ke, Zd)) / (ke * (1 + gam_ke(1, Zd)))


def Fermi(F, Zd, A, w, sc=False):
    """Fermi function F_F (without L0) for real electron energies w, through ln F to dodge over/underflow:
    F = [(F+1)(2F+1)!!]^2 4^(F+1) (2pR)^(2(gamma-F-1)) exp(pi y) |Gamma(gamma + i y)|^2 / Gamma(2 gamma + 1)^2,
    y = alpha Z w / p; zero where p -> 0.  Rose screening (sc) evaluates at w - V0 with the prefactor (p~/p)(w~/w)."""
    w = np.asarray(w, dtype=float)
    if sc:
        w_us, w = w, w_screen(Zd, w)
    ok = ~((np.abs(w - 1.0) < EPSILON) | (w < 1.0))
    out = np.zeros_like(w)
    ww = w[ok]
    p = np.sqrt(ww ** 2 - 1.0)
    pref = 1.0 if not sc else (p / np.sqrt(w_us[ok] ** 2 - 1.0)) * (ww / w_us[ok])
    R = R0 / HBAR_MEC * A ** (1.0 / 3.0)
    g = gam_ke(F + 1, Zd)
    factor = ((F + 1) * factorial2(2 * (F + 1) - 1)) ** 2 * 4 ** (F + 1) * (2.0 * p * R) ** (2.0 * (g - F - 1))
    y = ALPHA * Zd * ww / p
    lg = loggamma(g + 1j * y)
    ln_f = np.log(pref) + np.log(factor) + np.pi * y + lg + np.conj(lg) - 2 * loggamma(2.0 * g + 1.0)
    out[ok] = np.exp(np.real(ln_f))
    return out


def lambda_ke(ke, Zd, A, w, sc=False):
    """lambda_k = (k + gamma_k) / (k (1 + gamma_1)) F_{k-1} / F_0, zero where p -> 0."""
    w = np.asarray(w, dtype=float)
    ws = w_screen(Zd, w) if sc else w
    ok = ~((np.abs(ws - 1.0) < EPSILON) | (ws <= 1.0))
    out = np.zeros_like(w)
    out[ok] = lambda_ke_F0Fk(ke, Zd) * Fermi(ke - 1, Zd, A, w[ok], sc) / Fermi(0, Zd, A, w[ok], sc)
    return out


# ---- phase space -------------------------------------------------------------------------------------------------------
class phaseSpace(object):
    """Lepton phase-space integrals f_1..f_6 of allowed / first-forbidden beta decay (phase_space.py:26-100)."""

    def __init__(self, beta):
        if beta not in ("-", "+", "c"):
            raise ValueError("Invalid beta decay type requested in phaseSpace.")
        if beta == "c":
            raise NotImplementedError("electron capture is outside this path")
        self.beta = beta
        self._settings = dict(_PSI_DEFAULTS)

    approx = property(lambda s: s._settings["psi_approx"])
    GA = property(lambda s: s._settings["GA"])
    GV = property(lambda s: s._settings["GV"])
    LAM = property(lambda s: abs(s._settings["GA"] / s._settings["GV"]))

    def updateSettings(self, override):
        for h in override:
            if h not in self._settings:
                raise KeyError("Invalid override setting '{:}' for phase space approx '{:}'.".format(h, self.approx))
            self._settings[h] = override[h]

    @staticmethod
    def psi_integrand(wx, n, Zd, A, w0, sc=False):
        """p (W0 - w)^2 g_n(w) F_0(Z, w) L_0 (phase_space.py:253-312, zero-temperature beta decay branch)."""
        wx = np.array(wx, dtype=float)
        wx[wx < 1.0] = 1.0
        px = np.sqrt(wx ** 2 - 1.0)
        pv = w0 - wx
        if n == 1:
            gn = gam_ke(1, Zd) * mu_ke(1)
        elif n == 2:
            gn = wx
        elif n == 3:
            gn = wx * wx
        elif n == 4:
            gn = wx * wx * wx
        elif n == 5:
            gn = wx * pv * pv
        elif n == 6:
            gn = wx * px * px * lambda_ke(2, Zd, A, wx, sc)
        else:
            raise ValueError("Phase space factor n takes values 1 to 6.")
        f = px * pv * pv * gn * Fermi(0, Zd, A, wx, sc) * L0(Zd)
        f[~np.isfinite(f)] = 0
        return f

    def calcPsi(self, n, Zd, A, w0, sc=False):
        """f_n(W0) = integral from 1 to W0 on psi_glpts Gauss-Legendre nodes, real W0 only (phase_space.py:316-356)."""
        w0 = np.asarray(w0)
        if n not in range(1, 7):
            raise ValueError("Phase space factor n takes values 1 to 6.")
        if not np.isreal(w0).all():
            raise ValueError("calcPsi got complex argument w0.")
        x, wt = np.polynomial.legendre.leggauss(self._settings["psi_glpts"])

        def one(wmax):
            w, g = transform_interval(x, wt, 1.0, wmax)
            return np.sum(self.psi_integrand(w, n, Zd, A, wmax, sc=sc) * g)
        return np.vectorize(one)(np.real(w0))

    def analyticPsi(self, fpsi, w0min, w0max):
        """Continuation of f_n off the real axis: Thiele interpolant through ratint_pts Chebyshev nodes of
        [w0min, w0max] (phase_space.py:103-146)."""
        if self.approx != "RATINT":
            raise NotImplementedError("only the RATINT continuation is built")
        x, cw = np.polynomial.chebyshev.chebgauss(self._settings["ratint_pts"])
        x, _ = transform_interval(x, cw, w0min, w0max)
        return thieleInterpolator(x, fpsi(x))

    def psiFct(self, n, Zd, A, eqrpamax, eqrpamin, approx=True):
        """f_n as a function of W0 = 1 + (EQRPA_max - EQRPA)/me c^2 on [EQRPA_min, EQRPA_max]; beta-plus takes -|Z| and
        screening (phase_space.py:150-195)."""
        if self.beta == "-":
            Zd, sc = abs(Zd), False
        else:
            Zd, sc = -abs(Zd), True
        if self._settings["screening"] is not None:
            sc = self._settings["screening"]

        def fpsi(w0):
            return self.calcPsi(n, Zd, A, w0, sc=sc)
        if not approx:
            return fpsi
        return self.analyticPsi(fpsi, 1.0, 1.0 + (eqrpamax - eqrpamin) / MEC2)


# ---- shape factor -------------------------------------------------------------------------------------------------------
class shapeFactor(phaseSpace):
    """Phase-space weighted shape factor of allowed + first-forbidden decay from a set of famStrength objects, and its
    integral (shape_factor.py:27-166).  strengths: list of famStrength on one contour; beta_type '-' or '+'."""

    beta_totals = ["Total", "Total-Allowed", "Total-GT", "Total-Forbidden"]
    beta_al_contribs = ["Allowed-Fermi", "Allowed-GT_K=0", "Allowed-GT_K=1"]
    beta_ffK_totals = ["Forbidden-K=0", "Forbidden-K=1", "Forbidden-K=2"]
    beta_ffJ_totals = ["Forbidden-J=0", "Forbidden-J=1", "Forbidden-J=2"]
    beta_ff_contribs = ["Forbidden-(J,K)=(0,0)", "Forbidden-(J,K)=(1,0)", "Forbidden-(J,K)=(1,1)",
                        "Forbidden-(J,K)=(2,0)", "Forbidden-(J,K)=(2,1)", "Forbidden-(J,K)=(2,2)"]
    betaout_keys = beta_totals + beta_al_contribs + beta_ffK_totals + beta_ffJ_totals + beta_ff_contribs
    al_ops = ["F", "GT"]
    ff_ops = ["P", "R", "PS0", "RS0", "RS1", "RS2"]
    xts = ["RS0_PS0", "R_RS1", "P_RS1", "R_P"]
    CJ0_keys = ["J0_R1", "J0_R2"]
    CJ1_keys = ["J1_R1", "J1_R2", "J1_R3", "J1_R4", "J1_R5", "J1_R6"]
    CJ2_keys = ["J2_R5", "J2_R6"]
    C_keys = CJ0_keys + CJ1_keys + CJ2_keys

    def __init__(self, strengths, beta_type, ps_contour=None):
        if not isinstance(strengths, list):
            strengths = [strengths]
        if not all(isinstance(s, famStrength) for s in strengths):
            raise ValueError("Strength inputs must be famStrength objects.")
        phaseSpace.__init__(self, beta_type)
        s0 = strengths[0]
        dim = s0.contour.nr_points
        for s in strengths:
            if s.nucleus != s0.nucleus or s.str_df.shape[0] != dim or len(s.contour.ctr_z) != dim or s.beta != s0.beta:
                raise ValueError("inconsistent strengths")
        if s0.beta != beta_type:
            raise ValueError("beta type of the strengths and of the phase space differ")
        self.strengths = {s.genopname: s for s in strengths}
        self.contour = s0.contour
        self.ps_contour = ps_contour or self.contour
        self.nucleus = s0.nucleus
        self.sf_metadict = {"beta_type": self.beta, "quadratr": None, "FAM_ctr": self.contour.name_and_int}
        self.sf_df = self.ps_df = None

    # ---- constants (shape_factor.py:359-435) ---------------------------------------------------------------------------
    def prepConstants(self, hfb_gs):
        """hfb_gs: dict with 'HFB_Qval' and 'EQRPA_max' (what pynfam reads from the HFB log)."""
        qval, eqrpamax = float(hfb_gs["HFB_Qval"]), float(hfb_gs["EQRPA_max"])
        A, Zi = self.nucleus[2], self.nucleus[1]
        rad = R0 * A ** (1.0 / 3.0) / HBAR_MEC
        if self.beta == "-":
            Zf, Zd, LAMd = Zi + 1, Zi + 1, self.LAM
        else:
            Zf, Zd, LAMd = Zi - 1, -(Zi - 1), -self.LAM
        q_eff, q_mode = self._settings["Q_eff"], self._settings["Q_eff_mode"]
        if q_eff is not None and q_mode != 0:
            egs = eqrpamax - qval
            qval = q_eff if abs(q_mode) == 1 else qval + q_eff
            eqrpamax = qval + egs
        self.sf_metadict.update({"ec": +1, "Z": Zd, "LAM": LAMd, "ft_active": False, "temper": 0.0, "HFB_Qval": qval,
                                 "EQRPAmax": eqrpamax, "Half_Width": None if self.contour.closed else self.contour.half_width,
                                 "A": A, "Zi": Zi, "Zf": Zf, "|gA|/gV": self.LAM, "gA": self.GA, "gV": self.GV,
                                 "M_nucleon": MN, "Radius": rad, "alpha*Z": ALPHA * abs(Zd),
                                 "alpha*Z/2R": ALPHA * abs(Zd) / (2.0 * rad), "W0_max": 1.0 + qval / MEC2,
                                 "W0*R": (1.0 + qval) / MEC2 * rad})

    # ---- phase space on the contour (shape_factor.py:438-524) --------------------------------------------------------
    def calcPhaseSpace(self):
        m = self.sf_metadict
        eqrpa = self.contour.ctr_z if self.contour.closed else np.real(self.contour.ctr_z)
        w0 = (m["EQRPAmax"] - eqrpa) / MEC2 + 1.0
        ps = {}
        for n in range(1, 7):
            if m["HFB_Qval"] < 0:
                ps["f%d" % n] = np.zeros(len(eqrpa))
            else:
                fct = self.psiFct(n, m["Z"], m["A"], m["EQRPAmax"], self.ps_contour.energy_min, approx=self.contour.closed)
                ps["f%d" % n] = fct(w0)
        self.ps_df, self.ps_w0 = ps, w0
        # the phase-space settings travel in the header of the output files
        self.sf_metadict.update({k: v for k, v in self._settings.items() if k not in ("GA", "GV")})
        self.sf_metadict.update({"E_1stPeak": None, "FAM_Qval": None})

    # ---- strengths with their dimensionful prefactors (shape_factor.py:527-646) ----------------------------------------
    def prepStrengths(self):
        rt2, h = np.sqrt(2.0), HBAR_MEC
        pm = -h * MEC2 / MN
        pre = {"F_K0": 1.0, "GT_K0": 1.0, "GT_K1": rt2, "P_K0": pm, "P_K1": rt2 * pm, "R_K0": np.sqrt(3.0) / h,
               "R_K1": rt2 * np.sqrt(3.0) / h, "PS0_K0": pm, "RS0_K0": 1.0 / h, "RS1_K0": -1.0 / h, "RS1_K1": -rt2 / h,
               "RS2_K0": 1.0 / h, "RS2_K1": rt2 / h, "RS2_K2": rt2 / h}
        dim = len(self.contour.ctr_z)
        b = {o: {k: np.zeros(dim) for k in range(3)} for o in self.al_ops + self.ff_ops + self.xts}

        def check(sa, sb, label):
            """the cross-term computed from either operator must agree (tolerances of the reference)"""
            other = "x".join(reversed(label.split("x")))
            d = sa.cstr_df[label].values - sb.cstr_df[other].values
            for part, v in (("Real", np.abs(d.real)), ("Imag", np.abs(d.imag))):
                if np.any(v > 1e-3):
                    raise RuntimeError("{:} cross-terms {:} do not agree within tolerance.".format(part, label))

        pairs = {"RS0": [("RS0xPS0", "PS0", "RS0_PS0")], "R": [("RxRS1", "RS1", "R_RS1"), ("RxP", "P", "R_P")],
                 "P": [("PxRS1", "RS1", "P_RS1")]}
        for s in self.strengths.values():
            if s.genopname not in pre:
                continue
            c = s.cstr_df
            b[s.bareop][s.k] = pre[s.genopname] ** 2 * c["Strength"].values
            for label, partner, key in pairs.get(s.bareop, []):
                pk = "%s_K%d" % (partner, s.k)
                check(s, self.strengths[pk], label)
                b[key][s.k] = pre[s.genopname] * pre[pk] * c[label].values
        return b

    # ---- shape factor by multipole (shape_factor.py:649-783; Behrens-Buehring combinations) ----------------------------
    def calcSfByJ(self, b):
        m, ps = self.sf_metadict, self.ps_df
        ec, L = m["ec"], m["LAM"]
        g1 = gam_ke(1, m["Z"])
        xi = 0.5 * ALPHA * m["Z"] / m["Radius"]
        xp, xm = m["W0_max"] / 3.0 + ec * xi, m["W0_max"] / 3.0 - ec * xi
        r2, r3, r6 = np.sqrt(2.0), np.sqrt(3.0), np.sqrt(6.0)
        dim = len(self.contour.ctr_z)
        C = {t: {k: np.zeros(dim) for k in range(3)} for t in self.al_ops + self.C_keys}
        for k in range(3):
            F, GT, RS0, PS0, R, P, RS1, RS2 = (b[o][k] for o in ("F", "GT", "RS0", "PS0", "R", "P", "RS1", "RS2"))
            RS0PS0, RRS1, PRS1, RP = (b[o][k] for o in ("RS0_PS0", "R_RS1", "P_RS1", "R_P"))
            C["F"][k] = F * ps["f2"]
            C["GT"][k] = L ** 2 * GT * ps["f2"]
            C["J0_R1"][k] = (-ec * 2.0 / 3.0 * L ** 2 * (xp * RS0 + RS0PS0)) * ps["f1"]
            C["J0_R2"][k] = (L ** 2 * ((xp ** 2 + 1.0 / 9.0) * RS0 + PS0 + 2.0 * xp * RS0PS0)) * ps["f2"]
            C["J1_R1"][k] = (-2.0 / 9.0 * (ec * xp * R - ec * 2.0 * L ** 2 * xm * RS1 + L * r2 * (xp - xm) * RRS1
                                           - ec * r3 * RP - L * r6 * PRS1)) * ps["f1"]
            C["J1_R2"][k] = (P + xp ** 2 / 3.0 * R + 2.0 / 3.0 * L ** 2 * xm ** 2 * RS1
                             + (R + 2.0 * L ** 2 * RS1 + ec * 2.0 * r2 * L * RRS1) / 27.0
                             + np.sqrt(2.0 / 3.0) * (ec * 2.0 * L * xm * PRS1 - r2 * xp * RP
                                                     - ec * 2.0 / r3 * L * xm * xp * RRS1)) * ps["f2"] \
                + -8.0 / 27.0 * (L ** 2 * RS1 + ec * L / r2 * RRS1) * mu_ke(1) * g1 * ps["f2"]
            C["J1_R3"][k] = (4.0 / 3.0 * (r2 / 3.0 * L * xp * RRS1 - ec * 2.0 / 3.0 * L ** 2 * xm * RS1
                                          - np.sqrt(2.0 / 3.0) * L * PRS1)) * ps["f3"]
            C["J1_R4"][k] = (8.0 / 27.0 * L ** 2 * RS1) * ps["f4"]
            C["J1_R5"][k] = ((2.0 * R + L ** 2 * RS1 + ec * 2.0 * r2 * L * RRS1) / 27.0) * ps["f5"]
            C["J1_R6"][k] = ((2.0 * R + L ** 2 * RS1 - ec * 2.0 * r2 * L * RRS1) / 27.0) * ps["f6"]
            C["J2_R5"][k] = (L ** 2 * RS2 / 9.0) * ps["f5"]
            C["J2_R6"][k] = (L ** 2 * RS2 / 9.0) * ps["f6"]
        return C

    # ---- the 19 contributions of beta.out (shape_factor.py:786-839) ----------------------------------------------------
    def calcSfContributions(self, C):
        allk = self.al_ops + self.C_keys
        J0, J1, J2, K = self.CJ0_keys, self.CJ1_keys, self.CJ2_keys, self.betaout_keys
        spec = [(allk, [0, 1, 2]), (self.al_ops, [0, 1]), (["GT"], [0, 1]), (self.C_keys, [0, 1, 2]),
                (["F"], [0]), (["GT"], [0]), (["GT"], [1]),
                (self.C_keys, [0]), (self.C_keys, [1]), (self.C_keys, [2]),
                (J0, [0, 1, 2]), (J1, [0, 1, 2]), (J2, [0, 1, 2]),
                (J0, [0]), (J1, [0]), (J1, [1]), (J2, [0]), (J2, [1]), (J2, [2])]
        dim = len(self.contour.ctr_z)
        self.sf_df = {}
        for name, (terms, ks) in zip(K, spec):
            tot = np.zeros(dim, dtype=complex)
            for t in terms:
                for k in ks:
                    tot = tot + C[t][k]
            self.sf_df[name] = tot

    def calcShapeFactor(self, hfb_gs, zero_neg=False):
        """Strengths x phase space -> self.sf_df {contribution: complex array on the contour} (shape_factor.py:293-357)."""
        if zero_neg:
            raise NotImplementedError("zeroed negative strength (open contours) is not built")
        self.prepConstants(hfb_gs)
        self.calcPhaseSpace()
        self.calcSfContributions(self.calcSfByJ(self.prepStrengths()))

    def calcBetaRates(self, emin=None, emax=None, quad=None):
        """Integrate every contribution along the contour: DataFrame [Rate(s^-1), Half-Life(s)] indexed by contribution,
        in the order of beta.out (shape_factor.py:956-1031)."""
        import pandas as pd
        if self.sf_df is None:
            raise RuntimeError("calcShapeFactor must be run before calcBetaRates.")
        quad = self.contour.quadrature if quad is None else quad
        rows = []
        for h in self.betaout_keys:
            cint = complex_quadrature(quad, self.contour, self.sf_df[h], emin, emax)
            rate = np.log(2) / KAPPA * (np.imag(-1j * np.pi * cint) if self.contour.closed else np.imag(cint))
            with np.errstate(divide="ignore"):
                rows.append((rate, np.log(2) / rate))
        self.sf_metadict.update({"quadratr": quad})
        return pd.DataFrame(rows, columns=["Rate(s^-1)", "Half-Life(s)"], index=self.betaout_keys)

    # ---- output files (shape_factor.py:1329-1410) -----------------------------------------------------------------------
    @property
    def sf_metastr(self):
        return self.sfSummaryString(self.sf_metadict)

    def sfSummaryString(self, sd_in):
        """The '# Summary Data' header of beta.out / shapefactor.out / phasespace.out."""
        import datetime
        sd = {}
        for k, v in sd_in.items():
            if v is None:
                sd[k] = ("{:<10}", "N/A")
            elif isinstance(v, (float, np.floating)):
                sd[k] = ("{:<10.5f}", v)
            elif isinstance(v, (int, np.integer)) and not isinstance(v, bool):
                sd[k] = ("{:<10d}", v)
            else:
                try:
                    slen = max(10, len(v))
                except TypeError:
                    slen = 10
                sd[k] = ("{:<" + str(36 if slen > 10 else slen) + "}", v)
        lines = [["FAM_ctr", "temper"], ["beta_type", "quadratr", "Half_Width", "screening"],
                 ["psi_approx", "psi_glpts", "ratint_pts"], ["Zi", "A", "Zf", "HFB_Qval"],
                 ["FAM_Qval", "EQRPAmax", "E_1stPeak", "|gA|/gV"], ["gA", "gV", "M_nucleon", "alpha*Z"],
                 ["Radius", "alpha*Z/2R", "W0_max", "W0*R"]]
        date = str(datetime.datetime.now().replace(microsecond=0))[:-3]
        out = "# PynFAM code version: 2.0.0-b200\n# Run Date: {:}\n# Summary Data:\n".format(date)
        for keys in lines:
            fmt = ", ".join("{:<11} = " + sd[k][0] for k in keys)
            out += ("#   " + fmt + "\n").format(*[x for k in keys for x in (k, sd[k][1])])
        return out + "#"

    def writeOutput(self, df, title, fname, dest="./"):
        import os
        text = df.to_string(header=True, index=True, col_space=3, float_format=lambda x: "{:25.16e}".format(x),
                            index_names=False)
        with open(os.path.join(dest, fname), "w") as f:
            f.write(title + "\n" + self.sf_metastr + "\n" + text + "\n")

    def writeBetaOut(self, dest="./", fname="beta.out"):
        """beta.out: the rates and half-lives of every contribution."""
        df = self.calcBetaRates()
        self.writeOutput(df, "# Nuclear Beta Decay Rates and Half-Lives", fname, dest)
        return df
