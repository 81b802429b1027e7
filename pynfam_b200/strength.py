"""In-process contour driver and strength assembly.

The reference computes the strength function of one (operator, K) along a complex-energy contour by farming one
`pnfam_main.x` process per contour point (pynfam/strength/fam_strength.py:204-249), parsing the `.dat` files back and
concatenating them into `OP.out` (text) and `OP.out.ctr` (Fortran-unformatted, the file betadecay / shapeFactor read).
Here the whole contour of an operator is ONE batched GPU solve (`gpu.Context.solve`), and this module mirrors the
reference-side interface around it -- same class / method names, argument meaning and file formats:

    famContour      pynfam/strength/contour.py:19-677      (all seven types: CIRCLE, CONSTL, CONSTR, FERMIS, FERMIA, EXP, MONOMIAL)
    famStrength     pynfam/strength/fam_strength.py:28-680 (concatFamStr, writeStrengthOut, writeCtrBinary, readCtrBinary)

so a maintainer replaces the `getFamList` + task-farm + `concatFamData` sequence by `famStrength.compute(...)`
(INTEGRATION.md).  There is no CPU fallback: `compute` needs the CUDA library and a device.
"""
import os
import re
import struct
import time

import numpy as np

KAPPA = 6147.0   # pynfam/config.py:225, decay constant [s]
NSTR_MAX = 8     # strength + cross-term columns of one operator (the reference has at most 1 + 5, pnfam_extfield.f90:908-947)
TMIN = 1e-3   # pynfam/config.py: finite temperature is not supported by this path (fails loudly in the host set-up)

_INTERVAL = dict(energy_min=0.0, energy_max=6.0, hfb_emin_buff=0.0)          # pynfam/config.py:141-146
_PROFILE = dict(hw_min=0.01, hw_max=0.4, de_hw_ratio=1.0, beta_quadrature="TRAP")
_CONTOUR_DEFAULTS = {
    # pynfam/config.py:147-196
    "CIRCLE": dict(_INTERVAL, nr_points=60, use_gauleg_ctr=True, beta_quadrature="GAUSS", shift_imag=0.0, theta_init=np.pi,
                   max_height=30),
    "CONSTL": dict(_INTERVAL, nr_points=60, half_width=0.1, beta_quadrature="TRAP"),
    "CONSTR": dict(_INTERVAL, half_width=0.1, de_hw_ratio=1.0, beta_quadrature="TRAP"),
    "EXP": dict(_INTERVAL, p_percent_interval=-0.25, **_PROFILE),
    "MONOMIAL": dict(_INTERVAL, power=1.0, **_PROFILE),
    "FERMIS": dict(_INTERVAL, u_percent_interval=0.50, t_percent_interval=0.10, **_PROFILE),
    "FERMIA": dict(_INTERVAL, nr_points_min=10, nr_points_max=70, u_percent_interval=0.50, t_percent_interval=0.10, **_PROFILE),
}


class famContour(object):
    """Complex-energy contour and its integration data (pynfam/strength/contour.py:19-131).

    Args:
        contour (str): 'CIRCLE', 'CONSTL', 'CONSTR', 'FERMIS', 'FERMIA', 'EXP' or 'MONOMIAL'.
        override (dict): settings overriding the defaults (unknown keys raise KeyError, as in the reference).
    """

    def __init__(self, contour, override=None):
        self.name = contour.upper()
        if self.name not in _CONTOUR_DEFAULTS:
            raise ValueError("Requested contour {:} not implemented.".format(self.name))
        self._settings = dict(_CONTOUR_DEFAULTS[self.name])
        self._generateCtrData()
        if override is not None:
            self.updateSettings(override)

    name_and_int = property(lambda s: "{:} on ({:.2e}, {:.2e})".format(s.name, s.energy_min, s.energy_max))
    closed = property(lambda s: s._ctr_data["closed"])
    nr_points = property(lambda s: s._ctr_data["nr_points"])
    nr_compute = property(lambda s: s._ctr_data["nr_compute"])
    use_gauleg = property(lambda s: s._ctr_data["use_gl_ctr"])
    ctr_z = property(lambda s: s._ctr_data["ctr_z"])
    ctr_dzdt = property(lambda s: s._ctr_data["ctr_dzdt"])
    theta = property(lambda s: s._ctr_data["theta"])
    glwts = property(lambda s: s._ctr_data["glwts"])
    half_width = property(lambda s: s._ctr_data["half_width"])
    quadrature = property(lambda s: s._ctr_data["quad"])
    energy_min = property(lambda s: s._settings["energy_min"])
    energy_max = property(lambda s: s._settings["energy_max"])

    def updateSettings(self, override):
        for h in override:
            if h not in self._settings:
                raise KeyError("Invalid override setting {:} for contour {:}.".format(h, self.name))
            self._settings[h] = override[h]
        self._generateCtrData()

    def setHfbInterval(self, hfb, beta, shift=0.0):
        """Interval [min(E_gs - buffer, 0), EQRPA_max] (+ shift) from a dict of HFB properties (contour.py:150-205).
        Finite temperature: the strength at negative energies matters, the interval starts at -30 MeV, and for electron
        capture (beta = 'c') it ends at +30 MeV instead of EQRPA_max (contour.py:184-198)."""
        Egs, eqrpamax = hfb["E_gs"], hfb["EQRPA_max"]
        emin = min(Egs - self._settings["hfb_emin_buff"], 0.0)
        emax = eqrpamax
        if hfb.get("ft_active", False):
            emin = -30.0
            if beta == "c":
                emax = 30.0
        self._settings["energy_min"] = emin + shift
        self._settings["energy_max"] = emax + shift
        if np.isnan(Egs) or np.isnan(eqrpamax):
            self._settings["energy_min"] = self._settings["energy_max"] = 0.0
        self._generateCtrData()

    def _generateCtrData(self):
        self._ctr_data = {"CIRCLE": self._contourCircle, "CONSTL": self._contourConstL, "CONSTR": self._contourConstR,
                          "FERMIS": self._contourFermiS, "FERMIA": self._contourFermiA, "EXP": self._contourExp,
                          "MONOMIAL": self._contourMonomial}[self.name]()

    def _contourCircle(self):
        """contour.py:212-283: circle through (energy_min, energy_max) centred on the real axis (an ellipse of height
        max_height when the radius exceeds it), Gauss-Legendre or equally spaced in theta; the lower half is computed
        and the upper half follows from S(w*) = S(w)* unless the contour is shifted or rotated."""
        s = self._settings
        npts, t0 = s["nr_points"], s["theta_init"]
        if s["use_gauleg_ctr"]:
            theta, glwts = np.polynomial.legendre.leggauss(npts)
            t1 = t0 + 2.0 * np.pi
            f1, f2 = 0.5 * (t0 + t1), 0.5 * (t1 - t0)
            glwts = glwts * f2
            theta = theta * f2 + f1
        else:
            theta = np.linspace(t0, t0 + 2.0 * np.pi, npts)
            glwts = np.zeros(npts)
        r0 = 0.5 * (s["energy_max"] + s["energy_min"])
        r = s["energy_max"] - r0
        if r > s["max_height"]:
            cos, sin = np.cos(theta), np.sin(theta)
            ctr_z = r0 + r * cos + s["max_height"] * sin * 1j
            ctr_dzdt = -r * sin + s["max_height"] * cos * 1j
        else:
            ctr_z = r0 + r * np.exp(1j * theta)
            ctr_dzdt = 1j * (ctr_z - r0)
        nr_compute = (npts + 1) // 2
        if s["shift_imag"] != 0:
            ctr_z = ctr_z + s["shift_imag"] * 1j
            nr_compute = npts
        elif s["use_gauleg_ctr"]:
            if abs(np.mod(t0, np.pi)) > 1e-10:
                nr_compute = npts
        else:
            if abs(np.mod(t0, np.pi / (npts - 1))) > 1e-10:
                nr_compute = npts
        return dict(nr_points=npts, nr_compute=nr_compute, use_gl_ctr=s["use_gauleg_ctr"], ctr_z=ctr_z,
                    ctr_dzdt=ctr_dzdt, theta=theta, glwts=glwts, half_width=None, quad=s["beta_quadrature"], closed=True)

    def _line(self, w):
        hw = self._settings["half_width"]
        ctr_z = (w * 1j + np.imag(hw)) if not np.isreal(hw) else (w + np.real(hw) * 1j)
        n = len(ctr_z)
        return dict(nr_points=n, nr_compute=n, use_gl_ctr=False, ctr_z=ctr_z, ctr_dzdt=np.ones(n), theta=np.zeros(n),
                    glwts=np.zeros(n), half_width=hw, quad=self._settings["beta_quadrature"], closed=False)

    def _contourConstR(self):
        """contour.py:286-318: line at constant half width, spacing = half_width * de_hw_ratio."""
        s = self._settings
        de = np.real(s["half_width"]) * s["de_hw_ratio"]
        return self._line(np.arange(s["energy_min"], s["energy_max"] + de, de))

    def _contourConstL(self):
        """contour.py:321-349: line at constant half width, nr_points equally spaced."""
        s = self._settings
        return self._line(np.linspace(s["energy_min"], s["energy_max"], s["nr_points"]))

    # ---- open contours whose half width follows a profile hw(w) = a f(w) + b and whose spacing follows the half width
    def _open(self, ctr_z):
        n = len(ctr_z)
        return dict(nr_points=n, nr_compute=n, use_gl_ctr=False, ctr_z=ctr_z, ctr_dzdt=np.ones(n), theta=np.zeros(n),
                    glwts=np.zeros(n), half_width=None, quad=self._settings["beta_quadrature"], closed=False)

    def _profile(self, func, anchor_at_max):
        """March from the interval start with steps de = hw(w) |de_hw_ratio| until the end is passed, then pin the last
        point to it (contour.py:474-520 FERMIS, 523-568 EXP, 571-617 MONOMIAL).  hw runs from hw_min to hw_max; the
        offset b is fixed at the upper (FERMIS) or lower (EXP, MONOMIAL) end."""
        s = self._settings
        e0, emax = s["energy_min"], float(s["energy_max"] - s["energy_min"])
        ratio = abs(s["de_hw_ratio"])
        a = (s["hw_max"] - s["hw_min"]) / (func(emax) - func(0.0))
        b = s["hw_max"] - a * func(emax) if anchor_at_max else s["hw_min"] - a * func(0.0)
        w = np.zeros(int(np.ceil(emax / float(s["hw_min"] * ratio))) + 1)
        i = 0
        while True:
            w[i + 1] = w[i] + (a * func(w[i]) + b) * ratio
            if w[i + 1] >= emax:
                break
            i += 1
        w = np.trim_zeros(w, "b")
        w[-1] = emax
        return self._open((w + e0) + (a * func(w) + b) * 1j)

    def _fermi_profile(self):
        s = self._settings
        emax = float(s["energy_max"] - s["energy_min"])
        u, t = s["u_percent_interval"] * emax, s["t_percent_interval"] * emax
        return lambda x: 1.0 - 1.0 / (np.exp((x - u) / t) + 1.0)

    def _contourFermiS(self):
        """Static Fermi-function profile."""
        return self._profile(self._fermi_profile(), True)

    def _contourExp(self):
        """Static exponential profile exp(w / p), p = p_percent_interval x interval."""
        s = self._settings
        p = s["p_percent_interval"] * float(s["energy_max"] - s["energy_min"])
        return self._profile(lambda x: np.exp(x / p), False)

    def _contourMonomial(self):
        """Static monomial profile w^power."""
        power = self._settings["power"]
        if power <= 0.0:
            raise ValueError("Monomial power must be >= 0")
        return self._profile(lambda x: x ** power, False)

    def _contourFermiA(self):
        """Adaptive Fermi profile (contour.py:352-470): fill the interval with at most nr_points_max points, small widths
        at small energies first -- a line at hw_min if that fits; else raise the right end of a Fermi profile, then the
        left end, in steps of 1e-4 MeV until the marched grid reaches the end of the interval; a line at hw_max (with
        as many points as that needs) when even that is too sparse."""
        s = self._settings
        e0, emax = s["energy_min"], float(s["energy_max"] - s["energy_min"])
        hwmin, hwmax, ncap = s["hw_min"], s["hw_max"], s["nr_points_max"]
        ratio = abs(s["de_hw_ratio"])
        de_min, de_max = float(hwmin * ratio), float(hwmax * ratio)
        if ncap <= int(np.ceil(emax / de_max)) + 1:
            w = np.arange(0.0, emax + de_max, de_max)
            return self._open((w + e0) + np.ones(len(w)) * hwmax * 1j)
        fermi = self._fermi_profile()
        step, shift_max = 0.0001, hwmax - hwmin
        shift_u = shift_l = 0.0
        a = b = 0.0
        while True:
            if shift_u == 0.0 and shift_l == 0.0:
                kind = 0
                w = np.arange(0.0, emax + de_min, de_min)
                if len(w) < s["nr_points_min"]:
                    w = np.linspace(0.0, emax, s["nr_points_min"])
            elif shift_u >= shift_max and shift_l >= shift_max:
                kind = 2
                w = np.arange(0.0, emax + de_max, de_max)
            else:
                kind = 1
                w = np.zeros(ncap)
                for i in range(ncap - 1):
                    hw = a * fermi(w[i]) + b
                    if s["de_hw_ratio"] < 0:      # spacing by arc length
                        de = np.sqrt((hw * ratio) ** 2 - (hw - (a * fermi(w[i - 1]) + b)) ** 2)
                    else:
                        de = hw * ratio
                    w[i + 1] = w[i] + de
            if w[-1] >= emax and len(w) <= ncap:
                break
            if shift_u < shift_max:
                shift_u += step
            else:
                shift_u = shift_max
                shift_l = min(shift_l + step, shift_max)
            ymax, ymin = hwmin + shift_u, hwmin + shift_l
            a = (ymax - ymin) / (fermi(emax) - fermi(0.0))
            b = ymin - a * fermi(emax)
        if w[-1] != emax:
            w[-1] = emax
        hw = np.ones(len(w)) * hwmin if kind == 0 else np.ones(len(w)) * hwmax if kind == 2 else a * fermi(w) + b
        return self._open((w + e0) + hw * 1j)


def patch_namelist(text, **values):
    """Return the pnfam namelist `text` with `key = value` replaced for every keyword (the keys pnfamRun overrides per
    task, pynfam/fortran/pnfam_run.py:208-270)."""
    for k, v in values.items():
        if isinstance(v, str):
            rep = "'%s'" % v
        elif isinstance(v, bool):
            rep = ".true." if v else ".false."
        elif isinstance(v, float):
            rep = repr(v)
        else:
            rep = str(v)
        text, n = re.subn(r"(?im)^(\s*%s\s*=\s*)[^\n,/]*" % re.escape(k), lambda m: m.group(1) + rep, text)
        if n == 0:
            raise KeyError("namelist has no key %r" % k)
    return text


class _FortranRecords(object):
    """Sequential unformatted records with 4-byte markers (what gfortran and scipy.io.FortranFile write)."""

    def __init__(self, path, mode):
        self.f = open(path, mode + "b")

    def write_record(self, *arrays):
        payload = b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)
        m = struct.pack("<i", len(payload))
        self.f.write(m + payload + m)

    def read_record(self, dtype):
        head = self.f.read(4)
        if len(head) != 4:
            raise IOError("end of file")
        (n,) = struct.unpack("<i", head)
        data = self.f.read(n)
        (n2,) = struct.unpack("<i", self.f.read(4))
        if n2 != n:
            raise IOError("record markers disagree")
        return np.frombuffer(data, dtype=dtype).copy()

    def close(self):
        self.f.close()


def convString(conv_list):
    """pynfam/utilities/hfb_utils.py convString: 'Yes' only if every point converged."""
    return "Yes" if all(c == "Yes" for c in conv_list) else "No"


class famStrength(object):
    """Strength function of one (operator, K) along a contour (pynfam/strength/fam_strength.py:28-92).

    Args:
        operator (str): operator name with the beta type, e.g. 'GT-' (pnfam namelist's operator_name).
        k (int): K projection.
        contour (famContour, str): contour, or a contour type name for the defaults.
        nucleus (tuple): (N, Z, A) of the parent (the reference takes it from the hfbthoRun object).
    """

    def __init__(self, operator, k, contour, nucleus=None):
        self.op, self.k = operator, k
        self.nucleus = nucleus
        self.temperature, self.ft_active = 0.0, False
        self.contour = famContour(contour) if isinstance(contour, str) else contour
        self.str_df = None      # DataFrame: Re(Strength), Im(Strength), Re(x1), Im(x1), ...
        self.meta_df = None     # DataFrame: Time (minutes), Conv ('Yes' / 'No')
        self._meta = {"Version": "Unknown", "Interaction": "Unknown", "Time": None, "Conv": None}
        self.stats = None

    file_txt = property(lambda s: "{:}.out".format(s.opname))
    file_bin = property(lambda s: "{:}.out.ctr".format(s.opname))
    opname = property(lambda s: "{:}K{:1d}".format(s.op, s.k))
    bareop = property(lambda s: s.op[:-1])
    beta = property(lambda s: s.op[-1])
    genopname = property(lambda s: "{:}_K{:1d}".format(s.bareop, s.k))
    use_FT_prefactor = property(lambda s: False)

    @property
    def xterms(self):
        return None if self.str_df is None else [str(c[3:-1]) for c in self.str_df.columns[2:] if "Re" in c]

    @property
    def nxterms(self):
        return None if self.str_df is None else len(self.xterms)

    @property
    def meta(self):
        if self.meta_df is not None:
            self._meta["Conv"] = convString(list(self.meta_df["Conv"].values))
            self._meta["Time"] = float(np.sum(self.meta_df["Time"].values))
        return self._meta

    @property
    def ctr_param_df(self):
        import pandas as pd
        df = pd.DataFrame()
        if self.contour.closed:
            df["Theta"] = self.contour.theta
        df["Re(EQRPA)"] = np.real(self.contour.ctr_z)
        df["Im(EQRPA)"] = np.imag(self.contour.ctr_z)
        return df

    @property
    def ctr_integ_df(self):
        import pandas as pd
        df = pd.DataFrame()
        if self.contour.closed:
            df["GL_Weights"] = self.contour.glwts
            df["Re(dzdt)"] = np.real(self.contour.ctr_dzdt)
            df["Im(dzdt)"] = np.imag(self.contour.ctr_dzdt)
        return df

    @property
    def cstr_df(self):
        """One complex column per strength / cross-term (fam_strength.py:166-178)."""
        if self.str_df is None:
            return None
        import pandas as pd
        out = pd.DataFrame()
        cols = list(self.str_df.columns)
        for i in range(0, len(cols), 2):
            out[cols[i][3:-1]] = self.str_df[cols[i]].values + self.str_df[cols[i + 1]].values * 1j
        return out

    # ---- the batched replacement of getFamList + task farm + concatFamData -------------------------------------
    def setup(self, rundir, namelist, share_nucleus_with=None):
        """Host set-up of this operator in rundir (what pnfamRun does per task: the namelist's operator_name /
        beta_type / operator_k are overridden by this object's).  Returns the host.Problem."""
        from . import host
        text = open(os.path.join(rundir, namelist)).read()
        text = patch_namelist(text, operator_name=self.bareop, beta_type=self.beta, operator_k=int(self.k))
        try:    # pynfam names every operator's files after it (OP.dat, OP.tbc for the two-body-current field)
            text = patch_namelist(text, fam_output_filename=self.opname)
        except KeyError:
            pass
        name = "%s.b200.in" % self.opname
        tmp = os.path.join(rundir, ".%s.%d.tmp" % (name, os.getpid()))      # several ranks may share the rundir
        with open(tmp, "w") as f:
            f.write(text)
        os.replace(tmp, os.path.join(rundir, name))
        prob = host.Problem(rundir, name, share_nucleus_with=share_nucleus_with)
        if self.nucleus is None:
            n, z = prob.iscalar("npr_n"), prob.iscalar("npr_p")
            self.nucleus = (n, z, n + z)
        m = re.search(r"(?im)^\s*interaction_name\s*=\s*['\"]([^'\"]*)['\"]", text)
        self._meta["Interaction"] = m.group(1).strip() if m else "Unknown"
        self._meta["Version"] = "b200"
        return prob

    def solve_points(self, prob, ctx, points=None, **solve_kw):
        """One batched GPU solve of contour.ctr_z[points] (default: the nr_compute points that are computed).
        Returns the dict of gpu.Context.solve plus 'minutes' (per-point share of the wall time, by iterations)."""
        c = self.contour
        idx = np.arange(c.nr_compute) if points is None else np.asarray(points, dtype=int)
        t0 = time.time()
        res = ctx.solve(prob, omegas=np.asarray(c.ctr_z)[idx], **solve_kw)
        wall_min = (time.time() - t0) / 60.0
        it = np.maximum(res["iters"].astype(float), 1.0)
        res["minutes"] = wall_min * it / it.sum()
        res["points"] = idx
        return res

    def compute(self, rundir, namelist, ctx=None, device=0, share_nucleus_with=None, **solve_kw):
        """Solve the operator on contour.ctr_z[:nr_compute] in one batched GPU call and fill str_df / meta_df.

        rundir holds hfbtho_NAMELIST.dat + hfbtho_output.hel (+ .tbc); `namelist` is a pnfam namelist in rundir whose
        operator_name / beta_type / operator_k are overridden by this object's (what pnfamRun does per task).  Returns
        (problem, ctx) so the caller can reuse the device-resident nucleus for the next operator."""
        from . import gpu
        prob = self.setup(rundir, namelist, share_nucleus_with=share_nucleus_with)
        if ctx is None:
            ctx = gpu.Context(prob, device=device)
        res = self.solve_points(prob, ctx, **solve_kw)
        self.concatFamData(res["strength"], res["labels"], ["Yes" if c else "No" for c in res["conv"]], list(res["minutes"]))
        self.stats = res["stats"]
        self.iters = res["iters"]
        return prob, ctx

    def concatFamStr(self, strength, labels):
        """Computed points -> full-contour DataFrame; a closed contour with nr_compute = (nr_points+1)//2 is completed
        with S(w*) = S(w)* (fam_strength.py:293-337).  strength: complex [nr_compute, 1 + nxterms]; labels: the row
        labels of the .dat result table ('Strength', cross-term labels)."""
        import pandas as pd
        strength = np.asarray(strength)
        c = self.contour
        if strength.shape[0] != c.nr_compute:
            raise RuntimeError("fam data does not match contour data.")
        data = {}
        header_order = []
        for j, lab in enumerate(labels):
            lab = "Strength" if j == 0 else lab
            data["Re(%s)" % lab] = strength[:, j].real.copy()
            data["Im(%s)" % lab] = strength[:, j].imag.copy()
            header_order += ["Re(%s)" % lab, "Im(%s)" % lab]
        op_str_df = pd.DataFrame(data)
        if strength.shape[0] != c.nr_points:
            endpoints = c.ctr_z[0] == c.ctr_z[-1]
            sym = op_str_df.iloc[::-1].copy()
            if c.nr_compute * 2 != c.nr_points:
                sym.drop(sym.head(1).index, inplace=True)
            for h in header_order:
                if "Im" in h:
                    sym[h] = -sym[h].values
            if endpoints and c.nr_points != 1:
                sym.iloc[-1] = op_str_df.iloc[0].values
            op_str_df = pd.concat([op_str_df, sym], axis=0, ignore_index=True)
        return op_str_df[header_order]

    def concatFamData(self, strength, labels, conv_list, time_list):
        import pandas as pd
        self.str_df = self.concatFamStr(strength, labels)
        self.meta_df = pd.DataFrame({"Time": time_list, "Conv": conv_list})

    # ---- outputs ------------------------------------------------------------------------------------------------
    def writeStrengthOut(self, dest="./", fname=None):
        """OP.out, the text summary pynfam writes and re-reads (fam_strength.py:456-488)."""
        import pandas as pd
        if self.str_df is None:
            raise RuntimeError("Strength dataframe has not been populated. Cannot write to .out")
        out_df = pd.concat([self.ctr_param_df, self.str_df, self.ctr_integ_df, self.meta_df], axis=1)
        if fname is None:
            fname = self.file_txt
        pd_string = out_df.to_string(header=True, index=True, col_space=3,
                                     float_format=lambda x: "{:25.16e}".format(x))
        header = ["# pnFAM code version:   {:}\n".format(self.meta["Version"]),
                  "# Total run time:       {:<.6f} mins\n".format(self.meta["Time"]),
                  "# All points converged: {:}\n".format(self.meta["Conv"]),
                  "# Residual interaction: {:}\n".format(self.meta["Interaction"]),
                  "# Operator:             {:} with K={:d}\n".format(self.op, self.k),
                  "# Contour:              {:}\n".format(self.contour.name_and_int),
                  "# Temperature:          {:<.6f} (Prefactor={:})\n".format(self.temperature, self.use_FT_prefactor),
                  "#\n"]
        with open(os.path.join(dest, fname), "w") as f:
            f.writelines(header)
            f.write(pd_string + "\n")

    def writeCtrBinary(self, dest="./", fname=None):
        """OP.out.ctr, version 3 of the Fortran-unformatted contour file read by betadecay.x and shapeFactor
        (fam_strength.py:491-568): version | Z, A | operator(80) | K, nr_points, nxterms | labels(80 each) | theta |
        dz/dt | z | strength(nr_points, 1+nxterms) column-major | use_gauleg | GL weights."""
        if self.str_df is None:
            raise RuntimeError("Strength dataframe has not been populated. Cannot write to .ctr")
        if fname is None:
            fname = self.file_bin
        i4, f8, c16 = np.int32, np.float64, np.complex128
        c = self.contour
        w = _FortranRecords(os.path.join(dest, fname), "w")
        w.write_record(np.array([3], dtype=i4))
        w.write_record(np.array([self.nucleus[1], self.nucleus[2]], dtype=i4))
        pad = lambda s: np.frombuffer(s.encode() + (80 - len(s)) * b" ", dtype="S80")
        w.write_record(pad(self.bareop))
        w.write_record(np.array([self.k, c.nr_points, self.nxterms], dtype=i4))
        if self.nxterms > 0:
            w.write_record(*[pad(xt) for xt in self.xterms])
        else:
            w.write_record(pad(""))
        w.write_record(np.asarray(c.theta).astype(f8))
        w.write_record(np.asarray(c.ctr_dzdt).astype(c16))
        w.write_record(np.asarray(c.ctr_z).astype(c16))
        w.write_record(np.ascontiguousarray(self.cstr_df.values.T).astype(c16))
        w.write_record(np.array([int(c.use_gauleg)], dtype=i4))
        w.write_record(np.asarray(c.glwts).astype(f8))
        w.close()

    def readCtrBinary(self, src="./", fname=None):
        """Populate this object from an OP.out.ctr file (fam_strength.py:571-668)."""
        import pandas as pd
        if fname is None:
            fname = self.file_bin
        path = os.path.join(src, fname)
        if not os.path.exists(path):
            raise IOError("Binary file not found.")
        r = _FortranRecords(path, "r")
        i4, f8, c16 = np.int32, np.float64, np.complex128
        version = r.read_record(i4)[0]
        nucleus = r.read_record(i4)
        op = r.read_record("S80")[0]
        k, nr_points, nxterms = r.read_record(i4)
        raw = r.read_record("S1").tobytes()
        xterms = [raw[i:i + 80].decode().strip() for i in range(0, len(raw), 80)] if len(raw) >= 80 else []
        xterms = [x for x in xterms if x]
        theta = r.read_record(f8)
        ctr_dzdt = r.read_record(c16)
        ctr_z = r.read_record(c16)
        strength = r.read_record(c16).reshape(len(xterms) + 1, len(theta)).T
        use_gauleg = r.read_record(i4)[0]
        glwts = r.read_record(f8)
        r.close()
        bareop = op.decode().strip()
        if bareop != self.bareop or k != self.k:
            raise IOError("File contents do not match requested operator.\nExpected: {:}, {:}, Read: {:}, {:}."
                          .format(self.bareop, self.k, bareop, k))
        str_df = pd.DataFrame()
        for h, col in zip(["Strength"] + xterms, strength.T):
            str_df["Re({:})".format(h)] = np.real(col)
            str_df["Im({:})".format(h)] = np.imag(col)
        closed, quad = self.contour.closed, self.contour.quadrature
        if np.any(theta != 0.0) and not closed:
            closed, quad = True, "GAUSS"
        half_width = None if np.any(np.imag(ctr_z) != np.imag(ctr_z[0])) else np.imag(ctr_z[0])
        self.contour._ctr_data = dict(nr_points=int(nr_points), nr_compute=(int(nr_points) + 1) // 2 if closed else int(nr_points),
                                      use_gl_ctr=int(use_gauleg), ctr_z=ctr_z, ctr_dzdt=ctr_dzdt, theta=theta, glwts=glwts,
                                      half_width=half_width, quad=quad, closed=closed)
        self.str_df = str_df
        self.nucleus = (int(nucleus[1] - nucleus[0]), int(nucleus[0]), int(nucleus[1]))
        self.version = int(version)


def complex_quadrature(quad, contour, y, xmin=None, xmax=None):
    """Quadrature of y along the contour: complex contour integral / (2 pi i) on a closed contour, plain integral over
    Re(z) on an open one (shapeFactor.complex_quadrature, pynfam/strength/shape_factor.py:1161-1197)."""
    if quad not in ("TRAP", "GAUSS", "SIMPSON"):
        raise ValueError("Invalid quadrature requested.")
    y = np.asarray(y)
    if contour.closed:
        fac, dzdt, dt = 1.0 / (2.0 * np.pi * 1j), contour.ctr_dzdt, contour.theta
    else:
        fac, dzdt, dt = 1.0, 1.0, np.real(contour.ctr_z)
        xmin = min(dt) if xmin is None else xmin
        xmax = max(dt) if xmax is None else xmax
        mask = (dt >= xmin) & (dt <= xmax)
        dt, y = dt[mask], y[mask]
    if quad == "GAUSS":
        if not contour.use_gauleg:
            raise RuntimeError("Integration points do not lie on a gauss-legendre grid.")
        return fac * sum((y * dzdt) * contour.glwts)
    if quad == "TRAP":
        return fac * np.trapezoid(y * dzdt, dt)
    from scipy.integrate import simpson
    return fac * simpson(y * dzdt, x=dt)


def beta_rate(contour, weighted_shape_factor, quad=None, emin=None, emax=None):
    """Rate [1/s] and half-life [s] of one channel from its phase-space weighted shape factor on the contour
    (shapeFactor.calcBetaRates, shape_factor.py:956-1031): ln2/kappa * Re(contour integral) on a closed contour
    (the strengths carry -1/pi), ln2/kappa * Im(integral) on an open one."""
    quad = contour.quadrature if quad is None else quad
    cint = complex_quadrature(quad, contour, weighted_shape_factor, emin, emax)
    if contour.closed:
        rate = np.log(2) / KAPPA * np.imag(-1j * np.pi * cint)
    else:
        rate = np.log(2) / KAPPA * np.imag(cint)
    with np.errstate(divide="ignore"):
        return rate, np.log(2) / rate


def run_contours(rundir, namelist, operators, contour, dest=None, device=0, **solve_kw):
    """All (operator, K) of a nucleus on one contour: the nucleus (HFB reconstruction, basis tables) is set up and
    uploaded once and shared, each operator is one batched solve, and OP.out / OP.out.ctr are written to `dest`
    (default rundir) -- what mpi_workflow_fam + famStrength do with nr_compute x len(operators) process launches.
    operators: iterable of (operator_name, K).  Returns the list of famStrength objects."""
    dest = rundir if dest is None else dest
    out, first, ctx = [], None, None
    for op, k in operators:
        fs = famStrength(op, k, contour)
        prob, ctx = fs.compute(rundir, namelist, ctx=ctx, device=device, share_nucleus_with=first, **solve_kw)
        first = first or prob
        fs._keep = prob
        fs.writeStrengthOut(dest)
        fs.writeCtrBinary(dest)
        out.append(fs)
    return out


def _missing_tbc_files(rundir, namelist, fss):
    """Indices of the operators whose full-FAM two-body-current field has to be computed: GT with
    two_body_current_mode = x1x1xx or x5x1xx (pnfam_solver.f90:596-640) and no <OP>.tbc in rundir."""
    text = open(os.path.join(rundir, namelist)).read()
    m = re.search(r"(?im)^\s*two_body_current_mode\s*=\s*(\d+)", text)
    mode = int(m.group(1)) if m else 0
    if mode == 0 or (mode // 10000) % 10 not in (1, 5) or (mode // 100) % 10 == 0:
        return []
    return [o for o, fs in enumerate(fss) if fs.bareop.upper() == "GT" and not os.path.isfile(os.path.join(rundir, fs.opname + ".tbc"))]


def run_contours_sharded(rundir, namelist, operators, contour, dest=None, dist=None, device=0, solve_points=None,
                         concurrent_solves=None, **solve_kw):
    """run_contours over the ranks of a torch.distributed job (one process per GPU): the (operator, contour point) tasks
    are dealt to the ranks by shard.partition_tasks, every rank sets the nucleus and the operators it owns up once and solves its points of each
    operator as one batch, and the ONLY exchange is one all_reduce of the strengths (disjoint ownership, so a sum is a
    gather; the row labels travel beside it) -- NCCL on GPUs, gloo in the CPU tests.  Rank 0 assembles and writes OP.out / OP.out.ctr for every operator;
    every rank returns the list of famStrength objects.  `solve_points(fs, prob, ctx, points)` may replace the GPU
    solve (tests).  concurrent_solves (default 2, PNFAM_B200_CONCURRENT_SOLVES): operators of a rank solved side by side,
    each by its own host thread on its own context and stream, so that the tail of one batch (the few near-axis points
    that need twice the iterations of the others) overlaps with the bulk of the next operator's."""
    import torch
    from . import shard
    dest = rundir if dest is None else dest
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    world = dist.get_world_size() if multi else 1
    operators = list(operators)
    nc = contour.nr_compute
    mine = shard.partition_tasks(len(operators), contour.ctr_z[:nc], world)[rank]
    # a rank sets up the nucleus once and only the operators it owns points of; labels travel with the results
    t_start = time.perf_counter()
    fss = [famStrength(op, k, contour) for op, k in operators]
    probs, first = {}, None
    if multi:
        # Full-FAM two-body currents: a missing <OP>.tbc is computed by the set-up (csrc/host/tbc_generator.cpp, seconds to
        # minutes per operator).  Every missing file is generated by ONE rank, the files dealt round robin, before the
        # ranks that own points of those operators set them up and read the cached file.  The list is taken before any
        # rank starts writing.
        todo = _missing_tbc_files(rundir, namelist, fss)
        dist.barrier()
        # The HFB reconstruction is the same on every rank: rank 0 does it once with all host threads and leaves the
        # solution in the set-up cache of the run directory (csrc/host/hfb_front.cpp), the other ranks load it.
        if rank == 0:
            from . import host
            prev = host.set_threads(os.cpu_count() or 1)
            owned = [o for o, idx in mine if len(idx)]
            o0 = next((o for o in owned if o not in todo), todo[0] if todo else (owned[0] if owned else 0))
            probs[o0] = fss[o0].setup(rundir, namelist)
            first = probs[o0]
            host.set_threads(prev)
        dist.barrier()
        for j, o in enumerate(todo):
            if j % world == rank and o not in probs:
                probs[o] = fss[o].setup(rundir, namelist, share_nucleus_with=first)
                first = first or probs[o]
        if todo:
            dist.barrier()
    for o, idx in mine:
        if len(idx) and o not in probs:
            probs[o] = fss[o].setup(rundir, namelist, share_nucleus_with=first)
            first = first or probs[o]
    t_setup = time.perf_counter()
    nstr = NSTR_MAX
    buf = np.zeros((len(operators), nc, 2 * nstr + 3))          # re | im | conv, iterations, minutes
    labels = {}
    work = [(o, idx) for o, idx in mine if len(idx)]
    if concurrent_solves is None:
        concurrent_solves = int(os.environ.get("PNFAM_B200_CONCURRENT_SOLVES", "2"))
    nworkers = max(1, min(int(concurrent_solves), len(work)))
    import threading
    tls = threading.local()

    def job(item):
        o, idx = item
        if solve_points is not None:
            res = solve_points(fss[o], probs[o], None, idx)
        else:
            if getattr(tls, "ctx", None) is None:
                from . import gpu
                tls.ctx = gpu.Context(probs[o], device=device)       # one context (stream, side streams) per host thread
            res = fss[o].solve_points(probs[o], tls.ctx, points=idx, **solve_kw)
        n1 = res["strength"].shape[1]
        if n1 > nstr:
            raise RuntimeError("operator with more than %d strength columns" % nstr)
        labels[o] = list(res["labels"])
        buf[o, idx, :n1] = res["strength"].real
        buf[o, idx, nstr:nstr + n1] = res["strength"].imag
        buf[o, idx, 2 * nstr] = res["conv"]
        buf[o, idx, 2 * nstr + 1] = res["iters"]
        buf[o, idx, 2 * nstr + 2] = res["minutes"]

    if nworkers <= 1:
        for item in work:
            job(item)
    else:
        from concurrent.futures import ThreadPoolExecutor
        # the largest batches first: the pool drains with the small ones
        with ThreadPoolExecutor(nworkers) as ex:
            list(ex.map(job, sorted(work, key=lambda w: -len(w[1]))))
    t_solve = time.perf_counter()
    meta = {"nucleus": fss[mine[0][0]].nucleus, "meta": dict(fss[mine[0][0]]._meta)} if mine and len(mine[0][1]) else None
    if multi:
        on_gpu = dist.get_backend() == "nccl"
        t = torch.as_tensor(buf, device=("cuda:%d" % device) if on_gpu else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        buf = t.cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, (labels, meta))
        for lab, m in parts:
            labels.update(lab)
            meta = meta or m
    for o, fs in enumerate(fss):
        n1 = len(labels[o])
        if fs.nucleus is None:
            fs.nucleus = tuple(meta["nucleus"])
            fs._meta.update({k: meta["meta"][k] for k in ("Version", "Interaction")})
        fs.concatFamData(buf[o, :, :n1] + 1j * buf[o, :, nstr:nstr + n1], labels[o],
                         ["Yes" if c > 0.5 else "No" for c in buf[o, :, 2 * nstr]], list(buf[o, :, 2 * nstr + 2]))
        fs.iters = buf[o, :, 2 * nstr + 1].astype(int)
        fs._keep = probs.get(o)
        if rank == 0:
            fs.writeStrengthOut(dest)
            fs.writeCtrBinary(dest)
    if multi:
        dist.barrier()
    t_end = time.perf_counter()
    run_contours_sharded.last_timing = {"host_setup_s": t_setup - t_start, "solve_s": t_solve - t_setup,
                                        "gather_and_write_s": t_end - t_solve}
    return fss
