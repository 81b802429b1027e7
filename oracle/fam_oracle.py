"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/, smoke() and the
cpu_baseline / reference legs of bench.py may import this module.

numpy restatement of the reference's FAM iteration for one (operator, K, omega):

  ifam                         exes/pnfam/pnfam_solver.f90:93-221
  init_pnfam_solver            exes/pnfam/pnfam_solver.f90:226-460   (F -> qp basis, Greens, T)
  matrix_2qp                   exes/pnfam/pnfam_solver.f90:510-544
  triprod                      exes/pnfam/pnfam_type_blockmatrix.f90:140-212
  bigblockmatrix algebra       exes/pnfam/pnfam_type_bbm.f90:145-156, 217-576
  calc_hamiltonian             exes/pnfam/pnfam_hamiltonian_blas.f90:51-74
     density                   :124-711
     meanfield                 :717-1169
     pairingfield              :1175-1262
  qrpa_broyden/broyden_method  exes/pnfam/pnfam_broyden.f90:26-216

It follows the reference's own data model literally (block matrices with per-quadrant sign and
transpose flags; "represented quadrant = s * op_t(stored)"), which is deliberately different from
the product's CUDA formulation, so that the two can check each other.

Pinned against: the reference's golden per-point outputs tests/**/fam_meta/*.tar (committed as
tests/golden/*.json by tests/golden/make_golden.py) and against the live reference binary
(oracle/_ref/pnfam_main.x) -- see tests/test_oracle_golden.py.

Inputs come as a plain dict of numpy arrays (the `model`) that mirrors the Fortran module variables
(hfb_solution, pnfam_interaction, type_blockmatrix); pynfam_b200.host.Problem provides them.
"""
import numpy as np

# single-precision literal stored in a double: `real(pr) :: qrpa_alphamix = 0.7`
# (pnfam_broyden.f90:20)
ALPHAMIX = float(np.float32(0.7))


# ------------------------------------------------------------------------------------------------
# blockmatrix
# ------------------------------------------------------------------------------------------------
class BlockMatrix:
    """pnfam_type_blockmatrix.f90:15-26.  Indices kept 1-based (0 = no block) like the reference."""

    def __init__(self, nb, n=0):
        self.ir2c = np.zeros(nb, np.int64)
        self.ic2r = np.zeros(nb, np.int64)
        self.ir2m = np.zeros(nb, np.int64)
        self.ic2m = np.zeros(nb, np.int64)
        self.elem = np.zeros(n)

    def copy(self):
        b = BlockMatrix(len(self.ir2c))
        b.ir2c, b.ic2r, b.ir2m, b.ic2m = self.ir2c.copy(), self.ic2r.copy(), self.ir2m.copy(), self.ic2m.copy()
        b.elem = self.elem.copy()
        return b

    def copy_structure_from(self, a):
        self.ir2c, self.ic2r, self.ir2m, self.ic2m = a.ir2c.copy(), a.ic2r.copy(), a.ir2m.copy(), a.ic2m.copy()

    @staticmethod
    def from_rows(db, ir2c, elem):
        nb = len(db)
        b = BlockMatrix(nb, len(elem))
        b.elem = np.array(elem, float)
        ipt = 1
        for i in range(nb):
            j = int(ir2c[i])
            if j > 0:
                b.ir2c[i] = j
                b.ic2r[j - 1] = i + 1
                b.ir2m[i] = ipt
                b.ic2m[j - 1] = ipt
                ipt += db[i] * db[j - 1]
        return b


def block_diag_structure(db):
    nb = len(db)
    b = BlockMatrix(nb)
    b.ir2c = np.arange(1, nb + 1)
    b.ic2r = b.ir2c.copy()
    off = 1
    for i in range(nb):
        b.ir2m[i] = off
        off += db[i] ** 2
    b.ic2m = b.ir2m.copy()
    return b


def triprod(db, ta, a, tb, b, tc, c, alpha, beta, abc):
    """ABC = alpha * op(A) op(B) op(C) + beta*ABC  (pnfam_type_blockmatrix.f90:140-212).
    Recomputes abc's block maps; blocks of A and C are square."""
    nb = len(db)
    abc.ir2c[:] = 0
    abc.ic2r[:] = 0
    abc.ir2m[:] = 0
    abc.ic2m[:] = 0
    ipt = 1
    for i in range(1, nb + 1):
        if ta == 'n':
            k, ipa = a.ir2c[i - 1], a.ir2m[i - 1]
        else:
            k, ipa = a.ic2r[i - 1], a.ic2m[i - 1]
        if k == 0:
            continue
        if tb == 'n':
            l, ipb = b.ir2c[k - 1], b.ir2m[k - 1]
        else:
            l, ipb = b.ic2r[k - 1], b.ic2m[k - 1]
        if l == 0:
            continue
        if tc == 'n':
            j, ipc = c.ir2c[l - 1], c.ir2m[l - 1]
        else:
            j, ipc = c.ic2r[l - 1], c.ic2m[l - 1]
        if j == 0:
            continue
        abc.ir2c[i - 1] = j
        abc.ic2r[j - 1] = i
        abc.ir2m[i - 1] = ipt
        abc.ic2m[j - 1] = ipt
        nda, ndc = db[i - 1], db[j - 1]
        A = a.elem[ipa - 1:ipa - 1 + nda * nda].reshape(nda, nda, order='F')
        if ta != 'n':
            A = A.T
        if tb == 'n':
            B = b.elem[ipb - 1:ipb - 1 + nda * ndc].reshape(nda, ndc, order='F')
        else:
            B = b.elem[ipb - 1:ipb - 1 + nda * ndc].reshape(ndc, nda, order='F').T
        C = c.elem[ipc - 1:ipc - 1 + ndc * ndc].reshape(ndc, ndc, order='F')
        if tc != 'n':
            C = C.T
        res = alpha * (A @ B) @ C
        sl = slice(ipt - 1, ipt - 1 + nda * ndc)
        abc.elem[sl] = res.reshape(-1, order='F') + beta * abc.elem[sl]
        ipt += nda * ndc


# ------------------------------------------------------------------------------------------------
# bigblockmatrix
# ------------------------------------------------------------------------------------------------
class BBM:
    """pnfam_type_bbm.f90:30-35."""

    def __init__(self):
        self.m = {11: None, 12: None, 21: None, 22: None}
        self.t = {11: 'n', 12: 'n', 21: 'n', 22: 'n'}
        self.s = {11: 1.0, 12: 1.0, 21: 1.0, 22: 1.0}
        self.imag = False

    def shallow(self):
        b = BBM()
        b.m, b.t, b.s, b.imag = dict(self.m), dict(self.t), dict(self.s), self.imag
        return b

    def allocate(self, nb, a11, a12, a21, a22, n):
        for q, on in zip((11, 12, 21, 22), (a11, a12, a21, a22)):
            if on:
                self.m[q] = BlockMatrix(nb, n)

    def set_sign(self, s11, s12, s21, s22):
        self.s = {11: float(s11), 12: float(s12), 21: float(s21), 22: float(s22)}

    def set_trans(self, t11, t12, t21, t22):
        self.t = {11: t11, 12: t12, 21: t21, 22: t22}

    def transpose(self):
        """transpose_bbm :145-156"""
        self.m[12], self.m[21] = self.m[21], self.m[12]
        self.s[12], self.s[21] = self.s[21], self.s[12]
        for q in self.t:
            self.t[q] = 't' if self.t[q] == 'n' else 'n'

    def set_val(self, v):
        for q in self.m:
            if self.m[q] is not None:
                self.m[q].elem[:] = v


_TERMS = {  # (a quadrant, b quadrant, c quadrant) per output quadrant, in the reference's order
    11: ((11, 11, 11), (12, 21, 11), (11, 12, 21), (12, 22, 21)),
    12: ((11, 11, 12), (12, 21, 12), (11, 12, 22), (12, 22, 22)),
    21: ((21, 11, 11), (22, 21, 11), (21, 12, 21), (22, 22, 21)),
    22: ((21, 11, 12), (22, 21, 12), (21, 12, 22), (22, 22, 22)),
}


def triprod_bbm_quad(db, ta, ain, tb, bin_, tc, cin, qabc, sabc, tabc, out):
    """:428-552"""
    a, b, c = ain.shallow(), bin_.shallow(), cin.shallow()
    quadrant, s = qabc, sabc
    if ta == 't':
        a.transpose()
    if tb == 't':
        b.transpose()
    if tc == 't':
        c.transpose()
    if tabc == 't':
        a, c = c, a
        a.transpose()
        b.transpose()
        c.transpose()
        quadrant = {12: 21, 21: 12}.get(quadrant, quadrant)
    im = int(a.imag) + int(b.imag) + int(c.imag)
    if im >= 2:
        s = -s
    out.elem[:] = 0
    for qa, qb, qc in _TERMS[quadrant]:
        if a.m[qa] is not None and b.m[qb] is not None and c.m[qc] is not None:
            triprod(db, a.t[qa], a.m[qa], b.t[qb], b.m[qb], c.t[qc], c.m[qc],
                    s * a.s[qa] * b.s[qb] * c.s[qc], 1.0, out)


def triprod_bbm(db, ta, a, tb, b, tc, c, abc):
    """:557-576"""
    for q in (11, 12, 21, 22):
        if abc.m[q] is not None:
            triprod_bbm_quad(db, ta, a, tb, b, tc, c, q, abc.s[q], abc.t[q], abc.m[q])


def contract_bbm(a, b):
    """:393-421"""
    s = 0.0
    for q in (11, 12, 21, 22):
        if a.m[q] is not None and b.m[q] is not None:
            s += float(np.dot(a.m[q].elem, b.m[q].elem))
    if int(a.imag) + int(b.imag) == 2:
        s = -s
    return s


# ------------------------------------------------------------------------------------------------
# calc_hamiltonian
# ------------------------------------------------------------------------------------------------
def _isstart(db):
    s = np.zeros(len(db), np.int64)
    s[0] = 1
    s[1:] = np.cumsum(db)[:-1] + 1
    return s


def density(model, rho, kap):
    """density (:124-711).  rho, kap: complex BlockMatrix-like (re, im pairs).  Returns dict of the
    28 complex local densities (each nghl)."""
    db, nb, N, Ng = model['db'], model['nb'], model['dqp'], model['nghl']
    ns, nsu, isstart = model['ns'], model['num_spin_up'], _isstart(model['db'])
    wf, wfdr, wfdp, wfdz = model['wf'], model['wfdr'], model['wfdp'], model['wfdz']
    (rerho, imrho), (rek, imk) = rho, kap

    def wfa(tables, re, im):
        out = {s: [np.zeros((Ng, N), complex) for _ in tables] for s in (1, -1)}
        for ix in range(nb):
            iy = int(re.ir2c[ix])
            if iy == 0:
                continue
            ia, ib = isstart[ix] - 1, isstart[iy - 1] - 1
            d1, d2 = db[ix], db[iy - 1]
            o = re.ir2m[ix] - 1
            blk = (re.elem[o:o + d1 * d2] + 1j * im.elem[o:o + d1 * d2]).reshape(d1, d2, order='F')
            nu = nsu[ix]
            for t, tab in enumerate(tables):
                if nu > 0:
                    out[1][t][:, ib:ib + d2] = tab[:, ia:ia + nu] @ blk[:nu, :]
                if d1 - nu > 0:
                    out[-1][t][:, ib:ib + d2] = tab[:, ia + nu:ia + d1] @ blk[nu:, :]
        return out

    A = wfa((wf, wfdr, wfdp, wfdz), rerho, imrho)
    names = ('rho tau tjrr tjpr tjzr tjrp tjpp tjzp tjrz tjpz tjzz sr sp sz tr tp tz jr jp jz fr fp fz gs '
             'rb sbr sbp sbz').split()
    d = {k: np.zeros(Ng, complex) for k in names}
    I = 1j
    for b in range(N):
        s = int(ns[b])
        wb, drb, dpb, dzb = wf[:, b], wfdr[:, b], wfdp[:, b], wfdz[:, b]
        a0, ar, ap, az = (A[s][t][:, b] for t in range(4))
        # --- diagonal in spin ---
        prod = a0 * wb
        d['rho'] += prod
        d['sz'] += s * prod
        aux = (a0 * dpb + ap * wb) / 2
        d['jp'] += aux
        d['tjpz'] += s * aux
        aux = ar * drb + ap * dpb + az * dzb
        d['tau'] += aux
        d['tz'] += s * aux
        aux = (a0 * drb - ar * wb) / 2
        d['jr'] += I * aux          # dsrejr -= aux.im ; dsimjr += aux.re
        d['tjrz'] += s * I * aux
        aux = (a0 * dzb - az * wb) / 2
        d['jz'] += I * aux
        d['tjzz'] += s * I * aux
        aux = a0 * dzb + az * wb
        d['gs'] += s * aux
        aux = ar * dzb + az * drb
        d['fr'] += s * aux / 2
        aux = ap * dzb - az * dpb
        d['fp'] += s * I * aux / 2
        aux = az * dzb
        d['fz'] += s * aux
        # --- off-diagonal in spin ---
        a0, ar, ap, az = (A[-s][t][:, b] for t in range(4))
        prod = a0 * wb
        if s == -1:       # |a>=|+>, |b>=|->
            d['sr'] += prod
            d['sp'] += I * prod
            aux = (ap * wb + a0 * dpb) / 2
            d['tjpr'] += aux
            d['tjpp'] += I * aux
            aux = a0 * dpb - ap * wb
            d['gs'] += aux
            aux = ar * drb + ap * dpb + az * dzb
            d['tr'] += aux
            d['tp'] += I * aux
            aux = (a0 * drb - ar * wb) / 2
            d['tjrr'] += I * aux
            d['tjrp'] -= aux
            aux = a0 * drb + ar * wb
            d['gs'] += aux
            aux = (a0 * dzb - az * wb) / 2
            d['tjzr'] += I * aux
            d['tjzp'] -= aux
            aux = ar * (drb + 0.5 * dpb) - 0.5 * ap * drb
            d['fr'] += aux
            aux = ap * dpb + 0.5 * (ap * drb - ar * dpb)
            d['fp'] += I * aux
            aux = az * (drb + dpb) + ar * dzb - ap * dzb
            d['fz'] += aux / 2
        else:             # |a>=|->, |b>=|+>
            d['sr'] += prod
            d['sp'] -= I * prod
            aux = (ap * wb + a0 * dpb) / 2
            d['tjpr'] += aux
            d['tjpp'] -= I * aux
            aux = a0 * dpb - ap * wb
            d['gs'] -= aux
            aux = ar * drb + ap * dpb + az * dzb
            d['tr'] += aux
            d['tp'] -= I * aux
            aux = (a0 * drb - ar * wb) / 2
            d['tjrr'] += I * aux
            d['tjrp'] += aux
            aux = a0 * drb + ar * wb
            d['gs'] += aux
            aux = (a0 * dzb - az * wb) / 2
            d['tjzr'] += I * aux
            d['tjzp'] += aux
            aux = ar * (drb - 0.5 * dpb) + 0.5 * ap * drb
            d['fr'] += aux
            aux = -ap * dpb + 0.5 * (ap * drb - ar * dpb)
            d['fp'] += I * aux
            aux = az * (drb - dpb) + ar * dzb + ap * dzb
            d['fz'] += aux / 2
    # pairing densities
    K = wfa((wf,), rek, imk)
    for b in range(N):
        s = int(ns[b])
        wb = wf[:, b]
        prod = 2 * K[-s][0][:, b] * wb
        d['rb'] -= s * prod
        d['sbz'] += prod
        prod = 2 * K[s][0][:, b] * wb
        if s == -1:
            d['sbr'] += prod
            d['sbp'] -= I * prod    # dsresbp += im ; dsimsbp -= re
        else:
            d['sbr'] -= prod
            d['sbp'] -= I * prod
    w = model['wdcori']
    for k in d:
        d[k] = d[k] * w
    return d


def meanfield_tensor(model, ds):
    """The pointwise field tensor mf(ta,tb,sa,sb)(r) of meanfield (:775-1093), statement by
    statement; complex arithmetic replaces the (%re,%im) pairs.  Returns dict keyed (ta,tb,sa,sb)."""
    c = model
    crho, cs = c['crho'], c['cs']
    ctau, cj, ct, cdrho, cds, crdj, csdj = c['ctau'], c['cj'], c['ct'], c['cdrho'], c['cds'], c['crdj'], c['csdj']
    ctj0, ctj1, ctj2, cf, cgs = c['ctj0'], c['ctj1'], c['ctj2'], c['cf'], c['cgs']
    I = 1j
    Ng = model['nghl']
    mf = {}

    def add(key, sym, aux):
        mf[key] = mf[key] + sym * aux

    z = np.zeros(Ng, complex)
    for sa in (1, -1):
        for sb in (1, -1):
            for ta in range(4):
                for tb in range(4):
                    mf[(ta, tb, sa, sb)] = z.copy()
            mf[(0, 4, sa, sb)] = z.copy()
            mf[(4, 0, sa, sb)] = z.copy()
    t0 = ctj0 * (ds['tjrr'] + ds['tjpp'] + ds['tjzz'])
    t1_zr_rz = ctj1 * (ds['tjzr'] - ds['tjrz'])
    t1_pz_zp = ctj1 * (ds['tjpz'] - ds['tjzp'])
    t2_rz_zr = ctj2 * (ds['tjrz'] + ds['tjzr'])
    t2_pz_zp = ctj2 * (ds['tjpz'] + ds['tjzp'])
    P, M = 1, -1
    # wf_a, wf_b, same spin
    mf[(0, 0, P, P)] = 2.0 * crho * ds['rho'] + ctau * ds['tau']
    mf[(0, 0, M, M)] = mf[(0, 0, P, P)].copy()
    aux = 2.0 * cs * ds['sz'] + ct * ds['tz'] + cf * ds['fz']
    add((0, 0, P, P), 1, aux); add((0, 0, M, M), -1, aux)
    # (0,1)/(1,0) same spin
    mf[(0, 1, P, P)] = -crdj * (ds['tjpz'] - ds['tjzp'])
    mf[(0, 1, M, M)] = mf[(0, 1, P, P)].copy()
    aux = -csdj * ds['jp']
    add((0, 1, P, P), 1, aux); add((0, 1, M, M), -1, aux)
    mf[(1, 0, P, P)] = mf[(0, 1, P, P)].copy()
    mf[(1, 0, M, M)] = mf[(0, 1, M, M)].copy()
    aux = I * t1_zr_rz - 0.5 * I * t2_rz_zr     # re = -A.im + .5 B.im ; im = A.re - .5 B.re
    add((0, 1, P, P), 1, aux); add((0, 1, M, M), -1, aux); add((1, 0, P, P), -1, aux); add((1, 0, M, M), 1, aux)
    aux = -I * cj * ds['jr']                    # re = cj*im(jr) ; im = -cj*re(jr)
    add((0, 1, P, P), 1, aux); add((0, 1, M, M), 1, aux); add((1, 0, P, P), -1, aux); add((1, 0, M, M), -1, aux)
    # (0,2)/(2,0) same spin
    mf[(0, 2, P, P)] = cj * ds['jp']
    mf[(0, 2, M, M)] = mf[(0, 2, P, P)].copy()
    aux = t1_pz_zp + 0.5 * t2_pz_zp
    add((0, 2, P, P), 1, aux); add((0, 2, M, M), -1, aux)
    mf[(2, 0, P, P)] = mf[(0, 2, P, P)].copy()
    mf[(2, 0, M, M)] = mf[(0, 2, M, M)].copy()
    aux = I * csdj * ds['jr']                   # re = -csdj*im ; im = csdj*re
    add((0, 2, P, P), 1, aux); add((0, 2, M, M), -1, aux); add((2, 0, P, P), -1, aux); add((2, 0, M, M), 1, aux)
    aux = -I * crdj * (ds['tjzr'] - ds['tjrz'])  # re = crdj*im ; im = -crdj*re
    add((0, 2, P, P), 1, aux); add((0, 2, M, M), 1, aux); add((2, 0, P, P), -1, aux); add((2, 0, M, M), -1, aux)
    # (0,3)/(3,0) same spin
    mf[(0, 3, P, P)] = -crdj * (ds['tjrp'] - ds['tjpr'])
    mf[(0, 3, M, M)] = mf[(0, 3, P, P)].copy()
    aux = 2.0 * cgs * ds['gs']
    add((0, 3, P, P), 1, aux); add((0, 3, M, M), -1, aux)
    mf[(3, 0, P, P)] = mf[(0, 3, P, P)].copy()
    mf[(3, 0, M, M)] = mf[(0, 3, M, M)].copy()
    x = ds['tjrr'] + ds['tjpp'] - 2.0 * ds['tjzz']
    aux = -I * t0 + I * ctj2 * x / 3.0          # re = t0.im - ctj2*x.im/3 ; im = -t0.re + ctj2*x.re/3
    add((0, 3, P, P), 1, aux); add((0, 3, M, M), -1, aux); add((3, 0, P, P), -1, aux); add((3, 0, M, M), 1, aux)
    aux = -I * cj * ds['jz']
    add((0, 3, P, P), 1, aux); add((0, 3, M, M), 1, aux); add((3, 0, P, P), -1, aux); add((3, 0, M, M), -1, aux)
    # (0,4)/(4,0) same spin
    mf[(0, 4, P, P)] = 2.0 * cdrho * ds['rho']
    mf[(0, 4, M, M)] = mf[(0, 4, P, P)].copy()
    aux = 2.0 * cds * ds['sz']
    add((0, 4, P, P), 1, aux); add((0, 4, M, M), -1, aux)
    mf[(4, 0, P, P)] = mf[(0, 4, P, P)].copy()
    mf[(4, 0, M, M)] = mf[(0, 4, M, M)].copy()
    # (1,2)/(2,1) same spin
    mf[(1, 2, P, P)] = csdj * ds['sz']
    mf[(1, 2, M, M)] = mf[(1, 2, P, P)].copy()
    aux = crdj * ds['rho']
    add((1, 2, P, P), 1, aux); add((1, 2, M, M), -1, aux)
    mf[(2, 1, P, P)] = mf[(1, 2, P, P)].copy()
    mf[(2, 1, M, M)] = mf[(1, 2, M, M)].copy()
    # (1,3)/(3,1) same spin
    mf[(1, 3, P, P)] = 0.5 * cf * ds['sr']
    mf[(1, 3, M, M)] = -mf[(1, 3, P, P)]
    mf[(3, 1, P, P)] = mf[(1, 3, P, P)].copy()
    mf[(3, 1, M, M)] = mf[(1, 3, M, M)].copy()
    aux = I * csdj * ds['sp']                   # re = -csdj*im(sp) ; im = csdj*re(sp)
    add((1, 3, P, P), 1, aux); add((1, 3, M, M), 1, aux); add((3, 1, P, P), -1, aux); add((3, 1, M, M), -1, aux)
    # (2,3)/(3,2) same spin
    mf[(2, 3, P, P)] = -csdj * ds['sr']
    mf[(2, 3, M, M)] = mf[(2, 3, P, P)].copy()
    aux = -0.5 * I * cf * ds['sp']              # re = .5cf*im ; im = -.5cf*re
    add((2, 3, P, P), 1, aux); add((2, 3, M, M), -1, aux)
    mf[(3, 2, P, P)] = mf[(2, 3, M, M)].copy()
    mf[(3, 2, M, M)] = mf[(2, 3, P, P)].copy()
    # (1,1),(2,2),(3,3) same spin
    mf[(1, 1, P, P)] = (4.0 * cdrho + ctau) * ds['rho']
    mf[(1, 1, M, M)] = mf[(1, 1, P, P)].copy()
    aux = (4.0 * cds + ct) * ds['sz']
    add((1, 1, P, P), 1, aux); add((1, 1, M, M), -1, aux)
    mf[(2, 2, P, P)] = mf[(1, 1, P, P)].copy()
    mf[(2, 2, M, M)] = mf[(1, 1, M, M)].copy()
    mf[(3, 3, P, P)] = mf[(1, 1, P, P)].copy()
    mf[(3, 3, M, M)] = mf[(1, 1, M, M)].copy()
    aux = cf * ds['sz']
    add((3, 3, P, P), 1, aux); add((3, 3, M, M), -1, aux)
    # ---- opposite spin ----
    mf[(0, 0, P, M)] = 2.0 * cs * ds['sr'] + ct * ds['tr'] + cf * ds['fr']
    mf[(0, 0, M, P)] = mf[(0, 0, P, M)].copy()
    aux = -I * (2.0 * cs * ds['sp'] + ct * ds['tp'] + cf * ds['fp'])   # re = (..).im ; im = -(..).re
    add((0, 0, P, M), 1, aux); add((0, 0, M, P), -1, aux)
    mf[(0, 1, P, M)] = 2.0 * cgs * ds['gs']
    mf[(0, 1, M, P)] = mf[(0, 1, P, M)].copy()
    aux = -I * csdj * ds['jz']
    add((0, 1, P, M), 1, aux); add((0, 1, M, P), -1, aux)
    mf[(1, 0, P, M)] = mf[(0, 1, P, M)].copy()
    mf[(1, 0, M, P)] = mf[(0, 1, M, P)].copy()
    mf[(0, 2, P, M)] = mf[(0, 1, P, M)].copy()
    mf[(2, 0, M, P)] = mf[(1, 0, M, P)].copy()
    mf[(2, 0, P, M)] = -mf[(0, 2, P, M)]
    mf[(0, 2, M, P)] = -mf[(2, 0, M, P)]
    aux = -ctj1 * (ds['tjrp'] - ds['tjpr'])
    add((0, 1, P, M), 1, aux); add((0, 1, M, P), -1, aux); add((1, 0, P, M), -1, aux); add((1, 0, M, P), 1, aux)
    add((0, 2, P, M), 1, aux); add((0, 2, M, P), 1, aux); add((2, 0, P, M), 1, aux); add((2, 0, M, P), 1, aux)
    aux = -I * t0
    add((0, 1, P, M), 1, aux); add((0, 1, M, P), 1, aux); add((1, 0, P, M), -1, aux); add((1, 0, M, P), -1, aux)
    add((0, 2, P, M), 1, aux); add((0, 2, M, P), -1, aux); add((2, 0, P, M), 1, aux); add((2, 0, M, P), -1, aux)
    aux = -0.5 * ctj2 * (ds['tjrp'] + ds['tjpr'])
    add((0, 1, P, M), 1, aux); add((0, 1, M, P), -1, aux); add((1, 0, P, M), -1, aux); add((1, 0, M, P), 1, aux)
    add((0, 2, P, M), -1, aux); add((0, 2, M, P), -1, aux); add((2, 0, P, M), -1, aux); add((2, 0, M, P), -1, aux)
    aux = I * ctj2 * (-2.0 * ds['tjrr'] + ds['tjpp'] + ds['tjzz']) / 3.0   # re = -ctj2*x.im/3 ; im = ctj2*x.re/3
    add((0, 1, P, M), 1, aux); add((0, 1, M, P), 1, aux); add((1, 0, P, M), -1, aux); add((1, 0, M, P), -1, aux)
    aux = I * ctj2 * (ds['tjrr'] - 2.0 * ds['tjpp'] + ds['tjzz']) / 3.0
    add((0, 2, P, M), 1, aux); add((0, 2, M, P), -1, aux); add((2, 0, P, M), 1, aux); add((2, 0, M, P), -1, aux)
    # (0,3)/(3,0) opposite spin
    mf[(0, 3, P, M)] = csdj * ds['jp']
    mf[(0, 3, M, P)] = mf[(0, 3, P, M)].copy()
    aux = I * csdj * ds['jr']
    add((0, 3, P, M), 1, aux); add((0, 3, M, P), -1, aux)
    mf[(3, 0, P, M)] = mf[(0, 3, P, M)].copy()
    mf[(3, 0, M, P)] = mf[(0, 3, M, P)].copy()
    aux = -I * t1_zr_rz - 0.5 * I * t2_rz_zr    # re = A.im + .5B.im ; im = -A.re - .5B.re
    add((0, 3, P, M), 1, aux); add((0, 3, M, P), 1, aux); add((3, 0, P, M), -1, aux); add((3, 0, M, P), -1, aux)
    aux = t1_pz_zp - 0.5 * t2_pz_zp
    add((0, 3, P, M), 1, aux); add((0, 3, M, P), -1, aux); add((3, 0, P, M), -1, aux); add((3, 0, M, P), 1, aux)
    # (0,4)/(4,0) opposite spin
    mf[(0, 4, P, M)] = 2.0 * cds * ds['sr']
    mf[(0, 4, M, P)] = mf[(0, 4, P, M)].copy()
    aux = -I * 2.0 * cds * ds['sp']             # re = 2cds*im(sp) ; im = -2cds*re(sp)
    add((0, 4, P, M), 1, aux); add((0, 4, M, P), -1, aux)
    mf[(4, 0, P, M)] = mf[(0, 4, P, M)].copy()
    mf[(4, 0, M, P)] = mf[(0, 4, M, P)].copy()
    # (1,2)/(2,1) opposite spin
    mf[(1, 2, P, M)] = 0.5 * I * cf * ds['sp']  # re = -.5cf*im ; im = .5cf*re
    mf[(1, 2, M, P)] = mf[(1, 2, P, M)].copy()
    aux = 0.5 * cf * ds['sr']
    add((1, 2, P, M), 1, aux); add((1, 2, M, P), -1, aux)
    mf[(2, 1, P, M)] = -mf[(1, 2, P, M)]
    mf[(2, 1, M, P)] = -mf[(1, 2, M, P)]
    # (1,3),(3,1),(2,3),(3,2) opposite spin
    mf[(1, 3, P, M)] = 0.5 * cf * ds['sz']
    mf[(1, 3, M, P)] = mf[(1, 3, P, M)].copy()
    aux = crdj * ds['rho']
    add((1, 3, P, M), 1, aux); add((1, 3, M, P), -1, aux)
    mf[(3, 1, P, M)] = mf[(1, 3, M, P)].copy()
    mf[(3, 1, M, P)] = mf[(1, 3, P, M)].copy()
    mf[(3, 2, P, M)] = mf[(3, 1, P, M)].copy()
    mf[(2, 3, M, P)] = mf[(1, 3, M, P)].copy()
    mf[(2, 3, P, M)] = -mf[(1, 3, P, M)]
    mf[(3, 2, M, P)] = -mf[(3, 1, M, P)]
    # (3,3),(1,1),(2,2) opposite spin
    mf[(3, 3, P, M)] = (ct + 4.0 * cds) * ds['sr']
    mf[(3, 3, M, P)] = mf[(3, 3, P, M)].copy()
    aux = -I * (ct + 4.0 * cds) * ds['sp']      # re = (..)*im(sp) ; im = -(..)*re(sp)
    add((3, 3, P, M), 1, aux); add((3, 3, M, P), -1, aux)
    mf[(1, 1, P, M)] = mf[(3, 3, P, M)].copy()
    mf[(1, 1, M, P)] = mf[(3, 3, M, P)].copy()
    aux = cf * ds['sr']
    add((1, 1, P, M), 1, aux); add((1, 1, M, P), 1, aux)
    mf[(2, 2, P, M)] = mf[(3, 3, P, M)].copy()
    mf[(2, 2, M, P)] = mf[(3, 3, M, P)].copy()
    aux = -I * cf * ds['sp']                    # re = cf*im(sp) ; im = -cf*re(sp)
    add((2, 2, P, M), 1, aux); add((2, 2, M, P), -1, aux)
    return mf


def meanfield(model, ds, reh, imh):
    """meanfield (:717-1169): fills reh/imh (block structure preset by the caller)."""
    db, nb, N, Ng = model['db'], model['nb'], model['dqp'], model['nghl']
    ns, nsu, isstart = model['ns'], model['num_spin_up'], _isstart(model['db'])
    tabs = (model['wf'], model['wfdr'], model['wfdp'], model['wfdz'], model['wfd2_all'])
    mf = meanfield_tensor(model, ds)
    reh.elem[:] = 0
    imh.elem[:] = 0
    hpsi = {(t, sa): np.zeros((Ng, N), complex) for t in range(5) for sa in (1, -1)}
    for b in range(N):
        sb = int(ns[b])
        for sa in (1, -1):
            for ty in range(5):
                acc = hpsi[(ty, sa)][:, b]
                for tz in range(5):
                    key = (ty, tz, sa, sb)
                    if key in mf:
                        acc += mf[key] * tabs[tz][:, b]
    for ix in range(nb):
        iy = int(reh.ir2c[ix])
        if iy == 0:
            continue
        ia, ib = isstart[ix] - 1, isstart[iy - 1] - 1
        d1, d2 = db[ix], db[iy - 1]
        nu = nsu[ix]
        blk = np.zeros((d1, d2), complex)
        for t in range(5):
            if nu > 0:
                blk[:nu, :] += 2.0 * tabs[t][:, ia:ia + nu].T @ hpsi[(t, 1)][:, ib:ib + d2]
            if d1 - nu > 0:
                blk[nu:, :] += 2.0 * tabs[t][:, ia + nu:ia + d1].T @ hpsi[(t, -1)][:, ib:ib + d2]
        o = reh.ir2m[ix] - 1
        reh.elem[o:o + d1 * d2] = blk.real.reshape(-1, order='F')
        imh.elem[o:o + d1 * d2] = blk.imag.reshape(-1, order='F')


def pairingfield(model, ds, red, imd):
    """pairingfield (:1175-1262)."""
    db, nb, N, Ng = model['db'], model['nb'], model['dqp'], model['nghl']
    ns, nsu, isstart, wf = model['ns'], model['num_spin_up'], _isstart(model['db']), model['wf']
    cpair, cspair = model['cpair'], model['cspair']
    I = 1j
    red.elem[:] = 0
    imd.elem[:] = 0
    aux_pp = cspair * (ds['sbr'] - I * ds['sbp'])      # re = cs*(sbr.re + sbp.im), im = cs*(sbr.im - sbp.re)
    aux_mp = -cpair * ds['rb'] - cspair * ds['sbz']
    aux_pm = cpair * ds['rb'] - cspair * ds['sbz']
    aux_mm = cspair * (-ds['sbr'] - I * ds['sbp'])     # re = cs*(-sbr.re + sbp.im), im = cs*(-sbr.im - sbp.re)
    dpsi = {1: np.zeros((Ng, N), complex), -1: np.zeros((Ng, N), complex)}
    for b in range(N):
        if ns[b] == 1:
            dpsi[1][:, b] = aux_pp * wf[:, b]
            dpsi[-1][:, b] = aux_mp * wf[:, b]
        else:
            dpsi[1][:, b] = aux_pm * wf[:, b]
            dpsi[-1][:, b] = aux_mm * wf[:, b]
    for ix in range(nb):
        iy = int(red.ir2c[ix])
        if iy == 0:
            continue
        ia, ib = isstart[ix] - 1, isstart[iy - 1] - 1
        d1, d2 = db[ix], db[iy - 1]
        nu = nsu[ix]
        blk = np.zeros((d1, d2), complex)
        if nu > 0:
            blk[:nu, :] = 2.0 * wf[:, ia:ia + nu].T @ dpsi[1][:, ib:ib + d2]
        if d1 - nu > 0:
            blk[nu:, :] = 2.0 * wf[:, ia + nu:ia + d1].T @ dpsi[-1][:, ib:ib + d2]
        o = red.ir2m[ix] - 1
        red.elem[o:o + d1 * d2] = blk.real.reshape(-1, order='F')
        imd.elem[o:o + d1 * d2] = blk.imag.reshape(-1, order='F')


def calc_hamiltonian(model, dRsp_re, dRsp_im, dHsp_re, dHsp_im):
    """calc_hamiltonian (:51-74) with the argument wiring of pnfam_solver.f90:152-157."""
    ds = density(model, (dRsp_re.m[11], dRsp_im.m[11]), (dRsp_re.m[12], dRsp_im.m[12]))
    meanfield(model, ds, dHsp_re.m[11], dHsp_im.m[11])
    pairingfield(model, ds, dHsp_re.m[12], dHsp_im.m[12])
    ds = density(model, (dRsp_re.m[22], dRsp_im.m[22]), (dRsp_re.m[21], dRsp_im.m[21]))
    meanfield(model, ds, dHsp_re.m[22], dHsp_im.m[22])
    pairingfield(model, ds, dHsp_re.m[21], dHsp_im.m[21])


# ------------------------------------------------------------------------------------------------
# Broyden
# ------------------------------------------------------------------------------------------------
class Broyden:
    """broyden_method (pnfam_broyden.f90:117-216), state kept between calls like the SAVEd arrays."""

    def __init__(self, n, M, alpha=ALPHAMIX):
        self.n, self.M, self.alpha = n, M, alpha
        self.df = self.dv = None
        self.w0 = 0.01
        self.label = 'N'

    def step(self, it, vout, vin):
        """vout: new output, vin: previous input.  Returns (si, mixed vin)."""
        alpha, M = self.alpha, self.M
        vout = vout - vin
        si = float(np.max(np.abs(vout)))
        if M < 0:
            self.label = 'N'
            return si, vin + vout
        if M == 0 or it == 0:
            self.label = 'L'
            return si, vin + alpha * vout
        self.label = 'B'
        iter_used = min(it - 1, M)
        ipos = it - 1 - int((it - 2) / M) * M      # Fortran integer division truncates toward zero
        inext = it - int((it - 1) / M) * M
        if it == 1:
            self.w0 = 0.010
            self.df = np.zeros((self.n, M))
            self.dv = np.zeros((self.n, M))
        else:
            self.df[:, ipos - 1] = vout - self.df[:, ipos - 1]
            self.dv[:, ipos - 1] = vin - self.dv[:, ipos - 1]
            normi = 1.0 / np.sqrt(np.linalg.norm(self.df[:, ipos - 1]) ** 2)
            self.df[:, ipos - 1] *= normi
            self.dv[:, ipos - 1] *= normi
        curv = alpha * vout
        if iter_used > 0:
            df = self.df[:, :iter_used]
            beta = df.T @ df
            beta[np.diag_indices(iter_used)] = self.w0 * self.w0 + 1.0
            beta = np.linalg.inv(beta)
            work = df.T @ vout
            for i in range(iter_used):
                gamma = float(beta[:, i] @ work)
                curv = curv - gamma * (self.dv[:, i] + alpha * self.df[:, i])
        self.df[:, inext - 1] = vout
        self.dv[:, inext - 1] = vin
        return si, vin + curv


# ------------------------------------------------------------------------------------------------
# the solver
# ------------------------------------------------------------------------------------------------
def matrix_2qp(db, mat, a, b, c, d, e, f1, f2):
    """matrix_2qp (pnfam_solver.f90:510-544): complex M_ij = a*(b*f1_i + c*f2_j + e)**d on mat's blocks."""
    isstart = _isstart(db)
    out = np.zeros(len(mat.elem), complex)
    ipt = 0
    for ibr in range(len(db)):
        ibc = int(mat.ir2c[ibr])
        if ibc == 0:
            continue
        i1 = np.arange(isstart[ibr] - 1, isstart[ibr] - 1 + db[ibr])
        i2 = np.arange(isstart[ibc - 1] - 1, isstart[ibc - 1] - 1 + db[ibc - 1])
        blk = a * (b * f1[i1][:, None] + c * f2[i2][None, :] + e) ** d
        n = blk.size
        out[ipt:ipt + n] = blk.reshape(-1, order='F')
        ipt += n
    return out


class FamSolver:
    """State of pnfam_solve for one omega (pnfam_solver.f90:32-460)."""

    def __init__(self, model, f_ir2c, f_elem, g_list=(), beta_minus=True, omega=0j, quench=1.0,
                 broyden_history=50, energy_shift_prot=0.0, energy_shift_neut=0.0):
        self.model = m = model
        db, nb = m['db'], m['nb']
        self.db = db
        nxy = len(f_elem)
        self.nxy = nxy
        self.bminus = beta_minus
        self.quench = quench
        self.use_diag = bool(m.get('blo_active', False))
        a11 = a22 = self.use_diag
        Ep = m['Ep'] + energy_shift_prot
        En = m['En'] + energy_shift_neut
        # U, V block matrices (pnfam_setup.f90:292-321)
        U = block_diag_structure(db)
        V = BlockMatrix(nb)
        h = nb // 2
        V.ir2c = np.concatenate([np.arange(h) + h + 1, np.arange(h) + 1])
        V.ic2r = V.ir2c.copy()
        V.ir2m = U.ir2m.copy()
        V.ic2m = V.ir2m[V.ic2r - 1]

        def mk(struct, elem):
            b = struct.copy()
            b.elem = np.array(elem, float)
            return b
        self.W = {}
        for t in ('n', 'p'):
            W = BBM()
            W.m[11] = mk(U, m['U' + t]); W.m[12] = mk(V, m['V' + t])
            W.m[21] = mk(V, m['V' + t]); W.m[22] = mk(U, m['U' + t])
            self.W[t] = W
        Wn, Wp = self.W['n'], self.W['p']
        Fsp = BBM()
        Fsp.m[11] = BlockMatrix.from_rows(db, f_ir2c, f_elem)

        def new_qp(imag=False):
            x = BBM()
            x.allocate(nb, a11, True, True, a22, nxy)
            x.set_sign(1, 1, -1, -1)
            x.imag = imag
            return x
        self.Fqp = new_qp()
        self.Gqp = []
        self.dHqp_re, self.dHqp_im = new_qp(), new_qp(True)
        self.dRqp_re, self.dRqp_im = new_qp(), new_qp(True)
        self.dHsp_re, self.dHsp_im = BBM(), BBM()
        for x, sg in ((self.dHsp_re, (1, 1, -1, -1)), (self.dHsp_im, (1, 1, 1, -1))):
            x.allocate(nb, True, True, True, True, nxy)
            x.set_sign(*sg)
            x.set_trans('n', 'n', 'n', 't')
        self.dHsp_im.imag = True
        self.dRsp_re, self.dRsp_im = BBM(), BBM()
        for x, sg in ((self.dRsp_re, (1, -1, 1, -1)), (self.dRsp_im, (1, -1, -1, -1))):
            x.allocate(nb, True, True, True, True, nxy)
            x.set_sign(*sg)
            x.set_trans('n', 't', 't', 't')
        self.dRsp_im.imag = True
        if self.bminus:
            triprod_bbm(db, 't', Wp, 'n', Fsp, 'n', Wn, self.Fqp)
        else:
            triprod_bbm(db, 't', Wn, 'n', Fsp, 'n', Wp, self.Fqp)
        for (g_ir2c, g_elem) in g_list:
            Gsp = BBM()
            Gsp.m[11] = BlockMatrix.from_rows(db, g_ir2c, g_elem)
            G = new_qp()
            if self.bminus:
                triprod_bbm(db, 't', Wp, 'n', Gsp, 'n', Wn, G)
            else:
                triprod_bbm(db, 't', Wn, 'n', Gsp, 'n', Wp, G)
            self.Gqp.append(G)
        for x in (self.dRqp_re, self.dRqp_im, self.dHqp_re, self.dHqp_im):
            for q in (11, 12, 21, 22):
                if x.m[q] is not None:
                    x.m[q].copy_structure_from(self.Fqp.m[q])
        for x in (self.dHsp_re, self.dHsp_im):
            x.m[11].copy_structure_from(Fsp.m[11])
            x.m[12].copy_structure_from(self.dRqp_re.m[12])
            x.m[21].copy_structure_from(self.dRqp_re.m[21])
            x.m[22].copy_structure_from(Fsp.m[11])
        # Greens function and T (pnfam_solver.f90:415-458)
        f1, f2 = (Ep, En) if self.bminus else (En, Ep)
        w = complex(omega)
        self.G = {12: matrix_2qp(db, self.dRqp_re.m[12], -1, 1, 1, -1, -w, f1, f2),
                  21: matrix_2qp(db, self.dRqp_re.m[21], -1, 1, 1, -1, +w, f1, f2)}
        self.T = None
        if self.use_diag:
            self.G[11] = matrix_2qp(db, self.dRqp_re.m[11], -1, 1, -1, -1, -w, f1, f2)
            self.G[22] = matrix_2qp(db, self.dRqp_re.m[22], -1, 1, -1, -1, +w, f1, f2)
            q1, q2 = (m['qp_fp'], m['qp_fn']) if self.bminus else (m['qp_fn'], m['qp_fp'])
            self.T = {12: matrix_2qp(db, self.dRqp_re.m[12], 1, -1, -1, 1, 1.0, q1, q2).real,
                      21: matrix_2qp(db, self.dRqp_re.m[21], 1, -1, -1, 1, 1.0, q1, q2).real,
                      11: matrix_2qp(db, self.dRqp_re.m[11], 1, -1, 1, 1, 0.0, q1, q2).real,
                      22: matrix_2qp(db, self.dRqp_re.m[22], 1, -1, 1, 1, 0.0, q1, q2).real}
        nvec = 8 if self.use_diag else 4
        M = broyden_history if abs(quench) >= 1e-10 else -1
        self.bro = Broyden(nvec * nxy, M)
        self.broin = np.zeros(nvec * nxy)
        self.si = 1.0
        self.str = np.zeros(1 + len(self.Gqp), complex)
        self.trace = []

    # pack order of pnfam_broyden.f90:50-60
    def _pack(self):
        r, i = self.dRqp_re, self.dRqp_im
        v = [r.m[12].elem, r.m[21].elem, i.m[12].elem, i.m[21].elem]
        if self.use_diag:
            v += [r.m[11].elem, r.m[22].elem, i.m[11].elem, i.m[22].elem]
        return np.concatenate(v)

    def _unpack(self, v):
        n = self.nxy
        r, i = self.dRqp_re, self.dRqp_im
        r.m[12].elem, r.m[21].elem = v[0:n].copy(), v[n:2 * n].copy()
        i.m[12].elem, i.m[21].elem = v[2 * n:3 * n].copy(), v[3 * n:4 * n].copy()
        if self.use_diag:
            r.m[11].elem, r.m[22].elem = v[4 * n:5 * n].copy(), v[5 * n:6 * n].copy()
            i.m[11].elem, i.m[22].elem = v[6 * n:7 * n].copy(), v[7 * n:8 * n].copy()

    def iterate(self, it):
        db = self.db
        Wn, Wp = self.W['n'], self.W['p']
        if abs(self.quench) < 1e-10:
            self.dHqp_re.set_val(0.0)
            self.dHqp_im.set_val(0.0)
        else:
            if self.bminus:
                triprod_bbm(db, 'n', Wp, 'n', self.dRqp_re, 't', Wn, self.dRsp_re)
                triprod_bbm(db, 'n', Wp, 'n', self.dRqp_im, 't', Wn, self.dRsp_im)
            else:
                triprod_bbm(db, 'n', Wn, 'n', self.dRqp_re, 't', Wp, self.dRsp_re)
                triprod_bbm(db, 'n', Wn, 'n', self.dRqp_im, 't', Wp, self.dRsp_im)
            calc_hamiltonian(self.model, self.dRsp_re, self.dRsp_im, self.dHsp_re, self.dHsp_im)
            if self.bminus:
                triprod_bbm(db, 't', Wp, 'n', self.dHsp_re, 'n', Wn, self.dHqp_re)
                triprod_bbm(db, 't', Wp, 'n', self.dHsp_im, 'n', Wn, self.dHqp_im)
            else:
                triprod_bbm(db, 't', Wn, 'n', self.dHsp_re, 'n', Wp, self.dHqp_re)
                triprod_bbm(db, 't', Wn, 'n', self.dHsp_im, 'n', Wp, self.dHqp_im)
            for x in (self.dHqp_re, self.dHqp_im):
                for q in x.m:
                    if x.m[q] is not None:
                        x.m[q].elem *= self.quench
        for q in (11, 12, 21, 22):
            if self.Fqp.m[q] is not None:
                self.dHqp_re.m[q].elem += self.Fqp.m[q].elem
        for q in (11, 12, 21, 22):
            if self.dRqp_re.m[q] is None:
                continue
            z = self.G[q] * (self.dHqp_re.m[q].elem + 1j * self.dHqp_im.m[q].elem)
            if self.T is not None:
                z = self.T[q] * z
            self.dRqp_re.m[q].elem = z.real.copy()
            self.dRqp_im.m[q].elem = z.imag.copy()
        self.si, self.broin = self.bro.step(it, self._pack(), self.broin)
        self._unpack(self.broin)
        for k, F in enumerate([self.Fqp] + self.Gqp):
            re_s = contract_bbm(F, self.dRqp_re)
            im_s = contract_bbm(F, self.dRqp_im)
            self.str[k] = complex(-re_s / np.pi, -im_s / np.pi)

    def solve(self, max_iter=200, eps=1e-7):
        """The loop of ifam (:114-209).  Returns (iter_conv, si, strengths)."""
        it = 0
        while True:
            self.trace.append((it, self.bro.label, self.si, self.str[0]))
            if self.si < eps:
                return it, self.si, self.str.copy()
            if it == max_iter:
                return -1, self.si, self.str.copy()
            self.iterate(it)
            it += 1


# ------------------------------------------------------------------------------------------------
# glue: build the `model` dict and a solver from a pynfam_b200.host.Problem
# ------------------------------------------------------------------------------------------------
def model_from_problem(p):
    m = {k: p.iscalar(k) for k in ('nb', 'dqp', 'nghl')}
    for k in ('db', 'ns', 'nl', 'num_spin_up'):
        m[k] = p.i32(k).astype(np.int64)
    for k in ('wf', 'wfdr', 'wfdp', 'wfdz', 'wfd2_all'):
        m[k] = p.table(k)
    for k in ('wdcori', 'y', 'z', 'Ep', 'En', 'Up', 'Vp', 'Un', 'Vn', 'crho', 'cs', 'cpair', 'cspair'):
        m[k] = p.f64(k)
    for k in ('cdrho', 'ctau', 'ctj0', 'ctj1', 'ctj2', 'crdj', 'cds', 'ct', 'cj', 'cgs', 'cf', 'csdj'):
        m[k] = p.scalar(k)
    m['blo_active'] = bool(p.iscalar('statistical'))   # equal filling or finite temperature: P,Q quadrants + T factors
    if m['blo_active']:
        m['qp_fn'], m['qp_fp'] = p.f64('qp_fn'), p.f64('qp_fp')
    return m


def solver_from_problem(p, model=None, omega=None):
    m = model or model_from_problem(p)
    g = [(p.i32('g_ir2c_%d' % i), p.f64('g_elem_%d' % i)) for i in range(p.iscalar('nxterms'))]
    w = omega if omega is not None else complex(p.scalar('real_eqrpa'), p.scalar('imag_eqrpa'))
    return FamSolver(m, p.i32('f_ir2c'), p.f64('f_elem'), g, beta_minus=bool(p.iscalar('beta_minus')), omega=w,
                     quench=p.scalar('quench_residual_int'), broyden_history=p.iscalar('broyden_history_size'),
                     energy_shift_prot=p.scalar('energy_shift_prot'), energy_shift_neut=p.scalar('energy_shift_neut'))
