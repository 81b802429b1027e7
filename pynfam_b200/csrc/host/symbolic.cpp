#include "symbolic.hpp"

#include <stdexcept>

namespace pnfam {
namespace {

// Structure-only block matrix; 0-based with -1 for "no block".
struct SymBM {
  std::vector<int> r2c, c2r, r2m, c2m;
  bool alloc = false;
  int ident = -1;  // which W matrix (0..3) or which middle-operand quadrant (0..3 in bbm order 11,12,21,22)
  void init(int nb) { r2c.assign(nb, -1); c2r.assign(nb, -1); r2m.assign(nb, -1); c2m.assign(nb, -1); alloc = true; }
};

struct SymBBM {
  SymBM m[4];                                   // 11, 12, 21, 22
  char t[4] = {'n', 'n', 'n', 'n'};
  double s[4] = {1, 1, 1, 1};
  bool imag = false;
  void transpose() {                            // transpose_bbm, pnfam_type_bbm.f90:145-156
    std::swap(m[1], m[2]);
    std::swap(s[1], s[2]);
    for (char& c : t) c = (c == 'n') ? 't' : 'n';
  }
};

SymBM diag_struct(const std::vector<int>& db, int ident) {
  const int nb = (int)db.size();
  SymBM u; u.init(nb); u.ident = ident;
  int off = 0;
  for (int i = 0; i < nb; i++) { u.r2c[i] = u.c2r[i] = i; u.r2m[i] = u.c2m[i] = off; off += db[i] * db[i]; }
  return u;
}
SymBM v_struct(const std::vector<int>& db, int ident) {  // pnfam_setup.f90:312-321
  const int nb = (int)db.size(), h = nb / 2;
  SymBM u = diag_struct(db, ident), v; v.init(nb); v.ident = ident;
  for (int i = 0; i < nb; i++) { v.r2c[i] = i < h ? i + h : i - h; }
  for (int i = 0; i < nb; i++) v.c2r[i] = v.r2c[i];
  for (int i = 0; i < nb; i++) v.r2m[i] = u.r2m[i];
  for (int i = 0; i < nb; i++) v.c2m[i] = v.r2m[v.c2r[i]];
  return v;
}
SymBBM w_bbm(const std::vector<int>& db, int u_id, int v_id) {  // W = [U V; V U], pnfam_solver.f90:297-303
  SymBBM w;
  w.m[0] = diag_struct(db, u_id); w.m[1] = v_struct(db, v_id);
  w.m[2] = v_struct(db, v_id);    w.m[3] = diag_struct(db, u_id);
  return w;
}
SymBM from_rows(const std::vector<int>& db, const std::vector<int>& ir2c_1based, int ident) {
  const int nb = (int)db.size();
  SymBM b; b.init(nb); b.ident = ident;
  int off = 0;
  for (int i = 0; i < nb; i++) {
    const int j = ir2c_1based[i] - 1;
    if (j < 0) continue;
    b.r2c[i] = j; b.c2r[j] = i; b.r2m[i] = off; b.c2m[j] = off;
    off += db[i] * db[j];
  }
  return b;
}
SymBM from_struct(const BlockStruct& s, int nb, int ident) {
  SymBM b; b.init(nb); b.ident = ident; b.alloc = s.allocated;
  if (!s.allocated) return b;
  for (int i = 0; i < nb; i++) {
    b.r2c[i] = s.r2c[i]; b.r2m[i] = s.r2m[i];
    if (s.r2c[i] >= 0) { b.c2r[s.r2c[i]] = i; b.c2m[s.r2c[i]] = s.r2m[i]; }
  }
  return b;
}

struct RowTerm { int i, j, ipa, ipb, ipc; };

// Structure part of triprod (pnfam_type_blockmatrix.f90:140-212): which blocks meet in row i.
std::vector<RowTerm> sym_triprod(int nb, char ta, const SymBM& a, char tb, const SymBM& b, char tc, const SymBM& c) {
  std::vector<RowTerm> out;
  for (int i = 0; i < nb; i++) {
    int k, ipa, l, ipb, j, ipc;
    if (ta == 'n') { k = a.r2c[i]; ipa = a.r2m[i]; } else { k = a.c2r[i]; ipa = a.c2m[i]; }
    if (k < 0) continue;
    if (tb == 'n') { l = b.r2c[k]; ipb = b.r2m[k]; } else { l = b.c2r[k]; ipb = b.c2m[k]; }
    if (l < 0) continue;
    if (tc == 'n') { j = c.r2c[l]; ipc = c.r2m[l]; } else { j = c.c2r[l]; ipc = c.c2m[l]; }
    if (j < 0) continue;
    out.push_back({i, j, ipa, ipb, ipc});
  }
  return out;
}

const int kTerms[4][4][3] = {
    // output quadrant 11, 12, 21, 22 : (a quadrant, b quadrant, c quadrant), bbm order
    {{0, 0, 0}, {1, 2, 0}, {0, 1, 2}, {1, 3, 2}},
    {{0, 0, 1}, {1, 2, 1}, {0, 1, 3}, {1, 3, 3}},
    {{2, 0, 0}, {3, 2, 0}, {2, 1, 2}, {3, 3, 2}},
    {{2, 0, 1}, {3, 2, 1}, {2, 1, 3}, {3, 3, 3}},
};

struct QuadTerm { int qa, qb, qc; char ta, tb, tc; double alpha; };

// triprod_bbm_quad (pnfam_type_bbm.f90:428-552), structure + coefficient only.
std::vector<QuadTerm> quad_terms(char ta, SymBBM a, char tb, SymBBM b, char tc, SymBBM c, int qabc, double sabc, char tabc,
                                 SymBBM* ao, SymBBM* bo, SymBBM* co) {
  int quadrant = qabc;
  double s = sabc;
  if (ta == 't') a.transpose();
  if (tb == 't') b.transpose();
  if (tc == 't') c.transpose();
  if (tabc == 't') {
    std::swap(a, c);
    a.transpose(); b.transpose(); c.transpose();
    if (quadrant == 1) quadrant = 2; else if (quadrant == 2) quadrant = 1;
  }
  const int im = (int)a.imag + (int)b.imag + (int)c.imag;
  if (im >= 2) s = -s;
  std::vector<QuadTerm> out;
  for (int t = 0; t < 4; t++) {
    const int qa = kTerms[quadrant][t][0], qb = kTerms[quadrant][t][1], qc = kTerms[quadrant][t][2];
    if (a.m[qa].alloc && b.m[qb].alloc && c.m[qc].alloc)
      out.push_back({qa, qb, qc, a.t[qa], b.t[qb], c.t[qc], s * a.s[qa] * b.s[qb] * c.s[qc]});
  }
  *ao = a; *bo = b; *co = c;
  return out;
}

// storage index of a bbm quadrant (11,12,21,22 -> 0..3) for qp / sp objects
const int kQpStore[4] = {2, 0, 1, 3};
const int kSpStore[4] = {0, 1, 2, 3};

// Flatten  abc = op(A) op(B) op(C)  for all allocated output quadrants, for the re and im flavour.
TransformPlan make_plan(const std::vector<int>& db, char ta, const SymBBM& A, char tb, const SymBBM& B_re,
                        const SymBBM& B_im, char tc, const SymBBM& C, const bool out_alloc[4], const double s_re[4],
                        const double s_im[4], const char t_out[4], const int out_store[4], const int b_store[4]) {
  const int nb = (int)db.size();
  TransformPlan plan;
  for (int q = 0; q < 4; q++) {
    if (!out_alloc[q]) continue;
    SymBBM a1, b1, c1, a2, b2, c2;
    auto tr = quad_terms(ta, A, tb, B_re, tc, C, q, s_re[q], t_out[q], &a1, &b1, &c1);
    auto ti = quad_terms(ta, A, tb, B_im, tc, C, q, s_im[q], t_out[q], &a2, &b2, &c2);
    if (tr.size() != ti.size()) throw std::runtime_error("symbolic: re/im term mismatch");
    BlockStruct& os = plan.out[out_store[q]];
    os.r2c.assign(nb, -1); os.r2m.assign(nb, -1); os.allocated = true;
    std::vector<int> task_of_row(nb, -1);
    for (size_t t = 0; t < tr.size(); t++) {
      const QuadTerm& x = tr[t];
      const SymBM& am = a1.m[x.qa]; const SymBM& bm = b1.m[x.qb]; const SymBM& cm = c1.m[x.qc];
      auto rows = sym_triprod(nb, x.ta, am, x.tb, bm, x.tc, cm);
      int ipt = 0;
      size_t nrows_expected = 0;
      for (int i = 0; i < nb; i++) if (task_of_row[i] >= 0) nrows_expected++;
      if (t > 0 && rows.size() != nrows_expected) throw std::runtime_error("symbolic: inconsistent block structure between terms");
      for (const RowTerm& r : rows) {
        const int m = db[r.i], n = db[r.j];
        if (t == 0) {
          os.r2c[r.i] = r.j; os.r2m[r.i] = ipt;
          BlockTask bt{};
          bt.out_quad = out_store[q]; bt.out_off = ipt; bt.m = m; bt.n = n; bt.nterms = 0;
          task_of_row[r.i] = (int)plan.tasks.size();
          plan.tasks.push_back(bt);
        } else if (os.r2c[r.i] != r.j || os.r2m[r.i] != ipt) {
          throw std::runtime_error("symbolic: inconsistent block structure between terms");
        }
        BlockTask& bt = plan.tasks[task_of_row[r.i]];
        TripleTerm& tt = bt.t[bt.nterms++];
        tt.a_mat = am.ident; tt.a_off = r.ipa; tt.a_trans = x.ta != 'n';
        tt.b_quad = b_store[bm.ident]; tt.b_off = r.ipb; tt.b_trans = x.tb != 'n';
        tt.c_mat = cm.ident; tt.c_off = r.ipc; tt.c_trans = x.tc != 'n';
        tt.alpha_re = x.alpha; tt.alpha_im = ti[t].alpha;
        ipt += m * n;
      }
      if (t == 0) os.nelem = (size_t)ipt;
    }
  }
  return plan;
}

}  // namespace

OperatorPlan make_operator_plan(const std::vector<int>& db, const std::vector<int>& f_ir2c, bool use_diag, bool beta_minus) {
  OperatorPlan op;
  const int nb = (int)db.size();
  op.nb = nb; op.beta_minus = beta_minus; op.use_diag = use_diag;
  op.nxy = 0;
  for (int i = 0; i < nb; i++) if (f_ir2c[i] > 0) op.nxy += (size_t)db[i] * db[f_ir2c[i] - 1];
  // a = row-side isospin, b = column-side (pnfam_solver.f90:144-166): W ids 0 Ua, 1 Va, 2 Ub, 3 Vb
  SymBBM Wa = w_bbm(db, 0, 1), Wb = w_bbm(db, 2, 3);
  {
    SymBM u = diag_struct(db, 0);
    op.u_off = u.r2m; op.v_off = u.r2m;
  }
  const bool qp_alloc[4] = {op.use_diag, true, true, op.use_diag};   // bbm order 11,12,21,22
  const bool all_alloc[4] = {true, true, true, true};
  const double s_qp[4] = {1, 1, -1, -1};
  const char t_n[4] = {'n', 'n', 'n', 'n'};
  // --- structure of Fqp:  triprod_bbm('t',Wa,'n',Fsp,'n',Wb,Fqp), Fsp = [f 0; 0 0] --------------
  SymBBM Fsp;
  Fsp.m[0] = from_rows(db, f_ir2c, 0);
  {
    TransformPlan fp = make_plan(db, 't', Wa, 'n', Fsp, Fsp, 'n', Wb, qp_alloc, s_qp, s_qp, t_n, kQpStore, kSpStore);
    for (int k = 0; k < 4; k++) op.qp[k] = fp.out[k];
  }
  // --- forward: triprod_bbm('n',Wa,'n',dRqp,'t',Wb,dRsp) ---------------------------------------
  SymBBM Rre, Rim;
  for (int q = 0; q < 4; q++) {
    if (!qp_alloc[q]) continue;
    Rre.m[q] = from_struct(op.qp[kQpStore[q]], nb, q);
    Rre.s[q] = s_qp[q];
  }
  Rim = Rre; Rim.imag = true;
  {
    const double s_re[4] = {1, -1, 1, -1}, s_im[4] = {1, -1, -1, -1};
    const char t_out[4] = {'n', 't', 't', 't'};
    op.forward = make_plan(db, 'n', Wa, 'n', Rre, Rim, 't', Wb, all_alloc, s_re, s_im, t_out, kSpStore, kQpStore);
    for (int k = 0; k < 4; k++) op.sp[k] = op.forward.out[k];
  }
  // --- backward: triprod_bbm('t',Wa,'n',dHsp,'n',Wb,dHqp) --------------------------------------
  // dHsp block structures are PRESET by the caller (pnfam_solver.f90:402-413): h_pn and h_np take
  // Fsp's, Delta+ takes X's, Delta- takes Y's.
  SymBBM Hre, Him;
  Hre.m[0] = from_rows(db, f_ir2c, 0);
  Hre.m[1] = from_struct(op.qp[0], nb, 1);
  Hre.m[2] = from_struct(op.qp[1], nb, 2);
  Hre.m[3] = from_rows(db, f_ir2c, 3);
  const double sh_re[4] = {1, 1, -1, -1}, sh_im[4] = {1, 1, 1, -1};
  const char th[4] = {'n', 'n', 'n', 't'};
  for (int q = 0; q < 4; q++) { Hre.s[q] = sh_re[q]; Hre.t[q] = th[q]; }
  Him = Hre; Him.imag = true;
  for (int q = 0; q < 4; q++) Him.s[q] = sh_im[q];
  for (int q = 0; q < 4; q++) {
    const SymBM& m = Hre.m[q];
    BlockStruct& h = op.hsp[q];
    h.r2c = m.r2c; h.r2m = m.r2m; h.allocated = true; h.nelem = 0;
    for (int i = 0; i < nb; i++) if (m.r2c[i] >= 0) h.nelem += (size_t)db[i] * db[m.r2c[i]];
  }
  op.backward = make_plan(db, 't', Wa, 'n', Hre, Him, 'n', Wb, qp_alloc, s_qp, s_qp, t_n, kQpStore, kSpStore);
  for (int k = 0; k < 4; k++) {
    if (!op.qp[k].allocated) continue;
    if (op.backward.out[k].r2c != op.qp[k].r2c || op.backward.out[k].r2m != op.qp[k].r2m)
      throw std::runtime_error("symbolic: dHqp structure differs from Fqp structure");
  }
  // the hamiltonian outputs use the preset sp structures, which must hold nxy elements each
  for (int k = 0; k < 4; k++)
    if (op.sp[k].nelem != op.nxy || op.hsp[k].nelem != op.nxy)
      throw std::runtime_error("symbolic: sp quadrant size differs from nxy");
  return op;
}

}  // namespace pnfam
