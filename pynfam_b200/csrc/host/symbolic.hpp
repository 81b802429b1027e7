// Symbolic phase of the quasiparticle <-> single-particle transforms.
//
// The reference evaluates  dRsp = W_a . dRqp . W_b^T  and  dHqp = W_a^T . dHsp . W_b  through a
// generic 2x2 "bigblockmatrix" algebra with per-quadrant sign / transpose flags
// (exes/pnfam/pnfam_type_bbm.f90:428-576, pnfam_solver.f90:144-166,276-372) that is re-interpreted
// on every call.  Here that interpretation is done ONCE per operator on the host and flattened into
// a list of block tasks   out_block = sum_t alpha_t * op(A_t) op(B_t) op(C_t)   which the device
// executes as two grouped-GEMM launches per transform (csrc/cuda/transform.cu).
//
// Quadrant storage order used everywhere in the product (k = 0..3):
//   qp-basis objects (dRqp, dHqp, Fqp, Greens, T):  k=0 X/H20 (m12), k=1 Y/H02 (m21), k=2 P/H11 (m11), k=3 Q/H11t (m22)
//     -- the Broyden pack order of pnfam_broyden.f90:50-60
//   sp-basis objects (dRsp, dHsp):                  k=0 rho_pn/h_pn (m11), k=1 kappa+/Delta+ (m12),
//                                                   k=2 kappa-/Delta- (m21), k=3 rho_np/h_np (m22)
#pragma once
#include <cstdint>
#include <vector>


namespace pnfam {

struct BlockStruct {  // block maps of one stored quadrant; 0-based, -1 = no block
  std::vector<int> r2c, r2m;  // per block row: partner column block, element offset
  size_t nelem = 0;
  bool allocated = false;
};

struct TripleTerm {
  int a_mat, a_off, a_trans;   // A block: which W matrix (0 Ua,1 Va,2 Ub,3 Vb), element offset, transpose
  int b_quad, b_off, b_trans;  // B block: quadrant of the middle operand, element offset, transpose
  int c_mat, c_off, c_trans;   // C block
  double alpha_re, alpha_im;   // coefficient for the real / imaginary-part object
};

struct BlockTask {
  int out_quad, out_off, m, n;  // output block (m x n, column-major) in quadrant out_quad
  int nterms;
  TripleTerm t[4];
};

struct TransformPlan {
  std::vector<BlockTask> tasks;
  BlockStruct out[4];           // block structure of the four output quadrants
};

struct OperatorPlan {
  int nb = 0;
  size_t nxy = 0;
  bool beta_minus = true, use_diag = false;
  BlockStruct qp[4];            // X, Y, P, Q structures (from Fqp)
  BlockStruct sp[4];            // dRsp: rho_pn, kappa+, kappa-, rho_np (structures computed by the forward transform)
  BlockStruct hsp[4];           // dHsp: h_pn, Delta+, Delta-, h_np (structures preset, pnfam_solver.f90:402-413)
  TransformPlan forward;        // dRqp -> dRsp
  TransformPlan backward;       // dHsp -> dHqp   (also used once for F, G -> qp basis)
  // W matrices as the device sees them: [Ua | Va | Ub | Vb], a = row isospin (p for beta-), b = column
  std::vector<int> u_off, v_off;  // per block row offsets into U / V element arrays
};

// db: block dimensions (nb); f_ir2c: 1-based partner column block of the external field (0 = none);
// use_diag: P,Q quadrants active (odd-A equal filling / finite temperature).
OperatorPlan make_operator_plan(const std::vector<int>& db, const std::vector<int>& f_ir2c, bool use_diag, bool beta_minus);

}  // namespace pnfam
