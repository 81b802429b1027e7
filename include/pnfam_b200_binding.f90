! pnfam_b200_binding.f90 -- ISO_C_BINDING view of include/pnfam_b200.h, section 2 (libpnfam_b200.so, CUDA sm_100a).
!
! The derived types below are field for field the C structs of the header, in the same order (tests/test_host_and_abi.py
! parses both files and compares names, kinds and the resulting byte offsets).  A Fortran host keeps its own set-up
! (setup_pnfam, setup_extfield, init_pnfam_solver -- exes/pnfam/pnfam_solver.f90:44-48, 226-460) and replaces the call
! of ifam (pnfam_solver.f90:75) by pnfam_b200_solve, or only the procedure pointer calc_hamiltonian
! (pnfam_setup.f90:56-71, bound at :380) by a wrapper around pnfam_b200_calc_hamiltonian.  All arrays are passed as they
! are stored by the reference (column-major, 1-based ir2c / ir2m): c_loc() of the module variables, no copies.
!
! No Fortran compiler exists in the image this project is built in, so this file is shipped as (structurally tested)
! source; build it with the host code:  gfortran -c pnfam_b200_binding.f90 ; link with -lpnfam_b200 -lcudart.
module pnfam_b200_binding
   use iso_c_binding
   implicit none

   ! include/pnfam_b200.h: pnfam_b200_model -- module variables of hfb_solution / pnfam_interaction / type_blockmatrix
   type, bind(C) :: pnfam_b200_model
      integer(c_int32_t) :: nb
      integer(c_int32_t) :: dqp
      integer(c_int32_t) :: nghl
      type(c_ptr) :: db                 ! type_blockmatrix::db(nb)
      type(c_ptr) :: num_spin_up        ! hfb_solution::num_spin_up(nb)
      type(c_ptr) :: wf                 ! hfb_solution tables (nghl, dqp), column-major
      type(c_ptr) :: wfdr
      type(c_ptr) :: wfdp
      type(c_ptr) :: wfdz
      type(c_ptr) :: wfd2_all
      type(c_ptr) :: wdcori             ! hfb_solution::wdcori(nghl)
      type(c_ptr) :: crho               ! pnfam_interaction::crho(nghl) ...
      type(c_ptr) :: cs
      type(c_ptr) :: cpair
      type(c_ptr) :: cspair
      real(c_double) :: cdrho
      real(c_double) :: ctau
      real(c_double) :: ctj0
      real(c_double) :: ctj1
      real(c_double) :: ctj2
      real(c_double) :: crdj
      real(c_double) :: cds
      real(c_double) :: ct
      real(c_double) :: cj
      real(c_double) :: cgs
      real(c_double) :: cf
      real(c_double) :: csdj
      type(c_ptr) :: Ep                 ! pnfam_setup::Ep(dqp), En(dqp)
      type(c_ptr) :: En
      type(c_ptr) :: Up                 ! Up%elem, Vp%elem, Un%elem, Vn%elem (pnfam_setup.f90:292-321)
      type(c_ptr) :: Vp
      type(c_ptr) :: Un
      type(c_ptr) :: Vn
      type(c_ptr) :: qp_fp              ! c_null_ptr unless ft_active .or. hfb_blo_active (pnfam_setup.f90:323-362)
      type(c_ptr) :: qp_fn
      integer(c_int32_t) :: ngh         ! 0: no separable description (the general-table kernels run); the six fields
      integer(c_int32_t) :: ngl         ! below are then ignored but MUST be present -- they are part of the struct
      integer(c_int32_t) :: sep_nzrows
      type(c_ptr) :: sep_zrow           ! (dqp)
      type(c_ptr) :: sep_z              ! (ngh, sep_nzrows, 3)  Fortran order of the C array [3][sep_nzrows][ngh]
      type(c_ptr) :: sep_r              ! (ngl, dqp, 4)
   end type

   ! include/pnfam_b200.h: pnfam_b200_operator -- f and the cross-term fields g(k) (pnfam_extfield.f90:37-108, 882-949)
   type, bind(C) :: pnfam_b200_operator
      integer(c_int32_t) :: beta_minus
      integer(c_int32_t) :: nxterms
      type(c_ptr) :: f_ir2c             ! f%mat%ir2c(nb)
      type(c_ptr) :: f_elem             ! f%mat%elem(nxy)
      type(c_ptr) :: g_elem             ! array of nxterms c_ptr: c_loc(g(k)%mat%elem)
   end type

   ! include/pnfam_b200.h: pnfam_b200_solver_params -- &solver of the namelist (pnfam_setup.f90:98-108)
   type, bind(C) :: pnfam_b200_solver_params
      integer(c_int32_t) :: max_iter
      integer(c_int32_t) :: broyden_history_size
      real(c_double) :: convergence_epsilon
      real(c_double) :: quench_residual_int
      real(c_double) :: energy_shift_prot
      real(c_double) :: energy_shift_neut
      integer(c_int32_t) :: batch_slots  ! 0 = automatic
      integer(c_int32_t) :: reserved     ! 0
   end type

   ! include/pnfam_b200.h: pnfam_b200_stats
   type, bind(C) :: pnfam_b200_stats
      real(c_double) :: seconds_total
      real(c_double) :: seconds_device
      integer(c_int64_t) :: iterations
      integer(c_int64_t) :: kernel_launches
      integer(c_int64_t) :: h2d_bytes
      integer(c_int64_t) :: d2h_bytes
      real(c_double) :: seconds_density
      real(c_double) :: seconds_projection
      integer(c_int64_t) :: launches_density
      integer(c_int64_t) :: launches_projection
      real(c_double) :: flops_density
      real(c_double) :: flops_projection
      integer(c_int32_t) :: batch_slots
      integer(c_int32_t) :: lock_steps
   end type

   ! include/pnfam_b200.h: pnfam_b200_blockmatrix -- one type(blockmatrix) (pnfam_type_blockmatrix.f90:15-26)
   type, bind(C) :: pnfam_b200_blockmatrix
      type(c_ptr) :: elem               ! c_loc(bm%elem)
      type(c_ptr) :: ir2c               ! c_loc(bm%ir2c), 1-based, 0 = none
      type(c_ptr) :: ir2m               ! c_loc(bm%ir2m), 1-based
      integer(c_int64_t) :: nelem       ! size(bm%elem)
   end type

   interface
      integer(c_int) function pnfam_b200_ctx_create(model, device, ctx, err, errlen) bind(C, name="pnfam_b200_ctx_create")
         import
         type(pnfam_b200_model), intent(in) :: model
         integer(c_int), value :: device
         type(c_ptr), intent(out) :: ctx
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function

      subroutine pnfam_b200_ctx_destroy(ctx) bind(C, name="pnfam_b200_ctx_destroy")
         import
         type(c_ptr), value :: ctx
      end subroutine

      integer(c_int) function pnfam_b200_ctx_separable(ctx) bind(C, name="pnfam_b200_ctx_separable")
         import
         type(c_ptr), value :: ctx
      end function

      integer(c_int64_t) function pnfam_b200_ctx_h2d_bytes(ctx) bind(C, name="pnfam_b200_ctx_h2d_bytes")
         import
         type(c_ptr), value :: ctx
      end function

      ! replaces ifam (pnfam_solver.f90:93-221) for npoints frequencies of one operator
      integer(c_int) function pnfam_b200_solve(ctx, op, prm, npoints, omega_re, omega_im, strength, iters, conv, si, &
                                               trace, stats, err, errlen) bind(C, name="pnfam_b200_solve")
         import
         type(c_ptr), value :: ctx
         type(pnfam_b200_operator), intent(in) :: op
         type(pnfam_b200_solver_params), intent(in) :: prm
         integer(c_int32_t), value :: npoints
         real(c_double), intent(in) :: omega_re(*), omega_im(*)
         real(c_double), intent(out) :: strength(*)          ! (2, 1+nxterms, npoints)
         integer(c_int32_t), intent(out) :: iters(*), conv(*)
         real(c_double), intent(out) :: si(*)
         type(c_ptr), value :: trace                         ! c_null_ptr or c_loc of (4, max_iter+1, npoints)
         type(c_ptr), value :: stats                         ! c_null_ptr or c_loc of a pnfam_b200_stats
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function

      ! the plug-in point calc_hamiltonian (pnfam_setup.f90:56-71): 8 input and 8 output block matrices
      integer(c_int) function pnfam_b200_calc_hamiltonian(ctx, bm_in, bm_out, err, errlen) &
            bind(C, name="pnfam_b200_calc_hamiltonian")
         import
         type(c_ptr), value :: ctx
         type(pnfam_b200_blockmatrix), intent(in) :: bm_in(8)
         type(pnfam_b200_blockmatrix), intent(inout) :: bm_out(8)
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function

      integer(c_int) function pnfam_b200_dmma_peak(device, tflops, err, errlen) bind(C, name="pnfam_b200_dmma_peak")
         import
         integer(c_int), value :: device
         real(c_double), intent(out) :: tflops
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function

      ! host-only self-check of the transform plan (no device needed)
      integer(c_int) function pnfam_b200_check_transform_plan(nb, db, f_ir2c, use_diag, beta_minus, res, err, errlen) &
            bind(C, name="pnfam_b200_check_transform_plan")
         import
         integer(c_int32_t), value :: nb, use_diag, beta_minus
         integer(c_int32_t), intent(in) :: db(*), f_ir2c(*)
         real(c_double), intent(out) :: res(8)
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function

      ! libpnfam_host.so: the Yukawa part of the full-FAM two-body-current GT field, replaces the call
      !    call effective_2bc_extfield(f)            (pnfam_solver.f90:629, pnfam_extfield_2bc.f90:26-465)
      ! with the arrays that routine takes from hfb_solution / hfbtho / type_blockmatrix:
      !    ierr = pnfam_host_effective_2bc_extfield(ntx, nbx, hdb_half, hnz, hnr, hnl, hns, bz, bp, rmat, &
      !              int(size(rmat,1),c_int64_t), f%mat%ir2c, f%mat%ir2m, int(size(f%mat%elem),c_int64_t), f%k, &
      !              merge(1,0,f%beta_minus), merge(1,0,two_body_current_usep), 1, &
      !              f%gam%c3d, f%gam%c3e, f%gam%c4d, f%gam%c4e, f%gam%cpd, f%gam%cpe, err, len(err))
      ! (hdb_half = id(1:nbx), the block sizes of the undoubled HFBTHO basis); then mult_gamdel_lecs + calc_totgamdel as
      ! read_tbc does (pnfam_storage.f90:654-661).
      integer(c_int) function pnfam_host_effective_2bc_extfield(ntx, nbx, id, hnz, hnr, hnl, hns, bz, bp, rmat, ld_rmat, &
            ir2c, ir2m, nxy, k, beta_minus, use_p, spin_sorted, c3d, c3e, c4d, c4e, cpd, cpe, err, errlen) &
            bind(C, name="pnfam_host_effective_2bc_extfield")
         import
         integer(c_int32_t), value :: ntx, nbx, k, beta_minus, use_p, spin_sorted
         integer(c_int32_t), intent(in) :: id(*), hnz(*), hnr(*), hnl(*), hns(*), ir2c(*), ir2m(*)
         real(c_double), value :: bz, bp
         real(c_double), intent(in) :: rmat(*)
         integer(c_int64_t), value :: ld_rmat, nxy
         real(c_double), intent(out) :: c3d(*), c3e(*), c4d(*), c4e(*), cpd(*), cpe(*)
         character(kind=c_char) :: err(*)
         integer(c_int), value :: errlen
      end function
   end interface

   ! calc_hamiltonian-shaped wrapper (to be placed in the host's own module, after `contains`): point `calc_hamiltonian => calc_dHsp_b200` in setup_hamiltonian
   ! (pnfam_setup.f90:371-382).  b200_ctx is created once after init_pnfam_solver.
   ! (type(blockmatrix) comes from the host's own module type_blockmatrix.)
   !
   !   subroutine calc_dHsp_b200(rerho_pn, imrho_pn, rekp, imkp, rerho_np, imrho_np, rekm, imkm, &
   !                             reh_pn, imh_pn, redp, imdp, reh_np, imh_np, redm, imdm)
   !      use type_blockmatrix
   !      type(blockmatrix), intent(inout), target :: rerho_pn, imrho_pn, rekp, imkp, rerho_np, imrho_np, rekm, imkm
   !      type(blockmatrix), intent(inout), target :: reh_pn, imh_pn, redp, imdp, reh_np, imh_np, redm, imdm
   !      type(pnfam_b200_blockmatrix) :: bi(8), bo(8)
   !      character(kind=c_char) :: err(512)
   !      bi(1) = view(rerho_pn); bi(2) = view(imrho_pn); bi(3) = view(rekp); bi(4) = view(imkp)
   !      bi(5) = view(rerho_np); bi(6) = view(imrho_np); bi(7) = view(rekm); bi(8) = view(imkm)
   !      bo(1) = view(reh_pn);  bo(2) = view(imh_pn);  bo(3) = view(redp);  bo(4) = view(imdp)
   !      bo(5) = view(reh_np);  bo(6) = view(imh_np);  bo(7) = view(redm);  bo(8) = view(imdm)
   !      if (pnfam_b200_calc_hamiltonian(b200_ctx, bi, bo, err, 512) /= 0) call abort(err)
   !   contains
   !      function view(bm) result(v)
   !         type(blockmatrix), intent(in), target :: bm
   !         type(pnfam_b200_blockmatrix) :: v
   !         v%elem = c_loc(bm%elem); v%ir2c = c_loc(bm%ir2c); v%ir2m = c_loc(bm%ir2m); v%nelem = size(bm%elem, kind=c_int64_t)
   !      end function
   !   end subroutine

end module pnfam_b200_binding
