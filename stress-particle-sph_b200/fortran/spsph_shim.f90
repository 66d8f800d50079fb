!-------------------------------------------------------------------------------
! spsph_shim.f90 -- ISO_C_BINDING layer between the reference's Fortran driver and libspsph_cuda.so.
!
! Drop-in use (see INTEGRATION.md): compile this file instead of 2_SPH_main_2018.f90's Init_SPH /
! time_integration bodies; PROGRAM SPH_2018 (1_SPH_2018.f90) and 3_SPH_material_2018.f90 (reader, OutputMesh,
! OutputRes) stay unchanged. Module SPH_main_2018 keeps its public names:
!     Init_SPH          -> problem_input_data (host, unchanged) + spsph_create + spsph_upload
!     time_integration  -> spsph_step(itimestep_sph, time_sph, dt_sph)  [+ download on output steps]
!     Out_print_sph     -> unchanged (reads the module arrays refreshed by the download)
! NOT COMPILED IN THIS REPOSITORY'S CI: no Fortran compiler exists in the build image. The interfaces below
! mirror include/spsph.h field by field; tests/test_oracle_cpu.py::test_fortran_shim_matches_header checks
! that every struct member and entry point declared in the header appears here in the same order.
!-------------------------------------------------------------------------------
module spsph_c_api
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int), parameter :: SPSPH_MAX_TCURVES = 8, SPSPH_MAX_TCURVE_PTS = 128, SPSPH_MAX_BCS = 16
  integer(c_int), parameter :: SPSPH_NPROP = 20, SPSPH_NINT_VARS = 10

  type, bind(C) :: spsph_params
     integer(c_int32_t) :: struct_bytes, variant
     integer(c_int32_t) :: ndimn, nstre
     integer(c_int32_t) :: nnode, nstress, ntotal, ntotal2, ndummy
     integer(c_int32_t) :: ndummy2
     integer(c_int32_t) :: npoints
     integer(c_int32_t) :: sp_sph, inside_approach, sph_shift, vel_vector
     integer(c_int32_t) :: shift_update, dummy_nodes
     integer(c_int32_t) :: skf, sle, cspm, update_x, xsph
     integer(c_int32_t) :: cont_density, art_stress
     integer(c_int32_t) :: ntype_eco, ncrit, ntype_solid
     integer(c_int32_t) :: no_bcs, ifsigman, ic_grav, tcurve_grav
     integer(c_int32_t) :: bc_loop_ntotal
     real(c_float)      :: ae_threshold
     integer(c_int32_t) :: ntcurves
     integer(c_int32_t) :: nptstcurves(SPSPH_MAX_TCURVES)
     real(c_double) :: dx, dy, sml, r_x, r_y, disp_tol
     real(c_double) :: alpha, beta, damping
     real(c_double) :: ft_grav, cgrav(2)
     real(c_double) :: props(SPSPH_NPROP)
     real(c_double) :: D11, D22, D12, D33, D41, D42
     real(c_double) :: xmin_domain(2), xmax_domain(2)
     real(c_double) :: pi
     ! C row-major [curve][point] == Fortran (point, curve)
     real(c_double) :: ttcurves(SPSPH_MAX_TCURVE_PTS, SPSPH_MAX_TCURVES)
     real(c_float)  :: ftcurves(SPSPH_MAX_TCURVE_PTS, SPSPH_MAX_TCURVES)
     real(c_double) :: bc_list(8, SPSPH_MAX_BCS)
  end type spsph_params

  type, bind(C) :: spsph_state
     type(c_ptr) :: x, vel, stress, rho, mass, hsml, itype, internal_vars, f_drucker, x00, displ, x_10, disp_10
     type(c_ptr) :: wall_position, horizontal_or_not, n_int, bc_int, if_out_domain, bc_or_not, bc_info
  end type spsph_state

  interface
     integer(c_int) function spsph_create(h, p, device) bind(C, name="spsph_create")
       import :: c_ptr, c_int, spsph_params
       type(c_ptr), intent(out) :: h
       type(spsph_params), intent(in) :: p
       integer(c_int), value :: device
     end function
     integer(c_int) function spsph_upload(h, s) bind(C, name="spsph_upload")
       import :: c_ptr, c_int, spsph_state
       type(c_ptr), value :: h
       type(spsph_state), intent(in) :: s
     end function
     integer(c_int) function spsph_step(h, itimestep_sph, time_sph, dt_sph) bind(C, name="spsph_step")
       import :: c_ptr, c_int, c_int32_t, c_double
       type(c_ptr), value :: h
       integer(c_int32_t), value :: itimestep_sph
       real(c_double), value :: time_sph, dt_sph
     end function
     integer(c_int) function spsph_run(h, first_itimestep, time_sph, dt_sph, nsteps, time_sph_out) bind(C, name="spsph_run")
       import :: c_ptr, c_int, c_int32_t, c_double
       type(c_ptr), value :: h
       integer(c_int32_t), value :: first_itimestep, nsteps
       real(c_double), value :: time_sph, dt_sph
       real(c_double), intent(out) :: time_sph_out
     end function
     integer(c_int) function spsph_download(h, s) bind(C, name="spsph_download")
       import :: c_ptr, c_int, spsph_state
       type(c_ptr), value :: h
       type(spsph_state), intent(in) :: s
     end function
     integer(c_int) function spsph_pair_stats(h, npairs, maxiac, miniac, noiac) bind(C, name="spsph_pair_stats")
       import :: c_ptr, c_int, c_int32_t, c_int64_t
       type(c_ptr), value :: h
       integer(c_int64_t), intent(out) :: npairs
       integer(c_int32_t), intent(out) :: maxiac, miniac, noiac
     end function
     integer(c_int) function spsph_pairs(h, npairs, pair_i, pair_j, pint_type, w, dwdx, dwdy) bind(C, name="spsph_pairs")
       import :: c_ptr, c_int, c_int64_t
       type(c_ptr), value :: h
       integer(c_int64_t), intent(inout) :: npairs
       type(c_ptr), value :: pair_i, pair_j, pint_type, w, dwdx, dwdy
     end function
     integer(c_int) function spsph_last_run_ms(h, ms, kernel_launches) bind(C, name="spsph_last_run_ms")
       import :: c_ptr, c_int, c_float, c_int64_t
       type(c_ptr), value :: h
       real(c_float), intent(out) :: ms
       integer(c_int64_t), intent(out) :: kernel_launches
     end function
     integer(c_int) function spsph_profile(h, enable) bind(C, name="spsph_profile")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: enable
     end function
     integer(c_int) function spsph_profile_get(h, kid, name, total_ms, launches) bind(C, name="spsph_profile_get")
       import :: c_ptr, c_int, c_double, c_int64_t
       type(c_ptr), value :: h
       integer(c_int), value :: kid
       type(c_ptr), intent(out) :: name
       real(c_double), intent(out) :: total_ms
       integer(c_int64_t), intent(out) :: launches
     end function
     integer(c_int) function spsph_dist_unique_id(id128) bind(C, name="spsph_dist_unique_id")
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id128(128)
     end function
     integer(c_int) function spsph_dist_init(h, rank, nranks, id128, planes, halo_cells, halo_capacity) &
          bind(C, name="spsph_dist_init")
       import :: c_ptr, c_int, c_int32_t, c_char, c_double
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, nranks, halo_cells, halo_capacity
       character(kind=c_char), intent(in) :: id128(128)
       real(c_double), intent(in) :: planes(*)
     end function
     integer(c_int) function spsph_dist_flags(h, flags) bind(C, name="spsph_dist_flags")
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int32_t), intent(out) :: flags(*)
     end function
     integer(c_int) function spsph_local_counts(h, nloc3) bind(C, name="spsph_local_counts")
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int32_t), intent(out) :: nloc3(3)
     end function
     integer(c_int) function spsph_get_list_capacity(h, m_pairs) bind(C, name="spsph_get_list_capacity")
       import :: c_ptr, c_int, c_int64_t
       type(c_ptr), value :: h
       integer(c_int64_t), intent(out) :: m_pairs
     end function
     integer(c_int) function spsph_dist_set_planes(h, planes) bind(C, name="spsph_dist_set_planes")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: planes(*)
     end function
     integer(c_int) function spsph_upload_rows(h, s, ids, n) bind(C, name="spsph_upload_rows")
       import :: c_ptr, c_int, c_int32_t, spsph_state
       type(c_ptr), value :: h
       type(spsph_state), intent(in) :: s
       integer(c_int32_t), intent(in) :: ids(*)
       integer(c_int32_t), value :: n
     end function
     integer(c_int) function spsph_download_rows(h, s, ids, n) bind(C, name="spsph_download_rows")
       import :: c_ptr, c_int, c_int32_t, spsph_state
       type(c_ptr), value :: h
       type(spsph_state), intent(in) :: s
       integer(c_int32_t), intent(in) :: ids(*)
       integer(c_int32_t), value :: n
     end function
     ! output frame packed on the device: out(ncols, count) in Fortran order == the C side's (count, ncols) rows;
     ! cols = SPSPH_COL_* codes (0 x, 1 y, 2 vx, 3 vy, 4-7 stress, 8 plastic strain, 9 disp_10, 10 rho, 11 hsml,
     ! 12-13 displ, 14 f_drucker, 15 bc_or_not), first = 0-based particle number
     integer(c_int) function spsph_download_frame(h, cols, ncols, first, count, out) bind(C, name="spsph_download_frame")
       import :: c_ptr, c_int, c_int32_t, c_double
       type(c_ptr), value :: h
       integer(c_int32_t), intent(in) :: cols(*)
       integer(c_int32_t), value :: ncols, first, count
       real(c_double), intent(out) :: out(*)
     end function
     integer(c_int) function spsph_path_counts(h, tile_steps, list_steps) bind(C, name="spsph_path_counts")
       import :: c_ptr, c_int, c_int64_t
       type(c_ptr), value :: h
       integer(c_int64_t), intent(out) :: tile_steps, list_steps
     end function
     integer(c_int) function spsph_set_list_capacity(h, m_pairs) bind(C, name="spsph_set_list_capacity")
       import :: c_ptr, c_int, c_int64_t
       type(c_ptr), value :: h
       integer(c_int64_t), value :: m_pairs
     end function
     integer(c_int) function spsph_sync(h) bind(C, name="spsph_sync")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function spsph_destroy(h) bind(C, name="spsph_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     type(c_ptr) function spsph_last_error(h) bind(C, name="spsph_last_error")
       import :: c_ptr
       type(c_ptr), value :: h
     end function
     type(c_ptr) function spsph_version() bind(C, name="spsph_version")
       import :: c_ptr
     end function
  end interface
end module spsph_c_api

!-------------------------------------------------------------------------------
! Replacement bodies for the two hot entry points of MODULE SPH_main_2018 (2_SPH_main_2018.f90:32-46,78-184).
!-------------------------------------------------------------------------------
module spsph_shim
  use, intrinsic :: iso_c_binding
  use spsph_c_api
  use variable_types
  use SPH_time_vars_2018
  use SPH_global_vars_2018
  use SPH_material_vars_2018
  implicit none
  type(c_ptr), save :: handle = c_null_ptr
  type(spsph_state), save :: st
  type(spsph_params), save :: prm
  integer(c_int32_t), save :: variant = 1   ! SPSPH_VARIANT_BUI; set to the copy being built (0 code, 2 vs, 3 sl)
contains

  subroutine spsph_check(rc, what)
    integer(c_int), intent(in) :: rc
    character(*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc /= 0) then
       call c_f_pointer(spsph_last_error(handle), msg, [512])
       n = 1
       do while (n < 512 .and. msg(n) /= c_null_char)
          n = n + 1
       end do
       write(*,*) what, ': ', msg(1:n-1)      ! the reference reports errors with write(*,*) + stop
       stop
    end if
  end subroutine

  ! called at the end of Init_sph, after problem_input_data and x00 = x0 (2_SPH_main_2018.f90:40-42)
  subroutine spsph_init_device()
    integer :: i, j
    prm%variant = variant
    prm%ndimn = ndimn; prm%nstre = nstre
    prm%nnode = nnode; prm%nstress = nstress; prm%ntotal = ntotal; prm%ntotal2 = ntotal2
    prm%ndummy = ntotal2 - ntotal; prm%ndummy2 = ndummy2; prm%npoints = npoints
    prm%sp_sph = merge(1, 0, SP_SPH); prm%inside_approach = merge(1, 0, inside_approach)
    prm%sph_shift = merge(1, 0, SPH_shift); prm%vel_vector = merge(1, 0, vel_vector)
    prm%shift_update = shift_update; prm%dummy_nodes = merge(1, 0, dummy_nodes)
    prm%skf = skf; prm%sle = sle; prm%cspm = merge(1, 0, CSPM); prm%update_x = merge(1, 0, update_x)
    prm%xsph = merge(1, 0, XSPH); prm%cont_density = merge(1, 0, cont_density); prm%art_stress = merge(1, 0, art_stress)
    prm%ntype_eco = int(props(1,1)); prm%ncrit = int(props(1,2)); prm%ntype_solid = ntype_solid
    prm%no_bcs = no_bcs; prm%ifsigman = ifsigman; prm%ic_grav = ic_grav; prm%tcurve_grav = tcurve_grav
    prm%bc_loop_ntotal = merge(1, 0, variant == 1 .or. variant == 2)
    prm%ae_threshold = merge(1e-07, 1e-03, variant == 1)
    prm%ntcurves = ntcurves
    prm%nptstcurves = 0
    prm%nptstcurves(1:ntcurves) = nptstcurves(1:ntcurves)
    prm%dx = dx; prm%dy = dy; prm%sml = sml; prm%r_x = r_x; prm%r_y = r_y; prm%disp_tol = disp_tol
    prm%alpha = alpha; prm%beta = beta; prm%damping = DampingTG
    prm%ft_grav = ft_grav; prm%cgrav = 0
    if (ic_grav == 1) prm%cgrav(1:ndimn) = cgrav(1:ndimn)
    prm%props = props(1, 1:SPSPH_NPROP)
    prm%D11 = D11; prm%D22 = D22; prm%D12 = D12; prm%D33 = D33; prm%D41 = D41; prm%D42 = D42
    prm%xmin_domain = Xmin_Domain(1:2); prm%xmax_domain = Xmax_Domain(1:2)
    prm%pi = pi
    prm%ttcurves = 0; prm%ftcurves = 0; prm%bc_list = 0
    do i = 1, ntcurves
       do j = 1, nptstcurves(i)
          prm%ttcurves(j, i) = ttcurves(i, j)
          prm%ftcurves(j, i) = ftcurves(i, j)
       end do
    end do
    if (no_bcs > 0) prm%bc_list(1:8, 1:no_bcs) = bc_list(1:8, 1:no_bcs)
    prm%struct_bytes = int(c_sizeof(prm), c_int32_t)

    call spsph_check(spsph_create(handle, prm, 0_c_int), 'spsph_create')
    st%x = c_loc(x); st%vel = c_loc(vel); st%stress = c_loc(stress); st%rho = c_loc(rho); st%mass = c_loc(mass)
    st%hsml = c_loc(hsml); st%itype = c_loc(itype); st%internal_vars = c_loc(Internal_Vars)
    st%f_drucker = c_loc(f_drucker); st%x00 = c_loc(x00); st%displ = c_loc(displ); st%x_10 = c_loc(x_10)
    st%disp_10 = c_loc(disp_10); st%n_int = c_loc(n_int); st%bc_int = c_loc(bc_int)
    st%if_out_domain = c_loc(If_Out_Domain); st%bc_or_not = c_loc(BC_or_not); st%bc_info = c_loc(bc_info)
    st%wall_position = c_null_ptr; st%horizontal_or_not = c_null_ptr
    if (dummy_nodes) then
       st%wall_position = c_loc(wall_position); st%horizontal_or_not = c_loc(horizontal_or_not)
    end if
    call spsph_check(spsph_upload(handle, st), 'spsph_upload')
  end subroutine

  ! body of time_integration: one device step; the module arrays are refreshed only when the driver is about
  ! to write them (OutputRes / Out_print_sph cadence, 1_SPH_2018.f90:182-190)
  subroutine spsph_time_integration()
    call spsph_check(spsph_step(handle, int(itimestep_sph, c_int32_t), time_sph, dt_sph), 'spsph_step')
    if (t_plot_reset >= time_plot .or. t_print_reset >= time_print .or. time + dt > time_end) then
       call spsph_check(spsph_download(handle, st), 'spsph_download')
    end if
  end subroutine

  subroutine spsph_finalize()
    integer(c_int) :: rc
    rc = spsph_destroy(handle)
  end subroutine
end module spsph_shim
