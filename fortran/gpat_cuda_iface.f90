!*******************************************************************************
!< ISO_C_BINDING view of include/gpat_cuda.h (libgpat_cuda.so), for GPAT's driver
!< src/programs/stochastic-mhd.f90.  Each interface names the reference procedure
!< it replaces.  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no
!< Fortran compiler; layout agreement with the C header is checked from the C side
!< (tests/test_cpu_host.py::test_struct_layouts_match_the_c_header) and this file
!< mirrors the header field by field.
!*******************************************************************************
module gpat_cuda
    use, intrinsic :: iso_c_binding
    implicit none
    private
    public :: gpat_params, gpat_hist_spec, gpat_particle, gpat_counters, gpat_timings
    public :: gpat_init, gpat_set_params, gpat_finalize, gpat_last_error
    public :: gpat_upload_fields, gpat_upload_turbulence, gpat_upload_acc_surface, gpat_prefetch_fields, gpat_swap_fields
    public :: gpat_inject_uniform, gpat_inject_targeted, gpat_inject_at_shock, gpat_particle_mover, gpat_split
    public :: gpat_init_tracking, gpat_tracked_shape, gpat_download_tracked, gpat_reset_tracked
    public :: gpat_download_particles, gpat_upload_particles
    public :: gpat_download_escaped, gpat_reset_escaped
    public :: gpat_get_counters, gpat_set_counters
    public :: gpat_diagnostics, gpat_escaped_diagnostics, gpat_escaped_local_diagnostics, gpat_hist_edges
    public :: gpat_comm_unique_id, gpat_comm_init, gpat_comm_destroy
    public :: gpat_get_timings, gpat_check

    !< struct gpat_hist_spec: one set of local-distribution parameters
    !< (diagnostics.f90:53-65, 2118-2140)
    type, bind(C) :: gpat_hist_spec
        integer(c_int32_t) :: enabled, npbins, nmu, rx, ry, rz
        real(c_double) :: pmin, pmax
    end type gpat_hist_spec

    !< struct gpat_params: what read_particle_params, set_*_params,
    !< read_diagnostics_params, mhd_config and fconfig keep in module variables
    type, bind(C) :: gpat_params
        integer(c_int32_t) :: ndim, nx, ny, nz, time_interp
        integer(c_int32_t) :: pbc(3)
        real(c_double) :: dx, dy, dz, xmin, ymin, zmin, xmax, ymax, zmax, lx, ly, lz
        real(c_double) :: b0, p0, pmin, pmax, gamma_turb, pindex, kpara0, kret
        real(c_double) :: dt_min_rel, dt_max_rel
        integer(c_int32_t) :: momentum_dependency, mag_dependency, acc_region_flag, pad0
        real(c_double) :: acc_region(6)
        integer(c_int32_t) :: dpp_wave, dpp_shear, weak_scattering, keep_rho
        real(c_double) :: tau0, drift1, drift2
        integer(c_int32_t) :: pcharge, check_drift_2d, include_3rd_dim, nlgc
        real(c_double) :: kperp_kpara, duu0
        integer(c_int32_t) :: focused_transport, spherical_coord, nonuniform_grid
        integer(c_int32_t) :: deltab_flag, correlation_flag, acc_by_surface
        integer(c_int32_t) :: npp_global, nmu_global
        type(gpat_hist_spec) :: local(4)
        integer(c_int64_t) :: seed
        integer(c_int32_t) :: rng_mode, mpi_rank, strict_math
        integer(c_int32_t) :: surface_norm1, surface_norm2, surface2_existed, is_intersection, pad2
    end type gpat_params

    !< struct gpat_particle == particle_type (particle_module.f90:38-50), 104 bytes
    type, bind(C) :: gpat_particle
        integer(c_int8_t)  :: split_times, count_flag, pad_(2)
        integer(c_int32_t) :: origin, nsteps_tracked, nsteps_pushed, tag_injected, tag_splitted
        real(c_double) :: x, y, z, p, v, mu, weight, t, dt, padding
    end type gpat_particle

    type, bind(C) :: gpat_counters
        integer(c_int64_t) :: nptl_current, nptl_split, nptl_escaped, nptl_max, tag_max
        real(c_double) :: leak, leak_negp
    end type gpat_counters

    type, bind(C) :: gpat_timings
        real(c_float) :: mover_ms, push_ms, compact_ms, upload_ms, grad_ms, inject_ms, split_ms, diag_ms
        integer(c_int64_t) :: push_steps
        integer(c_int32_t) :: push_launches, total_launches
    end type gpat_timings

    interface
        !< init_particles (particle_module.f90:171) + init_prng (random_number_generator.f90:28)
        !< + init_field_data (mhd_data_parallel.f90:66) + init_particle_distributions (diagnostics.f90:178)
        integer(c_int) function gpat_init(h, device, nptl_max, params) bind(C, name="gpat_init")
            import :: c_ptr, c_int, c_int64_t, gpat_params
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: device
            integer(c_int64_t), value :: nptl_max
            type(gpat_params), intent(in) :: params
        end function gpat_init

        integer(c_int) function gpat_set_params(h, params) bind(C, name="gpat_set_params")
            import :: c_ptr, c_int, gpat_params
            type(c_ptr), value :: h
            type(gpat_params), intent(in) :: params
        end function gpat_set_params

        !< free_particles / delete_prng / free_field_data
        integer(c_int) function gpat_finalize(h) bind(C, name="gpat_finalize")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
        end function gpat_finalize

        type(c_ptr) function gpat_last_error(h) bind(C, name="gpat_last_error")
            import :: c_ptr
            type(c_ptr), value :: h
        end function gpat_last_error

        !< consumer side of read_field_data_parallel (mhd_data_parallel.f90:224) +
        !< calc_fields_gradients (mhd_data_parallel.f90:504); f = c_loc(farrayK(1,-1,-1,-1))
        integer(c_int) function gpat_upload_fields(h, slot, f, nvar, with_grad) &
                bind(C, name="gpat_upload_fields")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, f
            integer(c_int), value :: slot, nvar, with_grad
        end function gpat_upload_fields

        !< read_magnetic_fluctuation / read_correlation_length + their gradient passes
        !< (mhd_data_parallel.f90:306-497, 771-1604); data = slab array then 2-D array
        integer(c_int) function gpat_upload_turbulence(h, which, slot, data) bind(C, name="gpat_upload_turbulence")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, data
            integer(c_int), value :: which, slot
        end function gpat_upload_turbulence

        !< read_acc_surface (acc_region_surface.f90:121); heights = c_loc(acc_surfaceK1 or K2)
        integer(c_int) function gpat_upload_acc_surface(h, which, slot, heights) &
                bind(C, name="gpat_upload_acc_surface")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, heights
            integer(c_int), value :: which, slot
        end function gpat_upload_acc_surface

        !< frame pipeline: start the H2D copy of a frame already read into farray-shaped host
        !< memory (e.g. frame tf+1 during the push of frame tf); a later gpat_upload_fields with
        !< the same pointer only runs the gradient/pack kernel (stochastic-mhd.f90:401-447)
        integer(c_int) function gpat_prefetch_fields(h, f, nvar) bind(C, name="gpat_prefetch_fields")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, f
            integer(c_int), value :: nvar
        end function gpat_prefetch_fields

        !< copy_fields (mhd_data_parallel.f90:1920)
        integer(c_int) function gpat_swap_fields(h) bind(C, name="gpat_swap_fields")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
        end function gpat_swap_fields

        !< inject_particles_spatial_uniform (particle_module.f90:454)
        integer(c_int) function gpat_inject_uniform(h, nptl, dt, dist_flag, particle_v0, t_frame, &
                dt_mhd, part_box, power_index) bind(C, name="gpat_inject_uniform")
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: h
            integer(c_int64_t), value :: nptl
            real(c_double), value :: dt, particle_v0, t_frame, dt_mhd, power_index
            integer(c_int), value :: dist_flag
            real(c_double), intent(in) :: part_box(6)
        end function gpat_inject_uniform

        !< inject_particles_at_large_jz / _absj / _divv / _rho (particle_module.f90:785-1468) with
        !< get_ncells_large_* (mhd_data_parallel.f90:2211-2498); mode = 1 jz, 2 absj, 4 divv, 5 rho
        integer(c_int) function gpat_inject_targeted(h, mode, nptl, dt, dist_flag, particle_v0, &
                t_frame, dt_mhd, part_box, power_index, inject_same_nptl, vmin, ncells_norm, &
                nptl_injected, ncells) bind(C, name="gpat_inject_targeted")
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: mode, dist_flag, inject_same_nptl
            integer(c_int64_t), value :: nptl, ncells_norm
            real(c_double), value :: dt, particle_v0, t_frame, dt_mhd, power_index, vmin
            real(c_double), intent(in) :: part_box(6)
            integer(c_int64_t), intent(out) :: nptl_injected, ncells
        end function gpat_inject_targeted

        !< locate_shock_xpos (mhd_data_parallel.f90:1988) + inject_particles_at_shock (particle_module.f90:542)
        integer(c_int) function gpat_inject_at_shock(h, nptl, dt, dist_flag, particle_v0, t_frame, &
                power_index) bind(C, name="gpat_inject_at_shock")
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: h
            integer(c_int64_t), value :: nptl
            integer(c_int), value :: dist_flag
            real(c_double), value :: dt, particle_v0, t_frame, power_index
        end function gpat_inject_at_shock

        !< init_particle_tracking (particle_module.f90:5825) without the HDF5 read:
        !< tags = c_loc(tags_tracking), ncols = split_times_max + 2
        integer(c_int) function gpat_init_tracking(h, tags, ncols, nptl_tracking, nsteps_interval) &
                bind(C, name="gpat_init_tracking")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: h, tags
            integer(c_int), value :: ncols, nsteps_interval
            integer(c_int64_t), value :: nptl_tracking
        end function gpat_init_tracking

        integer(c_int) function gpat_tracked_shape(h, nsteps_tracking_max, nptl_tracking) &
                bind(C, name="gpat_tracked_shape")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: h
            integer(c_int64_t), intent(out) :: nsteps_tracking_max, nptl_tracking
        end function gpat_tracked_shape

        !< particles_tracked(nsteps_tracking_max, nptl_tracking) for dump_tracked_particles
        !< (particle_module.f90:6236); out = c_loc(particles_tracked)
        integer(c_int) function gpat_download_tracked(h, out) bind(C, name="gpat_download_tracked")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, out
        end function gpat_download_tracked

        !< reset_tracked_particles (particle_module.f90:5884)
        integer(c_int) function gpat_reset_tracked(h) bind(C, name="gpat_reset_tracked")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
        end function gpat_reset_tracked

        !< particle_mover (particle_module.f90:1846), both remove_particles passes included
        integer(c_int) function gpat_particle_mover(h, t0, dtf, nsteps_interval, num_fine_steps, &
                dump_escaped_dist, steps_done) bind(C, name="gpat_particle_mover")
            import :: c_ptr, c_int, c_double, c_int64_t
            type(c_ptr), value :: h
            real(c_double), value :: t0, dtf
            integer(c_int), value :: nsteps_interval, num_fine_steps, dump_escaped_dist
            integer(c_int64_t), intent(out) :: steps_done
        end function gpat_particle_mover

        !< split_particle (particle_module.f90:5430)
        integer(c_int) function gpat_split(h, split_ratio, pmin_split, nsteps_interval) &
                bind(C, name="gpat_split")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: h
            real(c_double), value :: split_ratio, pmin_split
            integer(c_int), value :: nsteps_interval
        end function gpat_split

        !< ptls(1:n) for dump_particles (diagnostics.f90:1811) / read_particles (particle_module.f90:5744)
        integer(c_int) function gpat_download_particles(h, out, nmax, n) bind(C, name="gpat_download_particles")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: h, out
            integer(c_int64_t), value :: nmax
            integer(c_int64_t), intent(out) :: n
        end function gpat_download_particles

        integer(c_int) function gpat_upload_particles(h, in, n) bind(C, name="gpat_upload_particles")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: h, in
            integer(c_int64_t), value :: n
        end function gpat_upload_particles

        integer(c_int) function gpat_download_escaped(h, out, nmax, n) bind(C, name="gpat_download_escaped")
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: h, out
            integer(c_int64_t), value :: nmax
            integer(c_int64_t), intent(out) :: n
        end function gpat_download_escaped

        !< reset_escaped_particles (stochastic-mhd.f90:533)
        integer(c_int) function gpat_reset_escaped(h) bind(C, name="gpat_reset_escaped")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
        end function gpat_reset_escaped

        integer(c_int) function gpat_get_counters(h, c) bind(C, name="gpat_get_counters")
            import :: c_ptr, c_int, gpat_counters
            type(c_ptr), value :: h
            type(gpat_counters), intent(out) :: c
        end function gpat_get_counters

        integer(c_int) function gpat_set_counters(h, c) bind(C, name="gpat_set_counters")
            import :: c_ptr, c_int, gpat_counters
            type(c_ptr), value :: h
            type(gpat_counters), intent(in) :: c
        end function gpat_set_counters

        !< calc_particle_distributions (diagnostics.f90:738) + quick_check (diagnostics.f90:116)
        !< + get_pmax_global (diagnostics.f90:1691); outputs already reduced over ranks.
        !< flocal: array of 4 c_ptr (c_loc(flocalK) or c_null_ptr)
        integer(c_int) function gpat_diagnostics(h, local_dist, fglobal, flocal, quick, pmax) &
                bind(C, name="gpat_diagnostics")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: h, fglobal
            integer(c_int), value :: local_dist
            type(c_ptr), intent(in) :: flocal(4)
            real(c_double), intent(out) :: quick(8), pmax
        end function gpat_diagnostics

        !< calc_escaped_distributions (diagnostics.f90:913), global part
        integer(c_int) function gpat_escaped_diagnostics(h, fescaped) bind(C, name="gpat_escaped_diagnostics")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, fescaped
        end function gpat_escaped_diagnostics

        !< calc_escaped_distributions, local part (diagnostics.f90:956-1170): fx/fy/fz are arrays of four
        !< c_ptr (c_loc(fescapedK_x) ... or c_null_ptr), K = 1..4
        integer(c_int) function gpat_escaped_local_diagnostics(h, fx, fy, fz) &
                bind(C, name="gpat_escaped_local_diagnostics")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            type(c_ptr), intent(in) :: fx(4), fy(4), fz(4)
        end function gpat_escaped_local_diagnostics

        integer(c_int) function gpat_hist_edges(h, which, pedges, muedges) bind(C, name="gpat_hist_edges")
            import :: c_ptr, c_int
            type(c_ptr), value :: h, pedges, muedges
            integer(c_int), value :: which
        end function gpat_hist_edges

        integer(c_int) function gpat_comm_unique_id(id) bind(C, name="gpat_comm_unique_id")
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
        end function gpat_comm_unique_id

        integer(c_int) function gpat_comm_init(h, id, nranks, rank) bind(C, name="gpat_comm_init")
            import :: c_ptr, c_int, c_char
            type(c_ptr), value :: h
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int), value :: nranks, rank
        end function gpat_comm_init

        integer(c_int) function gpat_comm_destroy(h) bind(C, name="gpat_comm_destroy")
            import :: c_ptr, c_int
            type(c_ptr), value :: h
        end function gpat_comm_destroy

        integer(c_int) function gpat_get_timings(h, t) bind(C, name="gpat_get_timings")
            import :: c_ptr, c_int, gpat_timings
            type(c_ptr), value :: h
            type(gpat_timings), intent(out) :: t
        end function gpat_get_timings
    end interface

    contains

    !< The reference's error style: print on every rank that fails, MPI_FINALIZE, stop
    !< (simulation_setup.f90:78-87, diagnostics.f90:1989-2014).
    subroutine gpat_check(ierr, h, what)
        use mpi
        integer(c_int), intent(in) :: ierr
        type(c_ptr), intent(in) :: h
        character(*), intent(in) :: what
        character(kind=c_char), pointer :: msg(:)
        integer :: i, n, mpierr
        if (ierr == 0) return
        call c_f_pointer(gpat_last_error(h), msg, [512])
        n = 0
        do i = 1, 512
            if (msg(i) == c_null_char) exit
            n = i
        enddo
        write(*, "(A,A,A,I0,A,512A1)") "gpat_cuda: ", what, " failed (", ierr, "): ", msg(1:n)
        call MPI_FINALIZE(mpierr)
        stop
    end subroutine gpat_check
end module gpat_cuda
