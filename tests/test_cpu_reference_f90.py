"""The oracle pinned to the reference's OWN Fortran.

tests/golden/ref_f90/<case>.npz were computed by executing the unmodified procedures of
/root/reference/src/modules/*.f90 (inject_particles_spatial_uniform, inject_one_particle, particle_mover,
particle_mover_one_cycle, particle_boundary_condition, get_interp_paramters, interp_fields,
calc_fields_gradients, calc_spatial_diffusion_coefficients[_nlgc], calc_dpp_*, every push_particle_*,
remove_particles, split_particle, calc_particle_distributions, calc_escaped_distributions, quick_check,
get_pmax_global, init_particle_distributions ...) with oracle/f90/f90run.py -- there is no Fortran compiler in
this image or on the B200 box (profiles/r02a_fortran_probe.log).  Three layers:

  1. oracle/gpat_oracle.c == golden, BIT FOR BIT, for every case (always runs; no /root/reference needed);
  2. the golden files are what the reference computes: regenerated live from /root/reference for a subset of
     the cases (skipped where the reference tree is absent, e.g. on the GPU box);
  3. the translator obeys Fortran's rules: literal kinds, mixed-kind promotion, integer division, real**integer,
     assignment conversion, derived-type copies, array sections and bounds, intent(out) copy-back, DO semantics --
     checked on small Fortran sources written for the purpose.
"""
import os
import sys
import textwrap

import numpy as np
import pytest

from helpers import REF_GOLDEN_CASES, assert_particles_identical, golden_collect, unpack_sparse
from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "f90"))
import f90run as F  # noqa: E402
import refsim  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ref_f90")
LIVE_CASES = ["c1_2d", "c3_shock_open", "c4_dpp_wave_shear", "c5_3d", "s1_shock_1d"]


def assert_same_as_golden(got, z, what):
    """bit-for-bit: float arrays are compared through their bit patterns"""
    assert sorted(got) == sorted(z.files), (what, set(got) ^ set(z.files))
    for k in z.files:
        a, b = np.asarray(got[k]), z[k]
        if a.dtype.names:
            assert_particles_identical(a, b, f"{what}:{k}")
            continue
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        if a.dtype.kind == "f":
            a, b = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64), np.ascontiguousarray(b, dtype=np.float64).view(np.uint64)
        bad = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
        assert len(bad) == 0, f"{what}: {k} differs at {len(bad)} of {a.size} elements, first {bad[:4]}"


@pytest.mark.parametrize("name", REF_GOLDEN_CASES)
def test_oracle_equals_the_reference_golden_bit_for_bit(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    got = golden_collect(lambda P, n: Oracle(P, n), name)
    assert_same_as_golden(got, z, name)
    # the run must have exercised what it claims to pin
    assert int(z["run_steps"]) > 1000 and int(z["steps41_count"]) > 40 * 40


def test_golden_runs_cover_escapes_splits_and_every_histogram():
    tot_escaped = tot_split = 0
    for name in REF_GOLDEN_CASES:
        z = np.load(os.path.join(GOLD, name + ".npz"))
        tot_split += int(z["run_counters_int"][1])
        tot_escaped += sum(len(z[k]) for k in z.files if k.endswith("escaped_particles"))
        assert any(k.endswith("flocal1.val") and len(z[k]) for k in z.files), name
    assert tot_split >= 50 and tot_escaped >= 40
    z = np.load(os.path.join(GOLD, "c3_shock_open.npz"))
    assert unpack_sparse(z, "run_f2_fescaped1_x").sum() >= 4 and z["run_f2_fescaped"].sum() >= 4


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("name", LIVE_CASES)
def test_golden_is_what_the_reference_fortran_computes(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    got = golden_collect(lambda P, n: refsim.RefSim(P, n), name)
    assert_same_as_golden(got, z, name)


# ---------------------------------------------------------------------------------------------------
# the translator itself
# ---------------------------------------------------------------------------------------------------
def run_f90(tmp_path, source, proc, *args, module="m"):
    p = tmp_path / "m.f90"
    p.write_text(textwrap.dedent(source))
    prog = F.Program()
    prog.ext_values.update(real32=4, real64=8)
    prog.load_file(str(p))
    prog.init_module_data()
    return prog, prog.get(module, proc)(*args)


HEAD = """
module m
    implicit none
    integer, parameter :: sp = kind(1.0), dp = kind(1.0d0)
    type pt
        integer :: n
        real(dp) :: x
    end type pt
    real(dp) :: acc = 0.0_dp
    type(pt), allocatable, dimension(:) :: arr
    real(sp), allocatable, dimension(:, :) :: g
contains
"""


def test_f90run_literal_kinds_and_promotion(tmp_path):
    src = HEAD + """
    subroutine s(p, a, b, c, d, e)
        real(dp), intent(in) :: p
        real(dp), intent(out) :: a, b, c, d, e
        integer :: q
        q = 3
        a = 0.1 * p                 ! default-real literal: single precision 0.1, promoted
        b = 0.1_dp * p
        c = 1.0 / (3 * q) / p       ! real(sp) / integer stays single precision, then promotes
        d = 2 / 3 + 7 / 2 + (-7) / 2  ! integer division truncates toward zero
        e = 0.5**(1.0 + q)          ! real(sp) ** real(sp)
    end subroutine s
    end module m
    """
    _, (a, b, c, d, e) = run_f90(tmp_path, src, "s", np.float64(3.0), None, None, None, None, None)
    assert a == np.float64(np.float32(0.1)) * 3.0 and a != 0.1 * 3.0
    assert b == np.float64(0.1) * 3.0
    assert c == np.float64(np.float32(1.0) / np.float32(9.0)) / 3.0
    assert d == 0.0 and isinstance(d, np.float64)
    assert e == 0.0625


def test_f90run_integer_powers_follow_libgcc_powi(tmp_path):
    src = HEAD + """
    subroutine s(x, n, a, b, c)
        real(dp), intent(in) :: x
        integer, intent(in) :: n
        real(dp), intent(out) :: a, b, c
        a = x**n
        b = x**(-2)
        c = x**2.0_dp
    end subroutine s
    end module m
    """
    x = np.float64(1.0000001234567)
    _, (a, b, c) = run_f90(tmp_path, src, "s", x, 5, None, None, None)
    x2 = x * x
    assert a == x * (x2 * x2)            # square-and-multiply: r = x; a = x^2; a = x^4; r = r * a
    assert b == 1.0 / (x * x)
    import math
    assert c == math.pow(x, 2.0)


def test_f90run_assignment_converts_and_structs_copy(tmp_path):
    src = HEAD + """
    subroutine s(r, i, y)
        real(dp), intent(in) :: r
        integer, intent(out) :: i
        real(dp), intent(out) :: y
        type(pt) :: a, b
        real(sp) :: h
        i = r                ! truncation toward zero
        h = r                ! rounding to single
        a%n = 1
        a%x = h
        b = a                ! a copy, not an alias
        b%x = 2.0
        allocate(arr(3))
        arr(2) = a
        a%n = 7
        y = arr(2)%x + b%x + arr(2)%n
    end subroutine s
    end module m
    """
    _, (i, y) = run_f90(tmp_path, src, "s", np.float64(-2.7), None, None)
    assert i == -2 and isinstance(i, int)
    assert y == np.float64(np.float32(-2.7)) + 2.0 + 1


def test_f90run_sections_bounds_and_mixed_kind_array_arithmetic(tmp_path):
    src = HEAD + """
    subroutine s(scale, out)
        real(dp), intent(in) :: scale
        real(dp), dimension(4), intent(out) :: out
        integer :: i, j
        allocate(g(4, -1:6))
        do j = -1, 6
            do i = 1, 4
                g(i, j) = 0.1 * i + j
            enddo
        enddo
        ! strided section on the left, single-precision difference times a double on the right
        g(2::2, 0:5) = (g(:2, 1:6) - g(:2, -1:4)) * scale
        out(1) = g(2, 0)
        out(2) = g(4, 5)
        out(3) = ubound(g, 2) + lbound(g, 2)
        out(4) = sum(g(1, :))
    end subroutine s
    end module m
    """
    out = F.FArray.alloc(np.float64, [(1, 4)])
    run_f90(tmp_path, src, "s", np.float64(1.0) / 3.0, out)
    f4 = np.float32
    g = lambda i, j: f4(f4(f4(0.1) * f4(i)) + f4(j))  # noqa: E731
    assert out[1] == f4(np.float64(g(1, 1) - g(1, -1)) * (np.float64(1.0) / 3.0))
    assert out[2] == f4(np.float64(g(2, 6) - g(2, 4)) * (np.float64(1.0) / 3.0))
    assert out[3] == 5.0
    s = f4(0)
    for j in range(-1, 7):
        s = f4(s + g(1, j))
    assert out[4] == s


def test_f90run_calls_copy_back_and_do_loops(tmp_path):
    src = HEAD + """
    subroutine bump(x, k, t)
        real(dp), intent(inout) :: x
        integer, intent(out) :: k
        type(pt), intent(inout) :: t
        x = x + 1.0
        k = 5
        t%n = t%n + 1
        acc = acc + x
    end subroutine bump

    function twice(v) result(w)
        real(dp), intent(in) :: v
        real(dp) :: w
        w = 2 * v
    end function twice

    subroutine s(res)
        real(dp), dimension(6), intent(out) :: res
        real(dp) :: x
        integer :: k, i, n
        type(pt) :: t
        t%n = 0
        x = 1.0
        call bump(x, k, t)
        call bump(res(2), k, t)          ! array element as an intent(inout) actual
        res(1) = x
        res(3) = k + t%n
        n = 0
        do i = 10, 1, -3
            if (i == 4) cycle
            n = n + i
        enddo
        res(4) = n + 100 * i             ! i = -2 after completion
        do while (n > 0)
            n = n - 7
            if (n < 5) exit
        enddo
        res(5) = n
        res(6) = twice(acc)
    end subroutine s
    end module m
    """
    res = F.FArray.alloc(np.float64, [(1, 6)])
    res.a[:] = 0.0
    run_f90(tmp_path, src, "s", res)
    assert list(res.a) == [2.0, 1.0, 7.0, 10 + 7 + 1 + 100 * -2, 4.0, 2 * (2.0 + 1.0)]


def test_f90run_reading_an_unassigned_variable_is_an_error(tmp_path):
    src = HEAD + """
    subroutine s(y)
        real(dp), intent(out) :: y
        real(dp) :: never_set
        y = never_set * 2.0
    end subroutine s
    end module m
    """
    with pytest.raises(RuntimeError, match="before assigning"):
        run_f90(tmp_path, src, "s", None)


def test_f90run_cpp_conditionals_and_continuations(tmp_path):
    src = HEAD + """
    subroutine s(y)
        integer, intent(out) :: y
        y = 1
#if (defined USE_OPENMP)
        y = 2
#endif
        !$OMP PARALLEL
        y = y + &
            10 * &   ! a comment after the continuation mark
            3
    end subroutine s
    end module m
    """
    _, (y,) = run_f90(tmp_path, src, "s", None)
    assert y == 31


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
def test_tracked_split_equals_the_reference_split_particle():
    """split_particle + is_particle_selected (PM:5430-5480, 5920-5959) executed from the reference's source on the
    hand-made tracked population, against the C oracle: particles and particles_tracked, field by field."""
    from helpers import make_case, tracked_split_population
    from stochastic_parker_b200.abi import PARTICLE_DTYPE
    w, P, _, _ = make_case("c1", grid=16, nptl=8, conf=dict(dt_min_rel=1e-2))
    tags, ptl = tracked_split_population(P)
    r, o = refsim.RefSim(P, 32), Oracle(P, 32)
    for s in (r, o):
        s.init_tracking(tags, 10)
        s.upload_particles(ptl)
        s.split(2.0, 2.0, 10)
    a, b = r.download_particles(), o.download_particles()
    assert len(a) == len(b) == 11
    assert_particles_identical(a, b, "split")
    ra, rb = r.download_tracked(), o.download_tracked()
    assert ra.shape == rb.shape
    for name in PARTICLE_DTYPE.names:
        if name != "padding":
            assert np.array_equal(ra[name], rb[name]), name
    assert (ra["tag_splitted"] != 0).sum() == 4


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
def test_tracking_run_equals_the_reference():
    """The reference's two-run tracking workflow (hooks in inject_one_particle PM:434-440, the mover PM:1697-1724 /
    1806-1812 and split_particle PM:5452-5473, locate_particle / is_particle_selected with findloc) from its own
    source against the C oracle: final population bit for bit, particles_tracked record by record."""
    from helpers import make_case
    from stochastic_parker_b200.abi import PARTICLE_DTYPE
    from stochastic_parker_b200.driver import run_intervals
    from stochastic_parker_b200.tracking import select_tags
    nptl, nsel = 40, 6
    w, P, frames, ts = make_case("c1", grid=24, nptl=nptl, nframes=3, conf=dict(dt_min_rel=1e-3))
    P.strict_math = 1
    kw = dict(nptl=nptl, dist_flag=2, particle_v0=w.particle_v0, inject_new_ptl=False, split_flag=1,
              pmin_split=1.02, split_ratio=1.02, nsteps_interval=20)
    a = Oracle(P, 8 * nptl)
    run_intervals(a, frames, ts, **kw)
    first = a.download_particles()
    assert first["split_times"].max() >= 2
    tags = select_tags(first, np.argsort(first["p"])[-nsel:])
    out = []
    for cls in (Oracle, refsim.RefSim):
        recs = []
        b = cls(P, 8 * nptl)
        run_intervals(b, frames, ts, track_tags=tags, on_tracked=lambda tf, rec: recs.append(rec.copy()), **kw)
        out.append((b.download_particles(), recs))
    (pa, ra), (pb, rb) = out
    assert_particles_identical(pa, pb, "tracking run")
    assert len(ra) == len(rb) == 2
    for x, y in zip(ra, rb):
        assert x.shape == y.shape and (x["tag_splitted"] < 0).sum() >= 10
        for n in PARTICLE_DTYPE.names:
            if n != "padding":
                assert np.array_equal(x[n], y[n]), n


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("key,grid,mode,same", [("c1", 32, 1, True), ("c1", 32, 2, False), ("c3", 32, 4, True),
                                                ("c4", 32, 5, True), ("c5", 16, 1, True)])
def test_targeted_injectors_equal_the_reference(key, grid, mode, same):
    """inject_particles_at_large_jz / _absj / _divv / _rho + get_ncells_large_* (PM:785-1468, MD:2211-2498) from the
    reference's source (including get_ncells_large_divv's reallocation-on-assignment shift, MD:2423) against the
    C oracle: same cell count, same number of particles, same accepted positions and momenta, bit for bit."""
    from helpers import box_of, make_case
    w, P, frames, _ = make_case(key, grid=grid, nptl=8)
    r, o = refsim.RefSim(P, 600), Oracle(P, 600)
    for s in (r, o):
        s.upload_fields(0, frames[0])
        s.upload_fields(1, frames[1])
    box = box_of(P)
    box[0] += 0.1 * (P.xmax - P.xmin)
    box[4] -= 0.15 * (P.ymax - P.ymin)
    fa = o.get_fields(0).reshape(-1, 32)
    crit = {1: np.abs(fa[:, 8 + 15] - fa[:, 8 + 13]),
            2: np.sqrt((fa[:, 8 + 17] - fa[:, 8 + 19]) ** 2 + (fa[:, 8 + 18] - fa[:, 8 + 14]) ** 2
                       + (fa[:, 8 + 13] - fa[:, 8 + 15]) ** 2),
            4: -(fa[:, 8] + fa[:, 8 + 4] + (fa[:, 8 + 8] if P.ndim == 3 else 0.0)), 5: fa[:, 3]}[mode]
    vmin = float(np.quantile(crit, 0.7))
    rr = r.inject_targeted(mode, 60, 1e-4, 0, w.particle_v0, 0.0, 0.1, box, 6.2, same, vmin, 3 * grid)
    ro = o.inject_targeted(mode, 60, 1e-4, 0, w.particle_v0, 0.0, 0.1, box, 6.2, same, vmin, 3 * grid)
    assert rr == ro and rr[0] > 0 and rr[1] > 0
    assert_particles_identical(r.download_particles(), o.download_particles(), f"{key} mode {mode}")


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
def test_shock_injection_equals_the_reference_in_1d_and_is_undefined_beyond():
    """locate_shock_xpos + inject_particles_at_shock (MD:1988-2006, PM:542-633).  1-D: bit for bit.  2-D / 3-D:
    interp_shock_location sums into sx1 / sx2 without ever initialising them (MD:2037-2049) -- executing the
    reference's statements proves it; the library and the oracle start them at zero."""
    from helpers import make_case
    w, P, frames, _ = make_case("s1", grid=64, nptl=8)
    r, o = refsim.RefSim(P, 600), Oracle(P, 600)
    for s in (r, o):
        s.upload_fields(0, frames[0])
        s.upload_fields(1, frames[1])
        s.inject_at_shock(50, 1e-4, 2, w.particle_v0, 0.0, 6.2)
    assert_particles_identical(r.download_particles(), o.download_particles(), "1-D shock injection")
    w, P, frames, _ = make_case("c3", grid=32, nptl=8)
    r = refsim.RefSim(P, 600)
    r.upload_fields(0, frames[0])
    r.upload_fields(1, frames[1])
    with pytest.raises(RuntimeError, match="sx1' before assigning"):
        r.inject_at_shock(50, 1e-4, 2, w.particle_v0, 0.0, 6.2)


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("key,grid,cli", [("c1", 24, {}), ("c1", 24, dict(nlgc=1, kperp_kpara=0.05)), ("c5", 12, {})])
def test_turbulence_maps_equal_the_reference(key, grid, cli):
    """deltab_flag + correlation_flag: init_magnetic_fluctuation / init_correlation_length, calc_grad_sigma2_slab/_2d,
    calc_grad_lc_slab/_2d, interp_magnetic_fluctuation, interp_correlation_length, copy_* and their terms in both kappa
    routines (MD:107-182, 771-1604, 1806-1941; PM:2246-2254, 2505-2517), two intervals from the reference's source
    against the C oracle, bit for bit."""
    from helpers import make_case
    from stochastic_parker_b200 import mhd
    from stochastic_parker_b200.driver import run_intervals
    w, P, frames, ts = make_case(key, grid=grid, nptl=16, conf=dict(dt_min_rel=1e-3), cli=cli)
    P.deltab_flag, P.correlation_flag, P.strict_math = 1, 1, 1
    maps = [mhd.make_turbulence_maps(w.nx, w.ny, w.nz, f, P.ndim, w.dt_out) for f in range(3)]
    out = []
    for cls in (refsim.RefSim, Oracle):
        s = cls(P, w.nptl_max)
        rec, steps = run_intervals(s, frames, ts, nptl=16, particle_v0=w.particle_v0, split_flag=1, pmin_split=1.05,
                                   split_ratio=1.05, maps=lambda which, f: (maps[f][2 * which], maps[f][2 * which + 1]))
        out.append((s.download_particles(), steps, rec))
    assert out[0][1] == out[1][1] > 2000
    assert_particles_identical(out[0][0], out[1][0], "turbulence maps")
    assert np.array_equal(out[0][2][-1]["fglobal"], out[1][2][-1]["fglobal"])


@pytest.mark.skipif(not refsim.available(), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("name", ["c5_3d_acc_surfaces_union", "c5_3d_acc_surface_no_time_interp",
                                  "c5_3d_ft_acc_surfaces_intersection"])
def test_acceleration_surfaces_equal_the_reference(name):
    """acc_by_surface: init_acc_surface, interp_acc_surface, check_above_acc_surface, copy_acc_surface
    (acc_region_surface.f90) and the gate in push_particle_3d / _3d_ft, from the reference's source, bit for bit."""
    from helpers import CASES, make_case
    from stochastic_parker_b200 import mhd
    from stochastic_parker_b200.driver import run_intervals
    w, P, frames, ts = make_case(**dict(CASES[name], grid=12), nptl=12)
    P.strict_math = 1
    out = []
    for cls in (refsim.RefSim, Oracle):
        s = cls(P, w.nptl_max)
        rec, steps = run_intervals(s, frames, ts, nptl=12, particle_v0=w.particle_v0, split_flag=1, pmin_split=1.05,
                                   split_ratio=1.05, surfaces=lambda which, f: mhd.make_acc_surface(P, which, f))
        out.append((s.download_particles(), steps))
    assert out[0][1] == out[1][1] > 2000
    assert_particles_identical(out[0][0], out[1][0], name)
