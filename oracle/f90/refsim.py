"""RefSim -- the reference's OWN Fortran procedures (unmodified text under /root/reference/src/modules)
executed by oracle/f90/f90run.py, behind the method names of oracle/oracle.py::Oracle.

TEST INFRASTRUCTURE ONLY.  Used (a) by tests/golden/make_ref_f90_golden.py to generate the golden vectors
that pin oracle/gpat_oracle.c to the reference's arithmetic and (b) by the CPU tests, live, when
/root/reference is present.  Nothing here restates physics: every number comes out of the reference's
statements.  What the harness adds is only what the reference gets from outside the hot path:

  * module variables that read_particle_params / read_diagnostics_params / FLAP would have set
    (assigned from gpat_params; the reference's own setters are called where they exist);
  * MPI / OpenMP entry points for a single rank (no-ops; MPI_REDUCE copies send -> recv);
  * `unif_01`: mt_stream is a third-party library that is not in the tree, and north_star replaces it
    anyway; the stub serves the SAME uniforms the C oracle and the GPU use -- a table of pre-generated
    uniforms (GPAT_RNG_TABLE) or the per-particle Philox stream (keys and counters as in DESIGN.md section 4;
    Philox itself is pinned by the Random123 known-answer vectors, tests/test_cpu_oracle.py).
    The per-particle step counter rides in the record's `padding` field, as in the C ABI.
"""
from __future__ import annotations

import ctypes as C
import glob
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))

import f90run as F  # noqa: E402
from stochastic_parker_b200.abi import PARTICLE_DTYPE, Counters, Params  # noqa: E402  (POD layouts only)

REF_ROOT = os.environ.get("GPAT_REFERENCE", "/root/reference")
MODULE_FILES = ["constants", "mpi_module", "mhd_config", "simulation_setup", "mhd_data_parallel",
                "acc_region_surface", "particle_module", "diagnostics"]
NVAR = 32
f8 = np.float64


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "modules"))


def u01(w):
    return f8(w) / f8(4294967295.0)


_philox = None


def philox(ctr, key):
    global _philox
    if _philox is None:
        from oracle.oracle import philox4x32_10  # the KAT-pinned implementation (not reference arithmetic)
        _philox = philox4x32_10
    return _philox(ctr, key)


class RefSim:
    quick_is_average = True  # diagnostics()["quick"][5] is the reference's average dt, not the sum

    PUSHERS = ("push_particle_1d", "push_particle_2d", "push_particle_2d_include_3rd", "push_particle_3d",
               "push_particle_1d_ft", "push_particle_2d_ft", "push_particle_2d_include_3rd_ft",
               "push_particle_3d_ft")

    def __init__(self, params: Params, nptl_max: int, local_dist=True, dump_escaped_dist=True):
        if not available():
            raise RuntimeError(f"{REF_ROOT} is not present: RefSim runs the reference's own sources")
        self.P = params.copy()
        self.nptl_max = int(nptl_max)
        self.prog = prog = F.Program(defines=())
        prog.ext_values.update(real32=4, real64=8, real128=16, int8=1, int16=2, int32=4, int64=8,
                               mpi_status_size=6, mpi_comm_world=0, mpi_integer=1, mpi_double_precision=2,
                               mpi_sum=3, mpi_min=4, mpi_max=5, mpi_info_null=0, mpi_integer1=6, mpi_integer4=7)
        for f in MODULE_FILES:
            prog.load_file(os.path.join(REF_ROOT, "src", "modules", f + ".f90"))
        prog.init_module_data()
        self.M = {n: m.space for n, m in prog.modules.items()}
        self._install_externals()
        self.written = []
        self.steps = 0
        self._table = None
        self._setup(local_dist, dump_escaped_dist)

    # ---- plumbing -------------------------------------------------------------------------------
    def call(self, module, name, *args):
        with np.errstate(all="ignore"):
            return self.prog.get(module, name)(*args)

    def _install_externals(self):
        E = self.prog.externals
        noop = lambda *a, **k: ()  # noqa: E731
        for n in ("mpi_barrier", "mpi_bcast", "mpi_allreduce", "mpi_sendrecv", "mpi_send", "mpi_recv", "mpi_wait",
                  "mpi_isend", "mpi_irecv", "mpi_finalize", "mpi_type_commit", "mpi_type_free", "mpi_gather",
                  "mpi_allgather", "mpi_comm_split", "mpi_comm_rank", "mpi_comm_size"):
            E[n] = noop

        def mpi_reduce(send, recv, count, dtype, op, root, comm, ierr):
            if isinstance(send, F.FArray):
                recv.assign(send)
                return (None,)
            return (send,)
        E["mpi_reduce"] = mpi_reduce
        self.prog.ext_out["mpi_reduce"] = [1]

        def mpi_allreduce(send, recv, count, dtype, op, comm, ierr):
            if isinstance(send, F.FArray):
                recv.assign(send)
                return (None,)
            return (send,)
        E["mpi_allreduce"] = mpi_allreduce
        self.prog.ext_out["mpi_allreduce"] = [1]
        E["omp_get_thread_num"] = lambda: 0
        E["omp_get_num_threads"] = lambda: 1
        E["__write__"] = lambda unit, items: self.written.append((unit, items))
        E["unif_01"] = self._unif_01
        # single rank: nothing to exchange (the routine is pure MPI plumbing)
        self.prog.compiled[("particle_module", "send_recv_particles")] = lambda *a: ()

    def _wrap_pushers(self):
        """count push calls and tell unif_01 which particle is being pushed"""
        for name in self.PUSHERS:
            fn = self.prog.get("particle_module", name)
            ptl_index = self.prog.modules["particle_module"].procs[name].args.index("ptl")

            def wrapped(*a, _fn=fn, _i=ptl_index):
                ptl = a[_i]
                self._cur, self._k = ptl, 0
                r = _fn(*a)
                ptl.padding = ptl.padding + 1.0  # one RNG step per push (gpat_particle.padding)
                self.steps += 1
                self._cur = None
                return r
            self.prog.compiled[("particle_module", name)] = wrapped

    # ---- RNG stub -------------------------------------------------------------------------------
    def set_rng_table(self, u: np.ndarray):
        self._table = np.ascontiguousarray(u, dtype=np.float64)

    def _unif_01(self, thread_id):
        P = self.P
        if self._cur is not None:  # inside a pusher: uniform number k of this particle's current step
            ptl, k = self._cur, self._k
            self._k += 1
            step = int(ptl.padding)
            if P.rng_mode == 1 and self._table is not None:
                slot = abs(ptl.tag_injected)
                if k >= 4:
                    raise RuntimeError("the uniform table holds four numbers per step")
                if slot >= self._table.shape[0] or step >= self._table.shape[1]:
                    return f8(0.5)
                return f8(self._table[slot, step, k])
            if k == 0:
                ctr = [step & 0xFFFFFFFF, step >> 32, abs(ptl.tag_injected), abs(ptl.tag_splitted)]
                key = [P.seed & 0xFFFFFFFF, ((P.seed >> 32) + ptl.origin) & 0xFFFFFFFF]
                self._blk = philox(ctr, key)
            if k == 4:  # fifth uniform of the 3-D / include-3rd focused-transport pushers
                ctr = [step & 0xFFFFFFFF, (step >> 32) ^ 0x80000000, abs(ptl.tag_injected), abs(ptl.tag_splitted)]
                key = [P.seed & 0xFFFFFFFF, ((P.seed >> 32) + ptl.origin) & 0xFFFFFFFF]
                return u01(philox(ctr, key)[0])
            return u01(self._blk[k])
        # injection: stream (k/4, 0, tag_injected, 0) of the particle that will receive tag_max
        pm = self.M["particle_module"]
        tag = int(pm.tag_max)
        if tag != self._inj_tag:
            self._inj_tag, self._inj_k = tag, 0
        k = self._inj_k
        self._inj_k += 1
        if (k & 3) == 0:
            key = [P.seed & 0xFFFFFFFF, ((P.seed >> 32) + int(self.M["mpi_module"].mpi_rank)) & 0xFFFFFFFF]
            self._inj_blk = philox([k >> 2, 0, tag & 0xFFFFFFFF, 0], key)
        return u01(self._inj_blk[k & 3])

    # ---- set-up ---------------------------------------------------------------------------------
    def _setup(self, local_dist, dump_escaped_dist):
        P, M = self.P, self.M
        self._cur, self._inj_tag, self._inj_k = None, None, 0
        mp = M["mpi_module"]
        mp.mpi_rank, mp.mpi_size = int(P.mpi_rank), 1
        mp.mpi_sub_rank, mp.mpi_sub_size, mp.mpi_cross_rank, mp.mpi_cross_size = 0, 1, 0, 1
        mp.mpi_sub_comm, mp.mpi_cross_comm, mp.ierr = 0, 0, 0
        mc = M["mhd_config_module"]
        c = mc.mhd_config
        for n in ("dx", "dy", "dz", "xmin", "ymin", "zmin", "xmax", "ymax", "zmax", "lx", "ly", "lz", "nx", "ny", "nz"):
            setattr(c, n, getattr(P, n))
        c.nxs, c.nys, c.nzs, c.topox, c.topoy, c.topoz = P.nx, P.ny, P.nz, 1, 1, 1
        mc.uniform_grid_flag = not P.nonuniform_grid
        mc.spherical_coord_flag = bool(P.spherical_coord)
        self.prog.allocate("mhd_config_module", "tstamps_mhd", [(1, 2)])
        ss = M["simulation_setup_module"]
        ss.mpi_sizex = ss.mpi_sizey = ss.mpi_sizez = 1
        ss.mpi_ix = ss.mpi_iy = ss.mpi_iz = 0
        ss.pbcx, ss.pbcy, ss.pbcz = int(P.pbc[0]), int(P.pbc[1]), int(P.pbc[2])
        self.call("simulation_setup_module", "set_field_configuration", int(P.ndim))
        self.call("simulation_setup_module", "set_neighbors")
        self.call("mhd_data_parallel", "init_field_data", int(P.time_interp))
        self.call("mhd_data_parallel", "init_grid_positions")            # MAIN:198-199
        self.call("mhd_data_parallel", "set_local_grid_positions", "")
        if not P.time_interp:  # farray2 is referenced by name only when time_interp is true
            pass
        pm = M["particle_module"]
        for n in ("b0", "p0", "pmin", "pmax", "gamma_turb", "pindex", "kpara0", "kret", "dt_min_rel", "dt_max_rel",
                  "momentum_dependency", "mag_dependency", "acc_region_flag"):
            setattr(pm, n, getattr(P, n))
        for i in range(6):
            pm.acc_region[i + 1] = P.acc_region[i]
        self.call("particle_module", "set_dpp_params", int(P.dpp_wave), int(P.dpp_shear), int(P.weak_scattering),
                  f8(P.tau0))
        self.call("particle_module", "set_duu_params", f8(P.duu0))
        self.call("particle_module", "set_flags_params", int(P.deltab_flag), int(P.correlation_flag),
                  int(P.include_3rd_dim), int(P.acc_by_surface))
        self.call("particle_module", "set_drift_parameters", f8(P.drift1), f8(P.drift2), int(P.pcharge))
        self.call("particle_module", "set_flag_check_drift_2d", int(P.check_drift_2d))
        self.call("particle_module", "init_particles", self.nptl_max)
        pm.nptl_old = 0
        pm.nptl_escaped, pm.nptl_escaped_max = 0, self.nptl_max
        self.prog.allocate("particle_module", "escaped_ptls", [(1, self.nptl_max)])
        for e in pm.escaped_ptls.a:
            e.padding = 0.0
        for e in pm.ptls.a:
            e.padding = 0.0
        # diagnostics parameters (what read_diagnostics_params leaves behind; DG:2060-2200)
        dg = M["diagnostics"]
        dg.npp_global = int(P.npp_global)
        dg.nmu_global = int(P.nmu_global) if P.focused_transport else 1
        fc = ss.fconfig
        for k in range(4):
            s, K = P.local[k], str(k + 1)
            setattr(dg, "dump_local_dist" + K, bool(s.enabled))
            setattr(dg, "pmin" + K, s.pmin)
            setattr(dg, "pmax" + K, s.pmax)
            setattr(dg, "npbins" + K, int(s.npbins))
            setattr(dg, "nmu" + K, int(s.nmu) if P.focused_transport else 1)
            if s.enabled:
                setattr(dg, "rx" + K, int(s.rx))
                setattr(dg, "ry" + K, int(s.ry))
                setattr(dg, "rz" + K, int(s.rz))
                setattr(dg, "nrx" + K, (fc.nx + s.rx - 1) // s.rx)
                setattr(dg, "nry" + K, (fc.ny + s.ry - 1) // s.ry)
                setattr(dg, "nrz" + K, (fc.nz + s.rz - 1) // s.rz)
        self.local_dist, self.dump_escaped_dist = bool(local_dist), bool(dump_escaped_dist)
        self.call("diagnostics", "init_particle_distributions", self.local_dist, self.dump_escaped_dist)
        self._wrap_pushers()

    def close(self):
        pass

    # ---- fields ---------------------------------------------------------------------------------
    @property
    def grid_shape(self):
        P = self.P
        return (P.nz + 4 if P.ndim > 2 else P.nz, P.ny + 4 if P.ndim > 1 else P.ny, P.nx + 4)

    def _farray(self, slot):
        md = self.M["mhd_data_parallel"]
        return md.farray1 if slot == 0 else md.farray2

    def upload_fields(self, slot: int, f: np.ndarray, with_grad: int = 0):
        f = np.ascontiguousarray(f, dtype=np.float32)
        nvar = f.shape[-1]
        fa = self._farray(slot)
        fa.a[:nvar] = f.reshape(self.grid_shape + (nvar,)).transpose(3, 2, 1, 0)
        if not with_grad:
            self.call("mhd_data_parallel", "calc_fields_gradients", slot)

    def get_fields(self, slot: int) -> np.ndarray:
        return self._farray(slot).a.transpose(3, 2, 1, 0).copy(order='C')

    def swap_fields(self):
        """what MAIN:537-547 does at the end of an interval: copy_fields (+ surfaces, maps)"""
        self.call("mhd_data_parallel", "copy_fields")
        if self.P.acc_by_surface:
            self.call("acc_region_surface", "copy_acc_surface")
        if self.P.deltab_flag:
            self.call("mhd_data_parallel", "copy_magnetic_fluctuation")
        if self.P.correlation_flag:
            self.call("mhd_data_parallel", "copy_correlation_length")

    # ---- turbulence maps (MD:107-182, 306-497, 771-1604) and acceleration surfaces (acc_region_surface.f90) -------------
    def upload_turbulence(self, which: int, slot: int, slab: np.ndarray, two_d: np.ndarray):
        """read_magnetic_fluctuation (which = 0) / read_correlation_length (1): the file content goes into component 1
        of the slab / 2-D arrays, then the reference's own gradient passes fill components 2..4"""
        md = self.M["mhd_data_parallel"]
        names = (("sigma2_slab", "sigma2_2d", "init_magnetic_fluctuation", "calc_grad_sigma2_slab", "calc_grad_sigma2_2d"),
                 ("lc_slab", "lc_2d", "init_correlation_length", "calc_grad_lc_slab", "calc_grad_lc_2d"))[which]
        if md.__dict__.get(names[0] + "_1") is None:
            self.call("mhd_data_parallel", names[2])
        suffix = "_1" if slot == 0 else "_2"
        for nm, data in ((names[0], slab), (names[1], two_d)):
            arr = getattr(md, nm + suffix)
            arr.a[0] = np.ascontiguousarray(data, dtype=np.float32).reshape(self.grid_shape).transpose(2, 1, 0)
        self.call("mhd_data_parallel", names[3], slot)
        self.call("mhd_data_parallel", names[4], slot)

    def upload_acc_surface(self, which: int, slot: int, heights: np.ndarray):
        """read_acc_surface, whole-plane branch: heights (n2, n1) C-order == acc_surfaceK{1,2}(-1:n1+2, -1:n2+2)"""
        P, ar = self.P, self.M["acc_region_surface"]
        if ar.__dict__.get("acc_surface11") is None:
            def norm(v):
                return ("+" if v > 0 else "-") + "xyz"[abs(v) - 1]
            if P.surface2_existed:
                self.call("acc_region_surface", "init_acc_surface", int(P.time_interp), norm(P.surface_norm1),
                          norm(P.surface_norm2), bool(P.is_intersection))
            else:
                self.call("acc_region_surface", "init_acc_surface", int(P.time_interp), norm(P.surface_norm1))
        arr = getattr(ar, f"acc_surface{which + 1}{slot + 1}")
        h = np.ascontiguousarray(heights, dtype=np.float64)
        arr.a[...] = h.reshape(arr.a.shape[1], arr.a.shape[0]).T

    def interp(self, x, y, z, rt) -> np.ndarray:
        """get_interp_paramters + interp_fields at given points (the px/py/pz arithmetic of PM:1622-1624)"""
        ss, mc = self.M["simulation_setup_module"], self.M["mhd_config_module"]
        fc, c = ss.fconfig, mc.mhd_config
        out = np.empty((len(x), NVAR))
        pos = F.FArray.alloc(np.int64, [(1, 3)])
        w = F.FArray.alloc(np.float64, [(1, 8)])
        fields = F.FArray.alloc(np.float64, [(1, NVAR)])
        for i in range(len(x)):
            px = (f8(x[i]) - fc.xmin) / c.dx
            py = (f8(y[i]) - fc.ymin) / c.dy
            pz = (f8(z[i]) - fc.zmin) / c.dz
            self.call("particle_module", "get_interp_paramters", px, py, pz, pos, w)
            self.call("mhd_data_parallel", "interp_fields", pos, w, f8(rt[i]), fields)
            out[i] = fields.a
        return out

    # ---- particles ------------------------------------------------------------------------------
    def _set_tstamps(self, t0, dtf):
        t0, dtf = f8(t0), f8(dtf)
        t1 = t0 + dtf
        for cand in (t1, np.nextafter(t1, np.inf), np.nextafter(t1, -np.inf)):
            if cand - t0 == dtf:
                t1 = cand
                break
        else:
            raise ValueError("no t1 with t1 - t0 == dtf: pass frame times the way the driver does")
        ts = self.M["mhd_config_module"].tstamps_mhd
        ts[1], ts[2] = t0, t1

    def inject_uniform(self, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box, power_index):
        pm = self.M["particle_module"]
        self._set_tstamps(t_frame, dt_mhd)
        n0 = int(pm.nptl_current)
        box = F.FArray(np.array(part_box, dtype=np.float64))
        self.call("particle_module", "inject_particles_spatial_uniform", int(nptl), f8(dt), int(dist_flag),
                  f8(particle_v0), 1, box, f8(power_index))
        for i in range(n0 + 1, int(pm.nptl_current) + 1):
            pm.ptls[i].padding = 0.0

    INJECTORS = {1: ("inject_particles_at_large_jz", "get_ncells_large_jz"),
                 2: ("inject_particles_at_large_absj", "get_ncells_large_absj"),
                 3: ("inject_particles_at_large_db2", "get_ncells_large_db2"),
                 4: ("inject_particles_at_large_divv", "get_ncells_large_divv"),
                 5: ("inject_particles_at_large_rho", "get_ncells_large_rho")}

    def inject_targeted(self, mode, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box, power_index,
                        inject_same_nptl=True, vmin=0.0, ncells_norm=1):
        """inject_particles_at_large_jz / _absj / _db2 / _divv / _rho (PM:785-1468) and their cell counters
        (MD:2211-2498), whole-field decomposition; returns (nptl_inject, ncells) like the C ABI"""
        pm = self.M["particle_module"]
        self._set_tstamps(t_frame, dt_mhd)
        inj, cnt = self.INJECTORS[mode]
        box = F.FArray(np.array(part_box, dtype=np.float64))
        n0 = int(pm.nptl_current)
        nargs = len(self.prog.modules["mhd_data_parallel"].procs[cnt].args)
        cargs = (f8(vmin), bool(self.P.spherical_coord), box) if nargs == 3 else (f8(vmin), box)
        ncells = self.call("mhd_data_parallel", cnt, *cargs)
        self.call("particle_module", inj, int(nptl), f8(dt), int(dist_flag), f8(particle_v0), 1,
                  bool(inject_same_nptl), f8(vmin), int(ncells_norm), box, f8(power_index))
        for i in range(n0 + 1, int(pm.nptl_current) + 1):
            pm.ptls[i].padding = 0.0
        return int(pm.nptl_inject), int(ncells)

    def inject_at_shock(self, nptl, dt, dist_flag, particle_v0, t_frame, power_index):
        """locate_shock_xpos (MD:1988-2006) + inject_particles_at_shock (PM:542-633), as MAIN:451-454 calls them"""
        pm, md = self.M["particle_module"], self.M["mhd_data_parallel"]
        if md.__dict__.get("shock_xpos1") is None:
            self.call("mhd_data_parallel", "init_shock_xpos")
        ts = self.M["mhd_config_module"].tstamps_mhd
        ts[1] = f8(t_frame)
        n0 = int(pm.nptl_current)
        self.call("mhd_data_parallel", "locate_shock_xpos")
        self.call("particle_module", "inject_particles_at_shock", int(nptl), f8(dt), int(dist_flag), f8(particle_v0), 1,
                  f8(power_index))
        for i in range(n0 + 1, int(pm.nptl_current) + 1):
            pm.ptls[i].padding = 0.0

    def particle_mover(self, t0, dtf, nsteps_interval=100, num_fine_steps=1, dump_escaped_dist=0) -> int:
        P = self.P
        self._set_tstamps(t0, dtf)
        s0 = self.steps
        self.call("particle_module", "particle_mover", bool(P.focused_transport), bool(P.nlgc), f8(P.kperp_kpara),
                  int(nsteps_interval), 1, int(num_fine_steps), bool(dump_escaped_dist))
        return self.steps - s0

    def debug_push_n(self, t0, dtf, nsteps) -> int:
        """Twin of gpat_debug_push_n / orc_debug_push_n: exactly `nsteps` passes through the body of the
        reference's inner loop (PM:1603-1694: leak / boundary test, interpolation parameters, interp_fields,
        kappa, pusher with fixed_dt = .false., nsteps_pushed) for every in-box particle -- the loop is this
        harness's, every statement it executes is the reference's."""
        P, M = self.P, self.M
        pm, ss, mc = M["particle_module"], M["simulation_setup_module"], M["mhd_config_module"]
        fc, c = ss.fconfig, mc.mhd_config
        t0, dtf = f8(t0), f8(dtf)
        self.call("particle_module", "set_dt_min_max", dtf)
        half = np.float32(0.5)
        e = [fc.xmin - c.dx * half, fc.xmax + c.dx * half, fc.ymin - c.dy * half, fc.ymax + c.dy * half,
             fc.zmin - c.dz * half, fc.zmax + c.dz * half]  # PM:1524-1529
        pos = F.FArray.alloc(np.int64, [(1, 3)])
        w = F.FArray.alloc(np.float64, [(1, 8)])
        fields = F.FArray.alloc(np.float64, [(1, NVAR)])
        kcls = self.prog.struct_class("particle_module", "kappa_type")
        ft, nd = bool(P.focused_transport), int(P.ndim)
        if P.deltab_flag or P.correlation_flag or P.acc_by_surface:
            raise NotImplementedError("debug_push_n twin: maps / surfaces")
        none4 = F.FArray.alloc(np.float64, [(1, 4)])
        s0 = self.steps
        for i in range(1, int(pm.nptl_current) + 1):
            ptl = pm.ptls[i].copy()
            for _ in range(nsteps):
                if ptl.count_flag != 1:
                    break
                if ptl.p < 0.0:
                    ptl.count_flag = 0
                    pm.leak_negp = pm.leak_negp + ptl.weight
                else:
                    self.call("particle_module", "particle_boundary_condition", ptl, *e)
                if ptl.count_flag != 1:
                    break
                px, py, pz = (ptl.x - fc.xmin) / c.dx, (ptl.y - fc.ymin) / c.dy, (ptl.z - fc.zmin) / c.dz
                rt = (ptl.t - t0) / dtf
                self.call("particle_module", "get_interp_paramters", px, py, pz, pos, w)
                self.call("mhd_data_parallel", "interp_fields", pos, w, rt, fields)
                kappa = kcls()
                if P.nlgc:
                    self.call("particle_module", "calc_spatial_diffusion_coefficients_nlgc", ptl, ft, f8(P.kperp_kpara),
                              fields, none4, none4, none4, none4, kappa)
                else:
                    self.call("particle_module", "calc_spatial_diffusion_coefficients", ptl, ft, fields, none4, none4,
                              kappa)
                z = f8(0.0)
                if not ft:
                    if nd == 1:
                        self.call("particle_module", "push_particle_1d", 0, rt, ptl, fields, kappa, False, z, z)
                    elif nd == 2 and P.include_3rd_dim:
                        self.call("particle_module", "push_particle_2d_include_3rd", 0, rt, ptl, fields, kappa, False,
                                  z, z, z, z)
                    elif nd == 2:
                        self.call("particle_module", "push_particle_2d", 0, rt, ptl, fields, kappa, False, z, z, z)
                    else:
                        self.call("particle_module", "push_particle_3d", 0, rt, z, z, ptl, fields, kappa, False,
                                  z, z, z, z)
                else:
                    if nd == 2 and P.include_3rd_dim:
                        self.call("particle_module", "push_particle_2d_include_3rd_ft", 0, rt, ptl, fields, none4, none4,
                                  kappa, False, z, z, z, z, z, z)
                    elif nd == 2:
                        self.call("particle_module", "push_particle_2d_ft", 0, rt, ptl, fields, none4, none4, kappa,
                                  False, z, z, z, z, z)
                    elif nd == 3:
                        self.call("particle_module", "push_particle_3d_ft", 0, rt, z, z, ptl, fields, none4, none4, kappa,
                                  False, z, z, z, z, z, z)
                    else:
                        raise NotImplementedError("push_particle_1d_ft reads an unassigned dx_dt")
                ptl.nsteps_pushed = (ptl.nsteps_pushed + 1) % (1 << 30)
            pm.ptls[i] = ptl
        return self.steps - s0

    def split(self, split_ratio, pmin_split, nsteps_interval=100):
        self.call("particle_module", "split_particle", f8(split_ratio), f8(pmin_split), int(nsteps_interval))

    @staticmethod
    def _to_records(structs):
        out = np.zeros(len(structs), dtype=PARTICLE_DTYPE)
        for i, s in enumerate(structs):
            for n in PARTICLE_DTYPE.names:
                if n == "padding":
                    continue
                out[n][i] = getattr(s, n)
        out["padding"] = np.array([int(s.padding) for s in structs], dtype=np.uint64).view(np.float64)
        return out

    def download_particles(self) -> np.ndarray:
        pm = self.M["particle_module"]
        return self._to_records([pm.ptls[i] for i in range(1, int(pm.nptl_current) + 1)])

    def upload_particles(self, rec: np.ndarray):
        pm = self.M["particle_module"]
        steps = np.ascontiguousarray(rec["padding"]).view(np.uint64)
        for i in range(len(rec)):
            s = pm.ptls[i + 1]
            for n in PARTICLE_DTYPE.names:
                if n != "padding":
                    setattr(s, n, rec[n][i].item() if rec[n].dtype.kind == "i" else f8(rec[n][i]))
            s.padding = float(steps[i])
        pm.nptl_current = len(rec)

    def download_escaped(self) -> np.ndarray:
        pm = self.M["particle_module"]
        return self._to_records([pm.escaped_ptls[i] for i in range(1, int(pm.nptl_escaped) + 1)])

    def reset_escaped(self):
        self.M["particle_module"].nptl_escaped = 0

    # ---- particle tracking (PM:5825-5990; the HDF5 read of init_particle_tracking is the harness's) -------------
    def init_tracking(self, tags: np.ndarray, nsteps_interval: int):
        pm = self.M["particle_module"]
        tags = np.ascontiguousarray(tags, dtype=np.int32)  # (ntrack, ncols) C == (ncols, ntrack) Fortran
        ntrack, ncols = tags.shape
        pm.track_particle_flag = True
        pm.split_times_max, pm.nptl_tracking = ncols - 2, ntrack
        t = self.prog.allocate("particle_module", "tags_tracking", [(1, ncols), (1, ntrack)])
        t.a[...] = tags.T
        # nsteps_tracking_max = ceiling((1.0 / dt_min_rel) / nsteps_interval) + 1      (PM:5876)
        pm.nsteps_tracking_max = int(math.ceil((np.float32(1.0) / pm.dt_min_rel) / int(nsteps_interval))) + 1
        self.prog.allocate("particle_module", "particles_tracked", [(1, int(pm.nsteps_tracking_max)), (1, ntrack)])
        for e in pm.particles_tracked.a.reshape(-1):
            e.padding = 0.0
        self.call("particle_module", "reset_tracked_particles")

    def download_tracked(self) -> np.ndarray:
        pm = self.M["particle_module"]
        a = pm.particles_tracked.a  # (nsteps_tracking_max, nptl_tracking)
        out = np.zeros((a.shape[1], a.shape[0]), dtype=PARTICLE_DTYPE)
        for j in range(a.shape[1]):
            out[j] = self._to_records(list(a[:, j]))
        return out

    def reset_tracked(self):
        self.call("particle_module", "reset_tracked_particles")

    def counters(self) -> Counters:
        pm = self.M["particle_module"]
        c = Counters()
        c.nptl_current, c.nptl_split, c.nptl_escaped = int(pm.nptl_current), int(pm.nptl_split), int(pm.nptl_escaped)
        c.nptl_max, c.tag_max = int(pm.nptl_max), int(pm.tag_max)
        c.leak, c.leak_negp = float(pm.leak), float(pm.leak_negp)
        return c

    # ---- diagnostics ----------------------------------------------------------------------------
    def diagnostics(self, local_dist: bool = True):
        dg, P = self.M["diagnostics"], self.P
        self.call("diagnostics", "clean_particle_distributions", bool(local_dist))
        self.call("diagnostics", "calc_particle_distributions", bool(local_dist))
        fglobal = dg.fglobal.a.T.copy(order='C')
        flocal = []
        for k in range(4):
            if P.local[k].enabled and local_dist:
                flocal.append(getattr(dg, f"flocal{k + 1}").a.transpose(4, 3, 2, 1, 0).copy(order="C"))
            else:
                flocal.append(None)
        self.written.clear()
        self.call("diagnostics", "quick_check", 0, True, "")
        row = [w for w in self.written if w[0] == 17 and len(w[1]) >= 5][-1][1]
        # iframe, var_global(1:5), pdt_min_g, pdt_max_g, var_global(6)
        quick = np.zeros(8)
        quick[:5] = np.asarray(row[1])
        # layout of gpat_diagnostics' quick[8], except that slot 5 holds the reference's AVERAGE dt
        # (var_global(6) / var_global(1), DG:153-157) where the ABI returns the sum
        quick[5], quick[6], quick[7] = row[4], row[2], row[3]
        self.written.clear()
        self.call("diagnostics", "get_pmax_global", 0, True, "")
        pmax = float([w for w in self.written if w[0] == 17][-1][1][0])
        return dict(fglobal=fglobal, flocal=flocal, quick=quick, pmax=pmax)

    def hist_edges(self, which: int = 0):
        dg = self.M["diagnostics"]
        if which == 0:
            return dg.pbins_edges_global.a.copy(), dg.mubins_edges_global.a.copy()
        return getattr(dg, f"pbins{which}_edges").a.copy(), getattr(dg, f"mubins{which}_edges").a.copy()

    def escaped_diagnostics(self) -> np.ndarray:
        dg = self.M["diagnostics"]
        self.call("diagnostics", "clean_escaped_distributions", self.local_dist)
        self.call("diagnostics", "calc_escaped_distributions", self.local_dist)
        self._esc_done = True
        return dg.fescaped.a.transpose(2, 1, 0).copy(order='C')

    def escaped_local_diagnostics(self):
        """the per-face arrays calc_escaped_distributions filled in the same pass (DG:956-1170)"""
        dg, P = self.M["diagnostics"], self.P
        if not getattr(self, "_esc_done", False):
            self.escaped_diagnostics()
        self._esc_done = False
        out = []
        for k in range(4):
            if not P.local[k].enabled:
                out.append(None)
                continue
            d = {}
            for ax, need in (("x", 1), ("y", 2), ("z", 3)):
                arr = getattr(dg, f"fescaped{k + 1}_{ax}") if P.ndim >= need else None
                d[ax] = arr.a.transpose(4, 3, 2, 1, 0).copy(order='C') if arr is not None else None
            out.append(d)
        return out

    def set_counters(self, c: Counters):
        pm = self.M["particle_module"]
        pm.nptl_current, pm.nptl_split, pm.nptl_escaped = int(c.nptl_current), int(c.nptl_split), int(c.nptl_escaped)
        pm.tag_max, pm.leak, pm.leak_negp = int(c.tag_max), c.leak, c.leak_negp
