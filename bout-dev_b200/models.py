"""Harness-side drivers of the fused model right-hand sides (Python over the C-ABI).

These mirror what the reference does around PhysicsModel::rhs for the explicit RK solvers
(Solver::run_rhs src/solver/solver.cxx:1380-1427: pre_rhs boundary conditions -> model rhs;
RK4Solver::take_step src/solver/impls/rk4/rk4.cxx:172-210), with device-resident SoA fields: the
load_vars / save_derivs gather-scatter passes of the reference do not exist here and the stage
updates are written by the last RHS kernel (bout_b200_rk_epilogue).  Every numerical step is a
C-ABI call; torch only owns the memory.  The C++ host layer (host/) drives the same entry points.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import capi


class HW3D:
    """examples/hasegawa-wakatani-3d/hw.cxx: rhs() lines 50-77, bndry_all = dirichlet_o2,
    y periodic (ixseps = nx), identity parallel transform, LaplaceCyclic with default flags."""

    def __init__(self, g: capi.Grid, metric: dict, alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3,
                 device="cuda", mesh=None):
        self.g, self.dev = g, torch.device(device)
        self.md = {k: v.to(self.dev) for k, v in metric.items()}
        self.alpha, self.kappa, self.Dn, self.Dvort = alpha, kappa, Dn, Dvort
        self.cm = capi.make_metric(self.md)
        self.mesh = mesh  # None = single rank (self exchange)
        zlength = float(self.md["dz"][0, 0].item()) * g.nz
        first_x = mesh is None or mesh.first_x
        last_x = mesh is None or mesh.last_x
        # Guard cells without separate passes (BOUT_B200_GUARD_FILL=0 keeps one launch per boundary / exchange):
        # the fused kernel's epilogue writes the x guard cells of the stage state it produces (the Dirichlet condition
        # Solver::pre_rhs would apply) and, on a single rank, its periodic y guard rows; the Laplacian solve writes
        # phi's y guard rows.  An RHS then is: solve (3 launches) + fused kernel.
        self.guard_fill = (not (mesh is not None and mesh.nxpe > 1) and self.dev.type == "cuda"
                           and os.environ.get("BOUT_B200_GUARD_FILL", "1") != "0")
        single = mesh is None
        self.lap = capi.Laplace(g, self.md, zlength, nonuniform=True, first_x=first_x, last_x=last_x,
                                wrap_yguards=single and self.guard_fill)
        self.lap.workspace(self.dev)
        self.gf = capi.make_guard_fill(g, 0.0, first_x, last_x, wrap_y=single) if self.guard_fill else None
        # (n, vort) pairs whose guard cells were written by the kernel that produced them: data pointers -> tensor
        # versions at that moment (any later in-place torch operation on them bumps the version and voids the entry)
        self._ready = {}
        self.x_split = mesh is not None and mesh.nxpe > 1
        if self.x_split:  # interface records of the X-decomposed tridiagonal solve
            nif = self.lap.iface_doubles()
            self.iface = torch.empty(nif, dtype=torch.float64, device=self.dev)
            self.iface_all = torch.empty(nif * mesh.nxpe, dtype=torch.float64, device=self.dev)
        self.coef = capi.fused_coefs(g, self.cm, capi.fused_coef_workspace(g, self.dev))
        self.phase_events = None
        self.kernel_events = None  # bench.py: list collecting (start, end) CUDA events of the fused kernel
        # NYPE > 1: the guard rows of n and vort travel on a second stream while the Laplacian solve runs on the
        # main one (they do not depend on it); phi's rows travel while the fused kernel works on the rows that do not
        # need them (all but the first and the last interior row), which then follow on the second stream
        self.overlap = (mesh is not None and mesh.nxpe == 1 and mesh.nype > 1 and self.dev.type == "cuda"
                        and os.environ.get("BOUT_B200_HALO_OVERLAP", "1") != "0")
        self.overlap_phi = self.overlap and g.yend - g.ystart >= 2 and os.environ.get("BOUT_B200_PHI_OVERLAP", "1") != "0"
        # row ranges of the fused kernel's launches: the rows within `w` of an end (they have y images / wait for a
        # neighbour) and all the others
        w = max(g.myg, 1) if self.guard_fill else 1
        self.edge_w = w
        self.split_ok = g.yend - g.ystart + 1 >= 2 * w + 1
        if self.split_ok:
            self.g_mid = capi.Grid(g.nx, g.ny, g.nz, g.mxg, g.myg)
            self.g_lo = capi.Grid(g.nx, g.ny, g.nz, g.mxg, g.myg)
            self.g_hi = capi.Grid(g.nx, g.ny, g.nz, g.mxg, g.myg)
            for gg, (a, b) in ((self.g_mid, (g.ystart + w, g.yend - w)), (self.g_lo, (g.ystart, g.ystart + w - 1)),
                               (self.g_hi, (g.yend - w + 1, g.yend))):
                gg.c.ystart, gg.c.yend = a, b
        self.overlap_phi = self.overlap_phi and self.split_ok
        self._comm_stream = None
        z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device=self.dev)
        # NYPE > 1 on one node (BOUT_B200_YPEER=0 keeps the exchange kernels): no exchange pass at all -- the fused
        # kernel and the Laplacian solve store the rows a neighbour needs straight into its guard rows over NVLink
        # (bout_b200_ypeer) and raise its flags; only the CTAs of the edge rows, launched last, ever wait.  The fields
        # that take part live in a peer-mapped arena; phi alternates between two buffers so that a neighbour that is
        # one solve ahead does not overwrite guard rows this rank's fused kernel is still reading.
        self.ypeer = (self.overlap and self.guard_fill and g.nz in (128, 256, 512) and getattr(mesh, "_peer_ok", False)
                      and mesh.rank not in (mesh.yup, mesh.ydown) and g.yend - g.ystart + 1 >= 2 * max(g.ystart, 1)
                      and os.environ.get("BOUT_B200_YPEER", "1") != "0")
        if self.ypeer:
            from .mesh import PeerFieldArena
            self.arena = PeerFieldArena(mesh, 8, self.dev)
            self.n, self.vort, self.sn, self.sv, self.tn, self.tv, ph0, ph1 = self.arena.fields
            self.phis = [ph0, ph1]
            self.phi = ph0
            self.an, self.av = z(), z()
            self._cstate = self._cphi = 0   # deliveries so far on the two channels (the same on every rank)
            self.overlap_phi = False
        else:
            self.n, self.vort, self.phi = z(), z(), z()
            # RK work fields: stage state (k5 in rk4.cxx) and the running combination
            self.sn, self.sv, self.tn, self.tv, self.an, self.av = z(), z(), z(), z(), z(), z()
        self.ddt_n, self.ddt_vort = z(), z()

    def set_state(self, n, vort):
        self.n.copy_(n)
        self.vort.copy_(vort)
        self._ready.clear()

    # -- one right-hand side on (n, vort) held in fields fn, fv
    def _pre_rhs(self, fn, fv):
        g = self.g
        e = []
        if self.mesh is None or self.mesh.first_x:
            e += [("dirichlet", "xin", fn), ("dirichlet", "xin", fv)]
        if self.mesh is None or self.mesh.last_x:
            e += [("dirichlet", "xout", fn), ("dirichlet", "xout", fv)]
        if e:
            capi.boundary_batch(g, e)

    def _mark(self, name):
        """bench.py / tools: optional per-phase CUDA events (None = off, no overhead)."""
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append((name, e))

    def _fused_split(self, fn, fv, epilogue, ddt, guards):
        """The fused kernel as three launches: the few edge rows (y images, waits for neighbour GPUs: the full guard
        variant of the kernel, short x chunks) on a second, high-priority stream beside all the other rows (the plain
        march + x guard cells) on this one."""
        if not self.split_ok:
            return self._fused(self.g, fn, fv, epilogue, ddt, guards)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        main, side = torch.cuda.current_stream(), self._comm_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._fused(self.g_lo, fn, fv, epilogue, ddt, guards)
            self._fused(self.g_hi, fn, fv, epilogue, ddt, guards)
        self._fused(self.g_mid, fn, fv, epilogue, ddt, guards)
        main.wait_stream(side)

    def _fused(self, gsel, fn, fv, epilogue, ddt, guards):
        capi.hw3d_rhs(gsel, None, self.alpha, self.kappa, self.Dn, self.Dvort, fn, fv, self.phi,
                      self.ddt_n if ddt else None, self.ddt_vort if ddt else None, epilogue=epilogue,
                      coef_ws=self.coef, guards=guards)

    def _rhs_ypeer(self, fn, fv, epilogue, ddt):
        """One right-hand side with the Y guard rows delivered by the producing kernels (bout_b200_ypeer)."""
        A = self.arena
        ready = self._ready.get((fn.data_ptr(), fv.data_ptr())) == (fn._version, fv._version)
        self._mark("start")
        if not ready:   # a state set from outside: boundary conditions and one ordinary exchange
            self._pre_rhs(fn, fv)
            self.mesh.communicate([fn, fv])
        self._mark("bc")
        self.phi = self.phis[self._cphi % 2]
        self._cphi += 1
        self.lap.set_ypeer(A.ypeer(1, self._cphi))
        self.lap.solve(fv, out=self.phi)
        self._mark("solve")
        self._mark("halo")
        nxt = None
        if epilogue is not None:
            k = epilogue._keep
            nxt = tuple((k[2][i] if k[2] is not None and k[2][i] is not None else k[3][i]) for i in range(2))
        deliver = nxt is not None and all(t is not None and t.data_ptr() in A.ptrs for t in nxt)
        # edge rows wait for: the neighbour's rows of this state (when they were delivered by its kernel, not by the
        # exchange above) and its rows of phi
        waits = [[(A.flag(0, d), self._cstate) if ready else None, (A.flag(1, d), self._cphi)] for d in range(2)]
        for d, nb in enumerate((self.mesh.ydown, self.mesh.yup)):
            if nb is None:
                waits[d] = [None, None]
        yp = A.ypeer(0, self._cstate + 1, deliver=deliver, waits=waits)
        gf = capi.make_guard_fill(self.g, 0.0, self.gf.x_inner and epilogue is not None, self.gf.x_outer and epilogue is not None,
                                  wrap_y=False, ypeer=yp)
        if self.kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._fused_split(fn, fv, epilogue, ddt, gf)
        if self.kernel_events is not None:
            e1.record()
            self.kernel_events.append((e0, e1))
        self._mark("fused")
        if deliver:
            self._cstate += 1
            self._ready = {(nxt[0].data_ptr(), nxt[1].data_ptr()): (nxt[0]._version, nxt[1]._version)}
        elif nxt is not None and nxt[0] is not None and nxt[1] is not None:
            self._ready.pop((nxt[0].data_ptr(), nxt[1].data_ptr()), None)

    def rhs(self, fn=None, fv=None, epilogue=None, ddt=True):
        fn = self.n if fn is None else fn
        fv = self.vort if fv is None else fv
        if self.ypeer:
            return self._rhs_ypeer(fn, fv, epilogue, ddt)
        # were the guard cells of this state written by the kernel that produced it?
        ready = self._ready.get((fn.data_ptr(), fv.data_ptr())) == (fn._version, fv._version)
        guards = self.gf if (self.guard_fill and epilogue is not None) else None
        self._mark("start")
        if not ready:
            self._pre_rhs(fn, fv)                               # Solver::pre_rhs
        self._mark("bc")
        overlapped = False
        if self.overlap and not self.x_split:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.dev, priority=-1)
            main = torch.cuda.current_stream()
            self._comm_stream.wait_stream(main)
            with torch.cuda.stream(self._comm_stream):
                self.mesh.communicate([fn, fv])
            overlapped = True
        if not self.x_split:                                    # phi = phiSolver->solve(vort, phi)
            self.lap.solve(fv, out=self.phi)
            self._mark("solve")
        else:
            self.lap.reduce(fv, iface=self.iface)
            self._mark("solve_reduce")
            self.mesh.all_gather_x(self.iface, self.iface_all)
            self._mark("solve_allgather")
            self.lap.finish(self.iface_all, self.mesh.nxpe, self.mesh.pe_xind, self.phi)
            self._mark("solve_finish")
        split_rows = False
        if self.mesh is None:                                   # mesh->communicate(n, vort, phi, ...)
            need = ([] if ready else [fn, fv]) + ([] if self.lap.c.wrap_yguards else [self.phi])
            if need:
                capi.halo_self(self.g, need, periodic_x=False, periodic_y=True)
        elif overlapped and self.overlap_phi:
            split_rows = True
        elif overlapped:
            self.mesh.communicate([self.phi])
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        else:
            self.mesh.communicate([fn, fv, self.phi])
        self._mark("halo")
        if self.kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if split_rows:
            # the rows that need no y guard row start at once; phi's rows travel meanwhile on the second (high
            # priority) stream, followed there by the first and the last interior row
            main = torch.cuda.current_stream()
            self._comm_stream.wait_stream(main)
            self._fused(self.g_mid, fn, fv, epilogue, ddt, guards)
            with torch.cuda.stream(self._comm_stream):
                self.mesh.communicate([self.phi])
                self._fused(self.g_lo, fn, fv, epilogue, ddt, guards)
                self._fused(self.g_hi, fn, fv, epilogue, ddt, guards)
            main.wait_stream(self._comm_stream)
        elif guards is not None:
            self._fused_split(fn, fv, epilogue, ddt, guards)
        else:
            self._fused(self.g, fn, fv, epilogue, ddt, guards)
        if self.kernel_events is not None:
            e1.record()
            self.kernel_events.append((e0, e1))
        self._mark("fused")
        if guards is not None:  # the field the next right-hand side reads now has its guard cells
            e = epilogue
            nxt = tuple((e._keep[2][i] if e._keep[2] is not None and e._keep[2][i] is not None else e._keep[3][i])
                        for i in range(2))
            if nxt[0] is not None and nxt[1] is not None:
                # x guard cells always; the y guard rows only on a single rank (otherwise they are exchanged anyway)
                self._ready[(nxt[0].data_ptr(), nxt[1].data_ptr())] = (nxt[0]._version, nxt[1]._version)

    def phase_summary(self):
        """Mean milliseconds per phase from the recorded events (after torch.cuda.synchronize())."""
        tot, cnt = {}, {}
        ev = self.phase_events
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            if n1 == "start":
                continue
            tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        return {k: tot[k] / cnt[k] for k in tot}

    def rk4_step(self, dt):
        """RK4Solver::take_step with the axpys folded into the RHS kernel:
        stage i reads state_i, writes state_{i+1} = start + a_i dt k_i and acc += b_i dt k_i."""
        n, v, sn, sv, tn, tv, an, av = self.n, self.vort, self.sn, self.sv, self.tn, self.tv, self.an, self.av
        E = capi.make_epilogue
        # the stage state a kernel writes must not be the one it reads (neighbour cells): ping-pong s/t
        self.rhs(n, v, E((n, v), None, (sn, sv), (an, av), 0.5 * dt, dt / 6.0), ddt=False)
        self.rhs(sn, sv, E((n, v), (an, av), (tn, tv), (an, av), 0.5 * dt, dt / 3.0), ddt=False)
        self.rhs(tn, tv, E((n, v), (an, av), (sn, sv), (an, av), dt, dt / 3.0), ddt=False)
        self.rhs(sn, sv, E((n, v), (an, av), (None, None), (n, v), 0.0, dt / 6.0), ddt=False)


class ELMpb:
    """examples/elm-pb/elm_pb.cxx rhs() (lines 1240-2030) with the shipped [highbeta] switches
    (examples/elm-pb/data/BOUT.inp): linear, evolve Psi, diamag + diamag_phi0, constn0, bm_*_flag = 0
    (BRACKET_STD -> b0xGrad_dot_Grad), hyper-resistivity, smooth_j_x, filter_z + low_pass_z; shifted-metric
    parallel transform, all-core grid (y periodic), methods C4 / C2 / C4 (first), W3 / C2 / W3 (upwind).

    Inputs are the POST-INIT quantities of the model (init() is set-up time, not on the per-RHS path):
    `md` the Coordinates Field2D set after ELMpb::init's metric assignment + geometry(); `eq` =
    {J0, P0, phi0, B0, b0xcv_x, b0xcv_y, b0xcv_z (Field2D), eta (Field3D)}; scalars dnorm, hyperresist.
    Every numerical step is a C-ABI call.  Evolving fields: U, P, Psi (solver order of SOLVE_FOR)."""

    def __init__(self, g: capi.Grid, md: dict, eq: dict, zshift, zlength, dnorm, hyperresist,
                 filter_z_mode=1, low_pass_z=16, zonal=(False, True, False), device="cuda", mesh=None):
        self.g, self.dev = g, torch.device(device)
        # mesh: a BoutMesh with NXPE = 1 (the domain split along Y only: every operator with X or Z coupling --
        # LaplaceCyclic, Delp2, smooth_x, the z FFTs -- stays rank-local and only communicate() crosses ranks);
        # g, md, eq, zshift are then this rank's slab.  None = single rank.
        if mesh is not None and mesh.nxpe != 1:
            raise ValueError("ELMpb: only NYPE decomposition (NXPE = 1) is implemented")
        self.mesh = mesh
        dv = lambda t: t.to(self.dev).contiguous()
        self.md = {k: dv(v) for k, v in md.items()}
        self.eq = {k: dv(v) for k, v in eq.items()}
        self.zs, self.zl = dv(zshift), float(zlength)
        self.mzs = (-self.zs).contiguous()
        self.dnorm, self.hyperresist = float(dnorm), float(hyperresist)
        self.filter_z_mode, self.low_pass_z = int(filter_z_mode), int(low_pass_z)
        self.zonal_flow, self.zonal_bkgd, self.zonal_field = zonal      # for ddt(U), ddt(P), ddt(Psi)
        m = self.md
        # phiSolver: [phiSolver] inner_boundary_flags = 4 (AC_LAP), outer = 1 + 4 (DC_GRAD + AC_LAP)
        self.lap = capi.Laplace(g, m, self.zl, all_terms=True, nonuniform=True, flags=(0, 4, 5))
        self.lap.workspace(self.dev)
        self.delp = capi.Laplace(g, m, self.zl, all_terms=True, nonuniform=True)   # Delp2's default Laplacian
        self.b0x = capi.make_b0xgrad(m, first=("C4", "C2", "C4"), upwind=("W3", "C2", "W3"))
        self.rdx = (1.0 / m["dx"]).contiguous()
        self.rdy = (1.0 / m["dy"]).contiguous()
        self.rdz = (1.0 / m["dz"]).contiguous()
        self.gradpar = (self.rdy / torch.sqrt(m["g_22"])).contiguous()       # DDY / dy * invSg
        self.B0 = self.eq["B0"]
        self.B0sq = (self.B0 * self.B0).contiguous()
        z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device=self.dev)
        self.U, self.P, self.Psi = z(), z(), z()
        self.phi, self.Jpar, self.Jpar2 = z(), z(), z()
        self.ddt_U, self.ddt_P, self.ddt_Psi = z(), z(), z()
        self.fused = True             # one kernel for the three equations + filter (csrc/elm_rhs.cu)
        self.two_streams = os.environ.get("BOUT_B200_ELM_STREAMS", "2") != "1"
        self.mode_slices = os.environ.get("BOUT_B200_ELM_MODES", "1") != "0"
        self._side = None
        self._side2 = None
        self.three_streams = os.environ.get("BOUT_B200_ELM_STREAMS", "2") == "3"   # measured: 1.003 ms vs 0.966 ms with two
        self.kernel_events = None     # bench.py: (start, end) CUDA events of the fused kernel
        e3 = self.eq["eta"]
        self.eta2d = e3[:, :, 0].contiguous() if (e3.dim() == 3 and bool((e3 == e3[:, :, :1]).all())) else None
        self.compute_jpar2 = False    # the reference evaluates Jpar2 = Delp2(Jpar) although no shipped term reads it

    @classmethod
    def from_setup(cls, s, device="cuda", mesh=None, **kw):
        """Build from elm_setup.elmpb_setup(...) of the GLOBAL grid (numpy, set-up time); with a mesh, each
        rank keeps its own y slab (guard rows included) of every array."""
        import numpy as np
        nx, ny, nz, mxg, myg = s["grid"]
        if mesh is None:
            g, ysl = capi.Grid(*s["grid"]), slice(None)
        else:
            g = mesh.grid
            if (ny - 2 * myg) != mesh.mysub * mesh.nype or g.nx != nx or g.nz != nz:
                raise ValueError("ELMpb.from_setup: mesh does not tile the set-up grid")
            ysl = slice(mesh.pe_yind * mesh.mysub, mesh.pe_yind * mesh.mysub + mesh.mysub + 2 * myg)
        cut = lambda v: torch.from_numpy(np.ascontiguousarray(v[:, ysl]))
        md = {k: cut(v) for k, v in s["md"].items()}
        eq = {k: cut(v) for k, v in s["eq"].items()}
        return cls(g, md, eq, cut(s["zShift"]), s["zlength"], s["dnorm"], s["hyperresist"], device=device,
                   mesh=mesh, **kw)

    def set_state(self, U, P, Psi):
        self.U.copy_(U), self.P.copy_(P), self.Psi.copy_(Psi)

    def _slices(self, f):
        return capi.parallel_slices(self.g, f, self.zs, self.zl, 1)

    def _comm(self, fields):
        """mesh->communicate(...) on one rank: y periodic self exchange; ShiftedMetric then builds the
        parallel slices of each field (mesh.cxx:349-378)."""
        self._halo(fields)
        return [self._slices(f) for f in fields]

    def _halo(self, fields):
        """Guard-cell exchange of Mesh::communicate: y periodic (all-core grid), no X neighbours (NXPE = 1)."""
        if self.mesh is None:
            capi.halo_self(self.g, fields, periodic_x=False, periodic_y=True)
        else:
            self.mesh.communicate(fields)

    def _bc_j(self, f):          # [J] bndry_core = neumann, rest dirichlet_o2 (both sides in one launch)
        capi.boundary_batch(self.g, [("neumann", "xin", f, 0.0, self.md["dx"]), ("dirichlet", "xout", f)])

    def _bc_psi(self, f):        # [Psi] zerolaplace on both X sides
        m = self.md
        capi.boundary_zerolaplace(self.g, "both", f, m["dx"], m["g11"], m["g33"], self.zl)   # one launch

    def pre_rhs(self, U, P, Psi):
        """Solver::pre_rhs (solver.cxx:1486-1500): boundary conditions of the evolving fields."""
        g = self.g
        capi.boundary_batch(g, [("dirichlet", "xin", U), ("dirichlet", "xout", U),
                                ("neumann", "xin", P, 0.0, self.md["dx"]),      # [P] bndry_core = neumann
                                ("dirichlet", "xout", P)])                      # one launch for the four
        self._bc_psi(Psi)

    def _grad_par_noslices(self, f):
        """Grad_par of a field WITHOUT parallel slices: DDY goes through toFieldAligned(RGN_NOX) ->
        index DDY C2 -> fromFieldAligned (index_derivs_interface.hxx:205-211), / dy, * invSg."""
        g = self.g
        fa = capi.zshift(g, f, self.mzs, self.zl, region="RGN_NOX")
        d = capi.stencil(g, "DDX_C2", "Y", fa, scale1=self.gradpar)
        return capi.zshift(g, d, self.zs, self.zl, region="RGN_NOBNDRY")

    def _buffers(self):
        if not hasattr(self, "_buf"):
            z = lambda: torch.zeros(self.g.shape3, dtype=torch.float64, device=self.dev)
            self._buf = {n: (z(), z()) for n in ("U", "P", "Psi", "phi", "J")}   # (yup, ydown) per field
            self._buf["sm"] = z()
            self._buf["d2J"] = z()
            nrc = capi.lib.bout_b200_elmpb_rowc_doubles()
            self._buf["rowc_ws"] = torch.zeros(self.g.nx * self.g.ny * nrc, dtype=torch.float64, device=self.dev)
            for n in ("mU", "mP", "mJ"):   # one Fourier coefficient per (x,y) row, complex
                self._buf[n] = torch.zeros(self.g.shape2 + (2,), dtype=torch.float64, device=self.dev)
        return self._buf

    def _slices_into(self, f, name):
        up, dn = self._buffers()[name]
        capi.parallel_slices(self.g, f, self.zs, self.zl, 1, yup=[up], ydown=[dn])
        return ([up], [dn])

    def _comm_into(self, fields, names):
        """mesh->communicate(...) with the parallel slices written into preallocated buffers."""
        self._halo(fields)
        return [self._slices_into(f, n) for f, n in zip(fields, names)]

    def rhs(self, U=None, P=None, Psi=None, rk=None):
        """rk = (start (U, P, Psi), stage state (U, P, Psi), c * dt): also form the next RK stage state."""
        g, m, eq = self.g, self.md, self.eq
        U = self.U if U is None else U
        P = self.P if P is None else P
        Psi = self.Psi if Psi is None else Psi
        ew = capi.elementwise
        buf = self._buffers()
        fused = self.fused and 0 <= self.filter_z_mode < g.nz // 2
        # y-derivative terms of U, P, Jpar in mode space (one Fourier coefficient per row) instead of full slices
        modes = fused and self.mode_slices
        self.pre_rhs(U, P, Psi)
        self._halo([U, P, Psi])                                               # mesh->communicate(comms)
        # The potential chain (LaplaceCyclic -> phi update -> communicate -> slices) runs on a second stream: its
        # Thomas kernel has only ny * (nz/2 + 1) systems (65 CTAs at 64 x 129) and its small launches leave SMs
        # idle, which the slice / Delp2 kernels of the main stream fill.  Both streams only READ U, P, Psi.
        two = self.two_streams and self.dev.type == "cuda"
        main = torch.cuda.current_stream() if two else None
        if two:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.dev)
            self._side.wait_stream(main)
        with (torch.cuda.stream(self._side) if two else contextlib.nullcontext()):
            # phi = phiSolver->solve(U) - 0.5 dnorm P / B0 (constn0, diamag)       elm_pb.cxx:1472-1477
            self.lap.solve(U, out=self.phi)
            ew(g, [(1.0, self.phi), (-0.5 * self.dnorm, P, None, self.B0)], out=self.phi)
            (sphi,) = self._comm_into([self.phi], ["phi"])
            capi.boundary_batch(g, [("extend", "xin", self.phi), ("extend", "xout", self.phi)])   # :1498-1516
        # The current chain (Delp2(Psi) -> BC -> communicate -> smooth_x -> ... -> row mode of Jpar: nine short launches
        # that depend on each other) is issued first; BOUT_B200_ELM_STREAMS=3 puts it on a third stream beside the
        # slice / row-mode kernels of the main one (no gain measured: the GPU is busy, not waiting)
        three = two and modes and self.three_streams
        if three:
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=self.dev)
            self._side2.wait_stream(main)
        with (torch.cuda.stream(self._side2) if three else contextlib.nullcontext()):
            # Jpar = Delp2(Psi); BC; communicate; smooth_x; BC; communicate         :1518-1563
            self.delp.delp2(Psi, out=self.Jpar)
            self._bc_j(self.Jpar)
            self._halo([self.Jpar])
            sm = capi.smooth_x(g, self.Jpar, out=buf["sm"])
            self._halo([sm])                                                      # smooth_x communicates its result
            buf["sm"], self.Jpar = self.Jpar, sm                                  # Jpar = smooth_x(Jpar)
            self._bc_j(self.Jpar)
            if modes:
                self._halo([self.Jpar])
                capi.row_mode(g, [self.Jpar], self.filter_z_mode, out=[buf["mJ"]])
                sJ = None
        if modes:
            sPsi = self._slices_into(Psi, "Psi")
            capi.row_mode(g, [U, P], self.filter_z_mode, out=[buf["mU"], buf["mP"]])
            sU = sP = None
        else:
            sU, sP, sPsi = [self._slices_into(f, n) for f, n in ((U, "U"), (P, "P"), (Psi, "Psi"))]
            (sJ,) = self._comm_into([self.Jpar], ["J"])
        if three:
            main.wait_stream(self._side2)
        # Delp2(Jpar): Jpar2 (:1565-1572) and the hyper-resistivity term (:1717-1723) evaluate the same field.
        # With a z-independent eta the term is linear with z-independent coefficients, so (after filter_z) the
        # fused kernel forms mode n0 of Delp2(Jpar) from the row modes of Jpar and the full-field pass is skipped
        # unless Jpar2 itself is wanted (no shipped term reads it and it is not an output).
        skip_d2j = modes and self.eta2d is not None and not self.compute_jpar2
        d2J = None if skip_d2j else self.delp.delp2(self.Jpar, out=buf["d2J"])
        if self.compute_jpar2:
            self.Jpar2.copy_(d2J)
            self._bc_j(self.Jpar2)
            self._halo([self.Jpar2])
        if two:
            main.wait_stream(self._side)
        if fused:
            keep = [1.0, 1.0, 1.0]
            if self.low_pass_z > 0:
                for i, zonal in enumerate((self.zonal_flow, self.zonal_bkgd, self.zonal_field)):
                    if self.filter_z_mode > self.low_pass_z or (self.filter_z_mode == 0 and not zonal):
                        keep[i] = 0.0
            fields = dict(m)
            fields.update(eq)
            if d2J is not None:
                fields["d2Jpar"] = d2J
            else:
                fields["eta"] = self.eta2d
            fields.update(U=U, P=P, Psi=Psi, phi=self.phi, Jpar=self.Jpar, Psi_up=sPsi[0][0],
                          Psi_dn=sPsi[1][0], phi_up=sphi[0][0], phi_dn=sphi[1][0],
                          ddt_U=self.ddt_U, ddt_P=self.ddt_P, ddt_Psi=self.ddt_Psi)
            fields["rowc_ws"] = buf["rowc_ws"]
            if modes:
                fields.update(U_mode=buf["mU"], P_mode=buf["mP"], J_mode=buf["mJ"], zShift=self.zs)
            else:
                fields.update(U_up=sU[0][0], U_dn=sU[1][0], P_up=sP[0][0], P_dn=sP[1][0], J_up=sJ[0][0], J_dn=sJ[1][0])
            if self.kernel_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            capi.elmpb_rhs(g, fields, self.hyperresist, self.filter_z_mode, keep=tuple(keep), rk=rk, zlength=self.zl)
            rk = None
            if self.kernel_events is not None:
                e1.record()
                self.kernel_events.append((e0, e1))
        else:
            self._rhs_terms(U, P, Psi, sU, sP, sPsi, sphi, sJ, d2J)
        if rk is not None:   # operator path: the stage update is its own pass
            for st0, st1, k in zip(rk[0], rk[1], (self.ddt_U, self.ddt_P, self.ddt_Psi)):
                capi.rk_stage(g, "axpy", rk[2], st0, k, out=st1)

    def _rhs_terms(self, U, P, Psi, sU, sP, sPsi, sphi, sJ, d2J):
        """Operator-by-operator evaluation of the three equations (exact also without filter_z: the terms the
        reference takes through field-aligned coordinates are taken through them here too)."""
        g, m, eq = self.g, self.md, self.eq
        ew = capi.elementwise
        # ---- ddt(Psi) = -Grad_par(phi B0)/B0 + eta Jpar - bracket(phi0, Psi) B0 - eta hyperresist Delp2(Jpar)
        # bracket(Field2D, Field3D, BRACKET_STD) = b0xGrad_dot_Grad / Bxy (difops.cxx:805-808)
        phiB0 = ew(g, [(1.0, self.phi, self.B0)])
        gp = self._grad_par_noslices(phiB0)
        bg = capi.b0xgrad_dot_grad(g, self.b0x, eq["phi0"], Psi, None, sPsi)
        ew(g, [(-1.0, gp, None, self.B0), (1.0, eq["eta"], self.Jpar), (-1.0, bg, self.B0, m["Bxy"]),
               (-self.hyperresist, eq["eta"], d2J)], out=self.ddt_Psi, region="RGN_NOBNDRY")
        # ---- ddt(U) = B0^2 b0xGrad_dot_Grad(Psi_loc, J0) + b0xcv . Grad(P) - B0^2 Grad_par(Jpar)
        #               - b0xGrad_dot_Grad(phi0, U)                                         :1755-1775
        # Psi_loc = Psi + the same zerolaplace boundary again (bitwise Psi); it keeps Psi's slices
        t1 = capi.b0xgrad_dot_grad(g, self.b0x, Psi, eq["J0"], sPsi, None, shifted=(self.zs, self.zl))
        dPx = capi.stencil(g, "DDX_C4", "X", P, scale1=self.rdx)
        dPy = capi.stencil_yslices(g, "DDX_C2", P, sP[0], sP[1], scale1=self.rdy)
        dPz = capi.stencil(g, "DDX_C4", "Z", P, scale1=self.rdz)
        gpJ = capi.stencil_yslices(g, "DDX_C2", self.Jpar, sJ[0], sJ[1], scale1=self.gradpar)
        t4 = capi.b0xgrad_dot_grad(g, self.b0x, eq["phi0"], U, None, sU)
        ew(g, [(1.0, t1, self.B0sq), (1.0, dPx, eq["b0xcv_x"]), (1.0, dPy, eq["b0xcv_y"]),
               (1.0, dPz, eq["b0xcv_z"]), (-1.0, gpJ, self.B0sq), (-1.0, t4)], out=self.ddt_U,
           region="RGN_NOBNDRY")
        # ---- ddt(P) = -b0xGrad_dot_Grad(phi, P0) - b0xGrad_dot_Grad(phi0, P)                :1869-1876
        p1 = capi.b0xgrad_dot_grad(g, self.b0x, self.phi, eq["P0"], sphi, None, shifted=(self.zs, self.zl))
        p2 = capi.b0xgrad_dot_grad(g, self.b0x, eq["phi0"], P, None, sP)
        ew(g, [(-1.0, p1), (-1.0, p2)], out=self.ddt_P, region="RGN_NOBNDRY")
        # ---- filter_z, low_pass_z in one pass per field                                      :1931-1955
        if self.filter_z_mode < 0:
            return
        for f, zonal in ((self.ddt_Psi, self.zonal_field), (self.ddt_U, self.zonal_flow),
                         (self.ddt_P, self.zonal_bkgd)):
            if self.low_pass_z > 0:
                capi.zfilter_lowpass(g, f, self.filter_z_mode, self.low_pass_z, keep_zonal=zonal, out=f)
            else:
                capi.zfilter(g, f, self.filter_z_mode, out=f)

    def rk4_step(self, dt):
        """RK4Solver::take_step (rk4.cxx:172-210): stage combinations over the evolving region.  Each RHS
        writes its time derivatives straight into that stage's k arrays and the stage states are built from
        the interior only (their guard cells are filled by the next pre_rhs + communicate), so a step is
        4 RHS (the stage states are written by the fused RHS kernel's epilogue) + the final combination, with no
        gather/scatter, copy or axpy passes."""
        g = self.g
        state = (self.U, self.P, self.Psi)
        if not hasattr(self, "_rk"):
            self._rk = [[torch.zeros_like(self.U) for _ in range(3)] for _ in range(5)]   # k1..k4, stage state
        k1, k2, k3, k4, st = self._rk
        keep = (self.ddt_U, self.ddt_P, self.ddt_Psi)

        if not hasattr(self, "_rk2"):
            self._rk2 = [torch.zeros_like(self.U) for _ in range(3)]   # second stage state (ping-pong)
        su = self._rk2

        def stage(src, k, coef, dst):
            # the stage state a kernel writes must not be the one it reads neighbours of: ping-pong st / su
            self.ddt_U, self.ddt_P, self.ddt_Psi = k
            self.rhs(*src, rk=None if coef is None else (state, dst, coef * dt))
        try:
            stage(state, k1, 0.5, st)
            stage(st, k2, 0.5, su)
            stage(su, k3, 1.0, st)
            stage(st, k4, None, None)
        finally:
            self.ddt_U, self.ddt_P, self.ddt_Psi = keep
        for i in range(3):
            capi.rk_stage(g, "rk4_final", dt, state[i], k1[i], k2[i], k3[i], k4[i], out=state[i])

def elmpb_initial_state(g: capi.Grid, device="cuda", amp=1e-5):
    """Synthetic smooth (U, P, Psi) of elm-pb magnitudes: a gaussian-in-x, ballooning-like-in-y envelope times
    the fundamental z harmonic plus weaker higher harmonics (the shipped [U] function is a ballooning transform
    of gauss(x-0.5,0.1) gauss(y-pi,0.6pi) sin(z); examples/elm-pb/data/BOUT.inp)."""
    x = ((torch.arange(g.nx, dtype=torch.float64, device=device) - g.xstart + 0.5) / (g.nx - 2 * g.xstart))[:, None, None]
    y = (2 * torch.pi * (torch.arange(g.ny, dtype=torch.float64, device=device) - g.ystart + 0.5)
         / (g.ny - 2 * g.ystart))[None, :, None]
    z = (2 * torch.pi * torch.arange(g.nz, dtype=torch.float64, device=device) / g.nz)[None, None, :]
    env = torch.exp(-((x - 0.5) / 0.1) ** 2) * torch.exp(-((y - torch.pi) / (0.6 * torch.pi)) ** 2)
    U = amp * env * (torch.sin(z) + 0.2 * torch.cos(3 * z + y))
    P = 0.3 * amp * env * (torch.cos(z) + 0.1 * torch.sin(2 * z - y))
    Psi = 0.04 * amp * env * (torch.sin(z + 0.5) + 0.1 * torch.cos(4 * z))
    return U.contiguous(), P.contiguous(), Psi.contiguous()


class Blob2D:
    """examples/blob2d/blob2d.cxx rhs() (lines 165-213) with the shipped switches (boussinesq, sheath, not compressible;
    [all] bndry_all = neumann, mesh:ddz first = C2, Delp2 through FFTs, LaplaceCyclic with the [phiBoussinesq] flags),
    on a BoutMesh decomposed in X (BASELINE configs[2]): X guard planes of n, omega and phi travel through
    Mesh::communicate, the tridiagonal solves through reduce -> all-gather of the interface rows -> finish
    (CyclicReduce across NXPE ranks, cyclic_reduction.hxx:215-440)."""

    def __init__(self, g: capi.Grid, metric: dict, Te0=5.0, B0=0.35, D_n=1e-6, D_vort=1e-6, R_c=1.5, L_par=10.0,
                 e=1.602e-19, m_i=2 * 1.667e-27, flags=(0, 0, 0), device="cuda", mesh=None):
        self.g, self.dev, self.mesh = g, torch.device(device), mesh
        self.md = {k: v.to(self.dev) for k, v in metric.items()}
        Omega_i = e * B0 / m_i                     # blob2d.cxx:80-82
        c_s = (e * Te0 / m_i) ** 0.5
        self.rho_s = c_s / Omega_i
        self.D_n, self.D_vort, self.R_c, self.L_par = D_n, D_vort, R_c, L_par
        self.first_x = mesh is None or mesh.first_x
        self.last_x = mesh is None or mesh.last_x
        self.x_split = mesh is not None and mesh.nxpe > 1
        zl = float(self.md["dz"][0, 0].item()) * g.nz
        self.lap = capi.Laplace(g, self.md, zl, all_terms=True, nonuniform=True, flags=flags,
                                first_x=self.first_x, last_x=self.last_x)
        self.lap.workspace(self.dev)
        self.delp = capi.Laplace(g, self.md, zl, all_terms=True, nonuniform=True)   # Delp2's default Laplacian
        self.rdz = (1.0 / self.md["dz"]).contiguous()
        if self.x_split:
            nif = self.lap.iface_doubles()
            self.iface = torch.empty(nif, dtype=torch.float64, device=self.dev)
            self.iface_all = torch.empty(nif * mesh.nxpe, dtype=torch.float64, device=self.dev)
        z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device=self.dev)
        self.n, self.omega, self.phi = z(), z(), z()
        self.ddt_n, self.ddt_omega = z(), z()
        self._t = [z() for _ in range(5)]      # bracket x 2, DDZ(n), Delp2 x 2

    def set_state(self, n, omega):
        self.n.copy_(n)
        self.omega.copy_(omega)

    def _bc(self, fields):                     # bndry_all = neumann (Solver::pre_rhs), one launch
        e = []
        for f in fields:
            if self.first_x:
                e.append(("neumann", "xin", f, 0.0, self.md["dx"]))
            if self.last_x:
                e.append(("neumann", "xout", f, 0.0, self.md["dx"]))
        if e:
            capi.boundary_batch(self.g, e)

    def _comm(self, fields):
        if self.mesh is not None:
            self.mesh.communicate(fields)      # ny = 1, MYG = 0: X guard planes only

    def rhs(self, n=None, omega=None):
        g = self.g
        n = self.n if n is None else n
        omega = self.omega if omega is None else omega
        ew = capi.elementwise
        self._bc([n, omega])
        self._comm([n, omega])                                            # :169
        if not self.x_split:                                              # phi = phiSolver->solve(omega)  :181
            self.lap.solve(omega, out=self.phi)
        else:
            self.lap.reduce(omega, iface=self.iface)
            self.mesh.all_gather_x(self.iface, self.iface_all)
            self.lap.finish(self.iface_all, self.mesh.nxpe, self.mesh.pe_xind, self.phi)
        self._comm([self.phi])                                            # :184
        bn, bo, dzn, d2n, d2o = self._t
        dx, dz = self.md["dx"], self.md["dz"]
        capi.bracket_arakawa(g, self.phi, n, dx, dz, out=bn)
        capi.bracket_arakawa(g, self.phi, omega, dx, dz, out=bo)
        capi.stencil(g, "DDX_C2", "Z", n, scale1=self.rdz, out=dzn)
        self.delp.delp2(n, out=d2n)
        self.delp.delp2(omega, out=d2o)
        curv, sh = 2.0 * self.rho_s / self.R_c, self.rho_s / self.L_par
        # ddt(n) = -bracket(phi, n) + 2 DDZ(n) rho_s/R_c + D_n Delp2(n) + n phi rho_s/L_par          :189-200
        ew(g, [(-1.0, bn), (curv, dzn), (self.D_n, d2n), (sh, n, self.phi)], out=self.ddt_n, region="RGN_NOBNDRY")
        # ddt(omega) = -bracket(phi, omega) + 2 DDZ(n) rho_s/R_c / n + D_vort Delp2(omega) / n + phi rho_s/L_par
        ew(g, [(-1.0, bo), (curv, dzn, None, n), (self.D_vort, d2o, None, n), (sh, self.phi)], out=self.ddt_omega,
           region="RGN_NOBNDRY")

    def rk4_step(self, dt):
        """RK4Solver::take_step (rk4.cxx:172-210)."""
        g = self.g
        if not hasattr(self, "_rk"):
            self._rk = [[torch.zeros_like(self.n) for _ in range(2)] for _ in range(5)]   # k1..k4, stage state
        k1, k2, k3, k4, st = self._rk
        state = (self.n, self.omega)
        keep = (self.ddt_n, self.ddt_omega)
        try:
            for src, k, coef in ((state, k1, 0.5), (st, k2, 0.5), (st, k3, 1.0), (st, k4, None)):
                self.ddt_n, self.ddt_omega = k
                self.rhs(*src)
                if coef is not None:
                    for i in range(2):
                        capi.rk_stage(g, "axpy", coef * dt, state[i], k[i], out=st[i])
        finally:
            self.ddt_n, self.ddt_omega = keep
        for i in range(2):
            capi.rk_stage(g, "rk4_final", dt, state[i], k1[i], k2[i], k3[i], k4[i], out=state[i])


def blob2d_initial_state(g: capi.Grid, mesh=None, device="cuda", height=0.5, width=0.05, gnx=None):
    """[n] function = 1 + height exp(-((x-0.25)/width)^2 - ((z/(2 pi) - 0.5)/width)^2), omega = 0
    (examples/blob2d/delta_1/BOUT.inp); x = GlobalX of the decomposed mesh."""
    jx = torch.arange(g.nx, dtype=torch.float64, device=device)
    if mesh is not None:
        jx = jx + mesh.pe_xind * mesh.mxsub
        gnx = mesh.nxpe * mesh.mxsub + 2 * mesh.mxg
    gnx = g.nx if gnx is None else gnx
    x = ((jx - g.xstart + 0.5) / (gnx - 2 * g.xstart))[:, None, None]
    z = (torch.arange(g.nz, dtype=torch.float64, device=device) / g.nz)[None, None, :]
    n = 1.0 + height * torch.exp(-((x - 0.25) / width) ** 2 - ((z - 0.5) / width) ** 2)
    n = n.expand(g.nx, g.ny, g.nz).contiguous()
    return n, torch.zeros_like(n)
