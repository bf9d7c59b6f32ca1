"""Harness-side drivers of the fused model right-hand sides (Python over the C-ABI).

These mirror what the reference does around PhysicsModel::rhs for the explicit RK solvers
(Solver::run_rhs src/solver/solver.cxx:1380-1427: pre_rhs boundary conditions -> model rhs;
RK4Solver::take_step src/solver/impls/rk4/rk4.cxx:172-210), with device-resident SoA fields: the
load_vars / save_derivs gather-scatter passes of the reference do not exist here and the stage
updates are written by the last RHS kernel (bout_b200_rk_epilogue).  Every numerical step is a
C-ABI call; torch only owns the memory.  The C++ host layer (host/) drives the same entry points.
"""
from __future__ import annotations

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
        self.lap = capi.Laplace(g, self.md, zlength, nonuniform=True, first_x=first_x, last_x=last_x)
        self.lap.workspace(self.dev)
        self.x_split = mesh is not None and mesh.nxpe > 1
        if self.x_split:  # interface records of the X-decomposed tridiagonal solve
            nif = self.lap.iface_doubles()
            self.iface = torch.empty(nif, dtype=torch.float64, device=self.dev)
            self.iface_all = torch.empty(nif * mesh.nxpe, dtype=torch.float64, device=self.dev)
        self.coef = capi.fused_coefs(g, self.cm, capi.fused_coef_workspace(g, self.dev))
        self.phase_events = None
        self.kernel_events = None  # bench.py: list collecting (start, end) CUDA events of the fused kernel
        z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device=self.dev)
        self.n, self.vort, self.phi = z(), z(), z()
        self.ddt_n, self.ddt_vort = z(), z()
        # RK work fields: stage state (k5 in rk4.cxx) and the running combination
        self.sn, self.sv, self.tn, self.tv, self.an, self.av = z(), z(), z(), z(), z(), z()

    def set_state(self, n, vort):
        self.n.copy_(n)
        self.vort.copy_(vort)

    # -- one right-hand side on (n, vort) held in fields fn, fv
    def _pre_rhs(self, fn, fv):
        g = self.g
        if self.mesh is None or self.mesh.first_x:
            capi.boundary(g, "dirichlet", "xin", fn)
            capi.boundary(g, "dirichlet", "xin", fv)
        if self.mesh is None or self.mesh.last_x:
            capi.boundary(g, "dirichlet", "xout", fn)
            capi.boundary(g, "dirichlet", "xout", fv)

    def _mark(self, name):
        """bench.py / tools: optional per-phase CUDA events (None = off, no overhead)."""
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append((name, e))

    def rhs(self, fn=None, fv=None, epilogue=None, ddt=True):
        fn = self.n if fn is None else fn
        fv = self.vort if fv is None else fv
        self._mark("start")
        self._pre_rhs(fn, fv)                                   # Solver::pre_rhs
        self._mark("bc")
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
        if self.mesh is None:                                   # mesh->communicate(n, vort, phi, ...)
            capi.halo_self(self.g, [fn, fv, self.phi], periodic_x=False, periodic_y=True)
        else:
            self.mesh.communicate([fn, fv, self.phi])
        self._mark("halo")
        if self.kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        capi.hw3d_rhs(self.g, None, self.alpha, self.kappa, self.Dn, self.Dvort, fn, fv, self.phi,
                      self.ddt_n if ddt else None, self.ddt_vort if ddt else None, epilogue=epilogue,
                      coef_ws=self.coef)
        if self.kernel_events is not None:
            e1.record()
            self.kernel_events.append((e0, e1))
        self._mark("fused")

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
