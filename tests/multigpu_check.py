"""Run under torchrun on N GPUs (N = 2, 4, 8): X-decomposed HW3D against the single-rank golden
trajectory of the reference (tests/golden/models.npz), an X-decomposed LaplaceCyclic against the
golden operator fixture, and the Y-decomposed elm-pb model against the reference's single-rank RHS and
trajectory (tests/golden/elmpb.npz).  Exit code 0 = parity.  Launched by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from util import rel_err  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    capi, models, meshmod = pkg.capi, pkg.models, pkg.mesh
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "models.npz")))
    gnx, gny, nz, mxg, myg = [int(v) for v in gold["hw3d_grid"]]
    MX = gnx - 2 * mxg
    assert MX % world == 0
    mxsub = MX // world
    mesh = meshmod.BoutMesh(world, 1, rank, mxsub, gny - 2 * myg, nz, mxg, myg, periodic_x=False, periodic_y=True,
                            device=dev)
    sl = slice(rank * mxsub, rank * mxsub + mxsub + 2 * mxg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev)
    metric = {n: t(gold[f"hw3d_metric_{n}"]) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13", "G1", "G3",
                                                        "d1_dx", "g_11")}
    m = models.HW3D(mesh.grid, metric, alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3, device=dev, mesh=mesh)
    m.set_state(t(gold["hw3d_init_n"]), t(gold["hw3d_init_vort"]))
    m.rhs()
    g = mesh.grid
    I = g.interior
    errs = [rel_err(m.ddt_n.cpu().numpy(), gold["hw3d_ddt_n"][sl], I), rel_err(m.ddt_vort.cpu().numpy(), gold["hw3d_ddt_vort"][sl], I)]
    ok = max(errs) < 1e-10
    dt, nsteps = gold["hw3d_dt_nsteps"]
    m.set_state(t(gold["hw3d_init_n"]), t(gold["hw3d_init_vort"]))
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    errs += [rel_err(m.n.cpu().numpy(), gold["hw3d_final_n"][sl], I), rel_err(m.vort.cpu().numpy(), gold["hw3d_final_vort"][sl], I)]
    ok = ok and max(errs[2:]) < 1e-8
    print(f"rank {rank}/{world}: ddt err {errs[0]:.2e} {errs[1]:.2e}; trajectory err {errs[2]:.2e} {errs[3]:.2e}", flush=True)
    ok = hw3d_y_check(pkg, dev, rank, world, gold) and ok
    ok = hw3d_ypeer_check(pkg, dev, rank, world) and ok
    ok = elmpb_check(pkg, dev, rank, world) and ok
    ok = blob2d_x_check(pkg, dev, rank, world, gold) and ok
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


def hw3d_ypeer_check(pkg, dev, rank, world):
    """Y guard rows delivered by the producing kernels over peer memory (bout_b200_ypeer: fused RHS epilogue + the
    Laplacian solve write the neighbours' guard rows, edge CTAs wait on flags) against the exchange kernels
    (BOUT_B200_YPEER=0) on the same NYPE = world decomposition: bitwise, after 3 RK4 steps and one more RHS."""
    capi, models, meshmod = pkg.capi, pkg.models, pkg.mesh
    mxsub, mysub, nz, mxg, myg = 24, 6, 128, 2, 2
    res = {}
    for mode in ("0", "1"):
        os.environ["BOUT_B200_YPEER"] = mode
        mesh = meshmod.BoutMesh(1, world, rank, mxsub, mysub, nz, mxg, myg, periodic_x=False, periodic_y=True, device=dev)
        g = mesh.grid
        one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device=dev)
        md = dict(dx=one(0.3), dy=one(0.7), dz=one(0.2), J=one(1.0), g_22=one(1.0), g11=one(1.0), g33=one(1.0),
                  g13=one(0.0), G1=one(0.0), G3=one(0.0), d1_dx=one(0.0), g_11=one(1.0))
        gen = torch.Generator(device=dev).manual_seed(11 + rank)
        n0 = torch.rand(g.shape3, dtype=torch.float64, device=dev, generator=gen) - 0.5
        v0 = torch.rand(g.shape3, dtype=torch.float64, device=dev, generator=gen) - 0.5
        m = models.HW3D(g, md, alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3, device=dev, mesh=mesh)
        assert m.ypeer == (mode == "1"), (mode, m.ypeer)
        m.set_state(n0, v0)
        for _ in range(3):
            m.rk4_step(1e-3)
        m.rhs()
        torch.cuda.synchronize()
        dist.barrier()
        I = g.interior
        res[mode] = [t[I].clone() for t in (m.n, m.vort, m.phi, m.ddt_n, m.ddt_vort)]
        if mode == "1":
            assert not m.arena.timed_out()
    os.environ.pop("BOUT_B200_YPEER", None)
    same = all(torch.equal(a, b) for a, b in zip(res["0"], res["1"]))
    big = float(res["1"][0].abs().max())
    print(f"rank {rank}/{world} (ypeer vs exchange kernels, NYPE={world}): bitwise equal {same}, max |n| {big:.3f}", flush=True)
    return same and big > 0.1 and all(bool(torch.isfinite(t).all()) for t in res["1"])


def blob2d_x_check(pkg, dev, rank, world, gold):
    """BASELINE configs[2]: models.Blob2D on NXPE = N ranks (X guard planes through Mesh::communicate, the tridiagonal
    solves through reduce -> all-gather -> finish) against the reference's single-rank run of the unchanged
    examples/blob2d source: one RHS (1e-10) and the 5-step RK4 trajectory (1e-8)."""
    capi, models, meshmod = pkg.capi, pkg.models, pkg.mesh
    gnx, gny, nz, mxg, myg = [int(v) for v in gold["blob2d_grid"]]
    MX = gnx - 2 * mxg
    if MX % world != 0 or MX // world < mxg:
        return True
    mxsub = MX // world
    mesh = meshmod.BoutMesh(world, 1, rank, mxsub, gny - 2 * myg, nz, mxg, myg, periodic_x=False, periodic_y=True, device=dev)
    sl = slice(rank * mxsub, rank * mxsub + mxsub + 2 * mxg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev)
    metric = {n: t(gold[f"blob2d_metric_{n}"]) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13", "G1", "G3",
                                                          "d1_dx", "g_11")}
    m = models.Blob2D(mesh.grid, metric, Te0=5.0, B0=0.35, D_n=1e-6, D_vort=1e-6, R_c=1.5, L_par=10.0, device=dev, mesh=mesh)
    I = mesh.grid.interior
    n0, _ = models.blob2d_initial_state(mesh.grid, mesh=mesh, device=dev, height=0.5, width=0.1)
    e0 = rel_err(n0.cpu().numpy(), gold["blob2d_init_n"][sl], I)
    m.set_state(t(gold["blob2d_init_n"]), t(gold["blob2d_init_omega"]))
    m.rhs()
    errs = [rel_err(m.ddt_n.cpu().numpy(), gold["blob2d_ddt_n"][sl], I), rel_err(m.ddt_omega.cpu().numpy(), gold["blob2d_ddt_omega"][sl], I)]
    dt, nsteps = gold["blob2d_dt_nsteps"]
    m.set_state(t(gold["blob2d_init_n"]), t(gold["blob2d_init_omega"]))
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    errs += [rel_err(m.n.cpu().numpy(), gold["blob2d_final_n"][sl], I), rel_err(m.omega.cpu().numpy(), gold["blob2d_final_omega"][sl], I)]
    ok = e0 < 1e-13 and max(errs[:2]) < 1e-10 and max(errs[2:]) < 1e-8
    print(f"rank {rank}/{world}: blob2d NXPE={world}: init {e0:.1e}; ddt err {errs[0]:.2e} {errs[1]:.2e}; trajectory err "
          f"{errs[2]:.2e} {errs[3]:.2e}", flush=True)
    return ok


def hw3d_y_check(pkg, dev, rank, world, gold):
    """HW3D split along Y (NXPE = 1, NYPE = world: bench.py's default decomposition, with the n / vort guard-row
    exchange overlapped with the Laplacian solve) vs the reference's single-rank golden trajectory."""
    gnx, gny, nz, mxg, myg = [int(v) for v in gold["hw3d_grid"]]
    my = gny - 2 * myg
    if my % world != 0 or my // world < myg:
        print(f"rank {rank}: hw3d fixture (MY = {my}) does not split over {world} ranks; skipped", flush=True)
        return True
    mysub = my // world
    mesh = pkg.mesh.BoutMesh(1, world, rank, gnx - 2 * mxg, mysub, nz, mxg, myg, periodic_x=False, periodic_y=True,
                             device=dev)
    ysl = slice(rank * mysub, rank * mysub + mysub + 2 * myg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, ysl])).to(dev)
    metric = {n: t(gold[f"hw3d_metric_{n}"]) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13", "G1", "G3",
                                                        "d1_dx", "g_11")}
    m = pkg.models.HW3D(mesh.grid, metric, alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3, device=dev, mesh=mesh)
    assert m.overlap
    m.set_state(t(gold["hw3d_init_n"]), t(gold["hw3d_init_vort"]))
    dt, nsteps = gold["hw3d_dt_nsteps"]
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    I = mesh.grid.interior
    e = [rel_err(m.n.cpu().numpy(), gold["hw3d_final_n"][:, ysl], I),
         rel_err(m.vort.cpu().numpy(), gold["hw3d_final_vort"][:, ysl], I)]
    print(f"rank {rank}/{world}: hw3d NYPE={world} trajectory err {e[0]:.2e} {e[1]:.2e} "
          f"(halo: {'peer memory' if mesh._peer else 'nccl'}, timed out: {mesh.halo_timed_out()})", flush=True)
    return max(e) < 1e-8 and not mesh.halo_timed_out()


def elmpb_check(pkg, dev, rank, world):
    """elm-pb split along Y over all ranks (NXPE = 1, NYPE = world) vs the reference's single-rank golden."""
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "elmpb.npz")))
    nx, ny, nz, mxg, myg = [int(v) for v in gold["grid"]]
    my = ny - 2 * myg
    if my % world != 0 or my // world < myg:
        print(f"rank {rank}: elm-pb fixture (MY = {my}) does not split over {world} ranks; skipped", flush=True)
        return True
    mysub = my // world
    mesh = pkg.mesh.BoutMesh(1, world, rank, nx - 2 * mxg, mysub, nz, mxg, myg, periodic_x=False, periodic_y=True,
                             device=dev)
    m = pkg.models.ELMpb.from_setup(pkg.elm_setup.elmpb_setup(nx, my, nz, mxg, myg), device=dev, mesh=mesh)
    ysl = slice(rank * mysub, rank * mysub + mysub + 2 * myg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, ysl])).to(dev)
    state = [t(gold["rhs_in_" + n]) for n in ("U", "P", "Psi")]
    I = mesh.grid.interior
    m.set_state(*state)
    m.rhs()
    e1 = [rel_err(f.cpu().numpy(), gold["rhs_ddt_" + n][:, ysl], I)
          for n, f in (("U", m.ddt_U), ("P", m.ddt_P), ("Psi", m.ddt_Psi))]
    m.set_state(*state)
    dt, nsteps = gold["dt_nsteps"]
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    e2 = [rel_err(f.cpu().numpy(), gold["final_" + n][:, ysl], I) for n, f in (("U", m.U), ("P", m.P), ("Psi", m.Psi))]
    print(f"rank {rank}/{world}: elm-pb NYPE={world} ddt err {max(e1):.2e}; trajectory err {max(e2):.2e} "
          f"(halo: {'peer memory' if mesh._peer else 'nccl'}, timed out: {mesh.halo_timed_out()})", flush=True)
    return max(e1) < 1e-10 and max(e2) < 1e-8 and not mesh.halo_timed_out()


if __name__ == "__main__":
    main()
