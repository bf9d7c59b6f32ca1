"""The C++ host layer's Mesh decomposed in y over NYPE ranks (bout::b200::YDecomposition; the reference's BoutMesh with
NXPE = 1, src/mesh/impls/bout/boutmesh.cxx:255-341, sendY / wait :1245-1555): one process per rank, launched by
torch.distributed.run or by tests/test_host_integrators.py with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT set.

Every rank runs the UNCHANGED examples/hasegawa-wakatani-3d/hw.cxx (the headline workload) and examples/conduction/
conduction.cxx (dy and g_22 vary along y: the metric's guard rows and d1_dy come from the neighbouring rank) on its
slab and compares with the slab of the reference build's SINGLE-rank run (tests/golden/models.npz): initial profile
(global y index in the expression evaluator), metric with its guard rows, one Solver::run_rhs (1e-10), the 5-step RK4
trajectory (1e-8) -- results that must not depend on the decomposition.  Then an adaptive integrator (rkgeneric /
cashkarp): the error norm is reduced over the ranks, so every rank must take the same number of RHS evaluations as
the NYPE = 1 run each process makes first (same host layer, same device) and end on its slab of that run.

On a GPU box: NCCL, real kernels.  On CPU: gloo and the host-memory test double (LD_PRELOAD=<libfake_b200.so>
B200_FAKE_ORACLE_PORT=<liboracle_port.so>), which pins the HOST logic.  Rank 0 prints one JSON line; exit code 0 iff ok."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))


def rel(a, b, whole=None):
    """max |a - b| over the slab, relative to the largest value of the WHOLE reference array (a slab far from the blob
    holds only round-off-sized numbers)"""
    return float(np.abs(a - b).max() / max(np.abs(b if whole is None else whole).max(), 1e-300))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import cases
    hostmod = ge.load_package().host
    rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl" if cuda else "gloo", rank=rank, world_size=size)
    gold = np.load(os.path.join(HERE, "golden", "models.npz"))
    res, bad = {"ranks": size}, []
    direction = sys.argv[1] if len(sys.argv) > 1 else "y"
    res["direction"] = direction
    yr = hostmod.YRanks(1 << 16, direction=direction, nxpe=2 if direction == "xy" else None)
    res["nxpe"], res["nype"] = yr.nxpe, yr.nype
    todo = (("hw3d", "hw3d", cases.HW3D_INI), ("conduction", "conduction", cases.CONDUCTION_INI)) if direction == "y" else \
        (("hw3d", "hw3d", cases.HW3D_INI), ("blob2d", "blob2d", cases.BLOB_INI))   # (hw is periodic in x: NXPE > 1 refuses it)

    def slabs(nx, ny, mxg, myg):
        """(guards-included slab of a global array, interior of a local array) for this rank, or None"""
        sx, sy = (nx - 2 * mxg) // yr.nxpe, (ny - 2 * myg) // yr.nype
        if (yr.nxpe > 1 and (mxg == 0 or sx * yr.nxpe != nx - 2 * mxg or sx < mxg)) \
                or (yr.nype > 1 and (myg == 0 or sy * yr.nype != ny - 2 * myg or sy < myg)):
            return None
        x0, y0 = yr.xind * sx, yr.yind * sy
        slab = lambda a: a[x0:x0 + sx + 2 * mxg, y0:y0 + sy + 2 * myg]
        inner = lambda a: a[mxg:mxg + sx, myg:myg + sy]
        return slab, inner, (sx + 2 * mxg, sy + 2 * myg)

    for tag, model, ini in todo:
        nx, ny, nz, mxg, myg = [int(v) for v in gold[tag + "_grid"]]
        sl = slabs(nx, ny, mxg, myg)
        if sl is None:
            res[tag] = "skipped: does not decompose"
            continue
        slab, inner, shape2 = sl
        shape = shape2 + (nz,)
        h = hostmod.Host(ini, yranks=yr)
        try:
            names = h.model_create(model, "rk4")
            assert h.shape3 == shape, h.shape3
            for v in names:
                e = rel(inner(h.model_get(v)), inner(slab(gold[f"{tag}_init_{v}"])), gold[f"{tag}_init_{v}"])
                res[f"{tag}_init_{v}"] = e
                if e > 1e-13:
                    bad.append(f"{tag} init {v}")
            for n in ("dx", "dy", "g11", "g33", "J", "g_22"):
                # guard cells included: between ranks they come from the neighbour (Mesh::exchangeY2D / exchangeX2D)
                if not np.allclose(h.metric(n), slab(gold[f"{tag}_metric_{n}"]), rtol=1e-14, atol=0):
                    bad.append(f"{tag} metric {n}")
            x0, g0 = yr.exchanges, yr.gathers
            h.model_rhs(0.0)
            res[f"{tag}_exchanges_per_rhs"] = yr.exchanges - x0
            res[f"{tag}_gathers_per_rhs"] = yr.gathers - g0
            for v in names:
                e = rel(inner(h.model_ddt(v)), inner(slab(gold[f"{tag}_ddt_{v}"])), gold[f"{tag}_ddt_{v}"])
                res[f"{tag}_ddt_{v}"] = e
                if e > 1e-10:
                    bad.append(f"{tag} ddt {v}")
            dt, nsteps = gold[tag + "_dt_nsteps"]
            for v in names:
                h.model_set(v, np.ascontiguousarray(slab(gold[f"{tag}_init_{v}"])))
            h.model_solve(1, float(dt) * int(nsteps))
            for v in names:
                e = rel(inner(h.model_get(v)), inner(slab(gold[f"{tag}_final_{v}"])), gold[f"{tag}_final_{v}"])
                res[f"{tag}_final_{v}"] = e
                if e > 1e-8:
                    bad.append(f"{tag} final {v}")
        finally:
            h.close()
    # averageY on a decomposed mesh (partial sums all-reduced over the Y column): tests/models/avg_ops.cxx against the
    # reference build's single-rank run
    if yr.nype > 1:
        ops = dict(np.load(os.path.join(HERE, "golden", "vec_ops.npz")))
        ops.update(np.load(os.path.join(HERE, "golden", "avg_ops.npz")))
        nx, ny, nz = ops["f0"].shape
        sl = slabs(nx, ny, 2, 2)
        if sl is not None:
            slab, inner, _ = sl
            h = hostmod.Host(cases.OPS_MESH_INI, yranks=yr)
            try:
                h.model_create("avg_ops", "rk4")
                h.model_set("f", np.ascontiguousarray(slab(ops["f0"])))
                h.model_set("g", np.ascontiguousarray(slab(ops["g0"])))
                h.model_rhs(0.0)
                for v in ("f", "g"):
                    e = rel(inner(h.model_ddt(v)), inner(slab(ops["avg_ddt_" + v])), ops["avg_ddt_" + v])
                    res[f"average_y_ddt_{v}"] = e
                    if e > 1e-12:
                        bad.append(f"averageY ddt {v}")
            finally:
                h.close()
    # adaptive integrator: global error norm.  Single-slab run first (every process makes it), then the decomposed one.
    tag, ini = "hw3d", cases.HW3D_INI.replace("[solver]", "[solver]\nadaptive = true\nscheme = cashkarp\nrtol = 1e-6\natol = 1e-9")
    nx, ny, nz, mxg, myg = [int(v) for v in gold[tag + "_grid"]]
    slab, inner, _ = slabs(nx, ny, mxg, myg)
    mine = lambda a: inner(slab(a))
    h = hostmod.Host(ini)
    try:
        names = h.model_create("hw3d", "rkgeneric")
        h.model_solve(1, 0.05)
        single = {v: h.model_get(v) for v in names}
        single_calls = h.rhs_calls()
    finally:
        h.close()
    h = hostmod.Host(ini, yranks=yr)
    try:
        names = h.model_create("hw3d", "rkgeneric")
        h.model_solve(1, 0.05)
        calls = torch.tensor([h.rhs_calls()], dtype=torch.int64, device="cuda" if cuda else "cpu")
        allc = [torch.zeros_like(calls) for _ in range(size)]
        dist.all_gather(allc, calls)
        res["adaptive_rhs_calls"] = [int(c.item()) for c in allc]
        res["adaptive_rhs_calls_single"] = single_calls
        if any(c != single_calls for c in res["adaptive_rhs_calls"]):
            bad.append("adaptive rhs calls")
        for v in names:
            e = rel(inner(h.model_get(v)), mine(single[v]), single[v])
            res[f"adaptive_final_{v}"] = e
            if e > 1e-10:
                bad.append(f"adaptive final {v}")
    finally:
        h.close()
    # per-rank dump files (the reference's BOUT.dmp.<rank>.nc): stitched the way boutdata.collect does, they must be the
    # records of the single-slab run's file
    import tempfile
    from scipy.io import netcdf_file
    tmp = [tempfile.mkdtemp(prefix="b200_dump_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    work = tmp[0]
    out_ini = lambda path: cases.HW3D_INI + f"\n[output]\npath = {path}\n"
    nx, ny, nz, mxg, myg = [int(v) for v in gold["hw3d_grid"]]
    if rank == 0:
        h = hostmod.Host(out_ini(os.path.join(work, "single.nc")))
        try:
            h.model_create("hw3d", "rk4")
            h.model_solve(2, 0.02)
        finally:
            h.close()
    h = hostmod.Host(out_ini(os.path.join(work, "BOUT.dmp.0.nc")), yranks=yr)
    try:
        h.model_create("hw3d", "rk4")
        h.model_solve(2, 0.02)
    finally:
        h.close()
    # restart files per rank: one output, then a restarted run of one more, ends bit-equal to the straight 2-output run
    straight = None
    for k, (nout, extra) in enumerate(((2, ""), (1, ""), (1, "restart = true\n"))):
        rp = os.path.join(work, "BOUT.restart.nc" if k else "unused.restart.nc")
        h = hostmod.Host(extra + cases.HW3D_INI + f"\n[restart_file]\npath = {rp}\n", yranks=yr)
        try:
            names = h.model_create("hw3d", "rk4")
            h.model_solve(nout, 0.02)
            final = {v: h.model_get(v) for v in names}
        finally:
            h.close()
        straight = final if k == 0 else straight
    if not os.path.exists(os.path.join(work, f"BOUT.restart.{rank}.nc")):
        bad.append("restart file name")
    for v in final:
        res[f"restart_{v}_bit_equal"] = bool(np.array_equal(inner(final[v]), inner(straight[v])))
        if not res[f"restart_{v}_bit_equal"]:
            bad.append(f"restart {v}")
    dist.barrier()
    if rank == 0:
        with netcdf_file(os.path.join(work, "single.nc"), "r", mmap=False) as nc:
            want = {v: np.array(nc.variables[v][:]) for v in ("n", "vort")}
            tw = np.array(nc.variables["t_array"][:])
        got = {v: np.full_like(want[v], np.nan) for v in want}
        for r in range(size):
            with netcdf_file(os.path.join(work, f"BOUT.dmp.{r}.nc"), "r", mmap=False) as nc:
                sc = {k: int(nc.variables[k].getValue()) for k in ("NXPE", "NYPE", "MXSUB", "MYSUB", "MXG", "MYG", "nx", "ny",
                                                                   "PE_XIND", "PE_YIND")}
                if (sc["NXPE"], sc["NYPE"], sc["nx"], sc["ny"]) != (yr.nxpe, yr.nype, nx, ny - 2 * myg) \
                        or r != sc["PE_YIND"] * sc["NXPE"] + sc["PE_XIND"] or not np.array_equal(nc.variables["t_array"][:], tw):
                    bad.append(f"dump scalars rank {r}")
                x0, y0 = sc["PE_XIND"] * sc["MXSUB"], sc["PE_YIND"] * sc["MYSUB"]
                for v in want:
                    a = np.array(nc.variables[v][:])
                    # boutdata.collect: interior cells from their owner, x boundary cells from the end ranks
                    xl = 0 if sc["PE_XIND"] == 0 else mxg
                    xh = a.shape[1] if sc["PE_XIND"] == sc["NXPE"] - 1 else a.shape[1] - mxg
                    got[v][:, x0 + xl:x0 + xh, myg + y0:myg + y0 + sc["MYSUB"]] = a[:, xl:xh, myg:myg + sc["MYSUB"]]
        for v in want:
            I = (slice(None), slice(None), slice(myg, ny - myg))
            e = float(np.nanmax(np.abs(got[v][I] - want[v][I])) / np.abs(want[v][I]).max()) if not np.isnan(got[v][I]).any() else 1.0
            res[f"dump_collect_{v}"] = e
            if e > 1e-12:
                bad.append(f"dump collect {v}")
        import shutil
        shutil.rmtree(work, ignore_errors=True)
    flag = torch.tensor([len(bad)], dtype=torch.int64, device="cuda" if cuda else "cpu")
    dist.all_reduce(flag)
    res.update(ok=int(flag.item()) == 0, bad=bad)
    if rank == 0 or bad:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
