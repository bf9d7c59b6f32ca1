"""Output step of the C++ host layer (bout-dev_b200/host/src/dump_writer.cxx): runs
tests/models/ode_pair.cxx with `[output] path = <file>` and checks the dump file with scipy's NetCDF reader.

    python tests/output_check.py <workdir>                       real device (GPU test)
    LD_PRELOAD=<libfake_b200.so> python tests/output_check.py …   host-memory test double (CPU: host logic only)

Checks: the file parses as NetCDF classic; dimensions / scalars as in BOUT.dmp.<rank>.nc; t_array and iteration per
record (the first record is the initial state, iteration -1, as Solver::solve's call_monitors(simtime, -1, NOUT));
record k of a 3-output run equals, bit for bit, the final state of a separate k-output run (fixed binary-exact
timestep), i.e. a snapshot is not disturbed by the steps that run while it is copied and written; a run restarted
from the restart file of a 2-output run (`restart = true`, `[restart_file] path`) ends bit-equal to the 3-output run.
Prints one JSON line; exit code 0 iff all checks pass."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))

SOLVER_OPTS = "timestep = 0.0625"
STEP = 0.25


def run(hostmod, nout, path=None, restart_path=None, restart=False):
    from cases import ode_ini
    ini = ("restart = true\n" if restart else "") + ode_ini(SOLVER_OPTS, False, nout, STEP)
    if path is not None:
        ini += f"\n[output]\npath = {path}\n"
    if restart_path is not None:
        ini += f"\n[restart_file]\npath = {restart_path}\n"
    h = hostmod.Host(ini)
    try:
        names = h.model_create("ode_pair", "rk4")
        init = {v: h.model_get(v) for v in names}
        h.model_solve(nout, STEP)
        return init, {v: h.model_get(v) for v in names}, h.info
    finally:
        h.close()


def main(workdir):
    from scipy.io import netcdf_file
    import __graft_entry__ as ge
    hostmod = ge.load_package().host
    path = os.path.join(workdir, "BOUT.dmp.0.nc")
    init, final, info = run(hostmod, 3, path)
    nx, ny, nz, xstart, xend, ystart, yend = info[:7]
    res = {"records": None, "checks": []}

    def check(name, ok):
        res["checks"].append([name, bool(ok)])

    with netcdf_file(path, "r", mmap=False) as nc:
        res["records"] = int(nc.dimensions["t"] is None and nc.variables["t_array"].shape[0])
        check("dims", nc.dimensions["x"] == nx and nc.dimensions["y"] == ny and nc.dimensions["z"] == nz)
        sc = {k: int(nc.variables[k].getValue()) for k in ("NXPE", "NYPE", "MXSUB", "MYSUB", "MXG", "MYG", "MZ", "nx", "ny", "nz")}
        check("scalars", sc == {"NXPE": 1, "NYPE": 1, "MXSUB": xend - xstart + 1, "MYSUB": yend - ystart + 1, "MXG": 2,
                                "MYG": 0, "MZ": nz, "nx": nx, "ny": yend - ystart + 1, "nz": nz})
        check("t_array", np.array_equal(nc.variables["t_array"][:], [0.0, 0.25, 0.5, 0.75]))
        check("iteration", np.array_equal(nc.variables["iteration"][:], [-1, 0, 1, 2]))
        recs = {v: np.array(nc.variables[v][:]) for v in ("f", "g")}
        # SAVE_REPEAT(energy), SAVE_ONCE(damp, answer, f_initial) of tests/models/ode_pair.cxx
        check("once_scalars", float(nc.variables["damp"].getValue()) == 0.3 and int(nc.variables["answer"].getValue()) == 42)
        check("once_field", np.array_equal(np.array(nc.variables["f_initial"][:]), init["f"])
              and nc.variables["f_initial"].dimensions == ("x", "y", "z"))
        en = np.array(nc.variables["energy"][:])
        I = (slice(None), slice(xstart, xend + 1), slice(ystart, yend + 1), slice(None))
        want = 0.5 * (recs["f"] * recs["f"] + recs["g"] * recs["g"])
        check("repeat_field", en.shape == (4, nx, ny, nz) and np.abs(en[I] - want[I]).max() <= 1e-15 * np.abs(want[I]).max())
    for v in ("f", "g"):
        check(f"shape_{v}", recs[v].shape == (4, nx, ny, nz))
        check(f"initial_{v}", np.array_equal(recs[v][0], init[v]))
        check(f"final_{v}", np.array_equal(recs[v][3], final[v]))
    for k in (1, 2):
        _, fin_k, _ = run(hostmod, k)
        for v in ("f", "g"):
            check(f"record{k}_{v}", np.array_equal(recs[v][k], fin_k[v]))
    # restart: two outputs with a restart file, then a restarted run of one more output == the straight 3-output run
    rpath = os.path.join(workdir, "BOUT.restart.0.nc")
    _, fin2, _ = run(hostmod, 2, restart_path=rpath)
    with netcdf_file(rpath, "r", mmap=False) as nc:
        check("restart_file", float(nc.variables["tt"].getValue()) == 0.5 and int(nc.variables["hist_hi"].getValue()) == 1
              and np.array_equal(np.array(nc.variables["f"][:]), fin2["f"]) and np.array_equal(np.array(nc.variables["g"][:]), fin2["g"]))
    path2 = os.path.join(workdir, "BOUT.dmp.restarted.nc")
    _, fin3, _ = run(hostmod, 1, path=path2, restart_path=rpath, restart=True)
    for v in ("f", "g"):
        check(f"restarted_{v}", np.array_equal(fin3[v], final[v]))
    with netcdf_file(path2, "r", mmap=False) as nc:
        check("restarted_times", np.array_equal(nc.variables["t_array"][:], [0.5, 0.75])
              and np.array_equal(nc.variables["iteration"][:], [1, 2]))
    with netcdf_file(rpath, "r", mmap=False) as nc:
        check("restart_file_advanced", float(nc.variables["tt"].getValue()) == 0.75 and int(nc.variables["hist_hi"].getValue()) == 2)
    res["ok"] = all(ok for _, ok in res["checks"])
    print(json.dumps(res), flush=True)
    return 0 if res["ok"] else 1


def main_hw(workdir):
    """Unchanged examples/hasegawa-wakatani/hw.cxx: evolving n, vort plus phi registered by the model's outputVars()."""
    from scipy.io import netcdf_file
    import __graft_entry__ as ge
    from cases import HW_INI
    hostmod = ge.load_package().host
    path = os.path.join(workdir, "BOUT.dmp.hw.nc")
    h = hostmod.Host("nout = 2\ntimestep = 0.05\n" + HW_INI + f"\n[output]\npath = {path}\n")
    try:
        names = h.model_create("hw", "rk4")
        h.model_solve(2, 0.05)
        final = {v: h.model_get(v) for v in names}
        phi = h.model_output("phi")
    finally:
        h.close()
    res = {"checks": []}
    with netcdf_file(path, "r", mmap=False) as nc:
        res["variables"] = sorted(nc.variables)
        ok = all(v in nc.variables for v in ("n", "vort", "phi", "t_array", "iteration"))
        res["checks"].append(["variables", ok])
        res["checks"].append(["records", ok and nc.variables["t_array"].shape[0] == 3])
        for v in names:
            res["checks"].append([f"final_{v}", ok and np.array_equal(np.array(nc.variables[v][2]), final[v])])
        res["checks"].append(["final_phi", ok and np.array_equal(np.array(nc.variables["phi"][2]), phi)])
        res["checks"].append(["evolves", ok and not np.array_equal(np.array(nc.variables["n"][0]), np.array(nc.variables["n"][2]))])
    res["ok"] = all(c for _, c in res["checks"])
    print(json.dumps(res), flush=True)
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main_hw(sys.argv[1]) if sys.argv[2:] == ["hw"] else main(sys.argv[1]))
