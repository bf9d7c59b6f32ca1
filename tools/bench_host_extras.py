"""Child process of bench.py's `other` extras (never part of the headline number): the unchanged examples/blob2d
source at 1024 x 1024 (BASELINE configs[2]) through the C++ host layer with the integrators beyond rk4 and with the
output step enabled.  Runs in its own process and without torch, so that nothing here can disturb the main line.
Prints cumulative JSON lines (the last one is complete):
  integrators[<solver>]: wall ms per RHS evaluation over one Solver::solve of `nsteps` internal steps,
  output_step: wall ms of the same rk4 solve with nout output records written ([output] path) and without, and the
               difference per record (what the solver waits for; the device -> host -> disk work overlaps the steps).
"""
import importlib.util
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def load_host():
    spec = importlib.util.spec_from_file_location("b200_host_binding", os.path.join(ROOT, "bout-dev_b200", "host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def blob_ini(solver_opts, extra=""):
    import cases
    body = cases.BLOB_INI[cases.BLOB_INI.index("[mesh:ddx]"):]
    assert "[solver]\ntimestep = 0.05" in body
    body = body.replace("[solver]\ntimestep = 0.05", "[solver]\n" + solver_opts)
    return cases.make_ini(1028, 1, 1024, mxg=2, myg=0, dx="0.3", dz="0.3", extra=body) + extra


def solve(hostmod, ini, solver, nout, step):
    h = hostmod.Host(ini)
    try:
        names = h.model_create("blob2d", solver)
        h.model_rhs(0.0)
        h.model_get(names[0])
        c0 = h.rhs_calls()
        t0 = time.time()
        h.model_solve(nout, step)
        final = h.model_get(names[0])
        dt = time.time() - t0
        return dt, h.rhs_calls() - c0, bool((final == final).all())
    finally:
        h.close()


def main():
    hostmod = load_host()
    res = {"integrators": {}, "workload": "examples/blob2d unchanged, 1024x1x1024, through bout-dev_b200/host"}
    dt = 0.05
    cases_ = {"rk4": ("rk4", f"timestep = {dt}"),
              "rk3ssp": ("rk3ssp", f"timestep = {dt}"),
              "euler": ("euler", f"timestep = {dt}\nmxstep = 100000"),
              "rkgeneric_cashkarp_fixed": ("rkgeneric", f"scheme = cashkarp\nadaptive = false\ntimestep = {dt}\nmax_timestep = 100"),
              "rkgeneric_rkf45_adaptive": ("rkgeneric", f"timestep = {dt}\nmax_timestep = 100\nrtol = 1e-5\natol = 1e-10\nmxstep = 2000"),
              "splitrk_fixed": ("splitrk", f"adaptive = false\ntimestep = {dt}\nnstages = 4"),
              "adams_bashforth_adaptive": ("adams-bashforth", f"timestep = {dt}\nmax_timestep = 100\nrtol = 1e-5\nmxstep = 4000")}
    solve(hostmod, blob_ini(cases_["rk4"][1]), "rk4", 1, 5 * dt)   # warm-up instance (library load, pool)
    try:
        nout, per = 8, 5
        plain, calls, _ = solve(hostmod, blob_ini(cases_["rk4"][1]), "rk4", nout, per * dt)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "BOUT.dmp.0.nc")
            dumped, _, _ = solve(hostmod, blob_ini(cases_["rk4"][1], f"\n[output]\npath = {path}\n"), "rk4", nout, per * dt)
            size = os.path.getsize(path)
        res["output_step"] = {"records": nout + 1, "file_MB": size / 1e6, "solve_ms_without": plain * 1e3,
                              "solve_ms_with": dumped * 1e3, "added_ms_per_record": (dumped - plain) * 1e3 / (nout + 1),
                              "rhs_calls": calls}
    except Exception as e:  # noqa: BLE001
        res["output_step"] = {"error": repr(e)[:200]}
    print(json.dumps(res), flush=True)
    for tag, (solver, opts) in cases_.items():
        try:
            wall, calls, finite = solve(hostmod, blob_ini(opts), solver, 1, 40 * dt)
            res["integrators"][tag] = {"wall_ms": wall * 1e3, "rhs_calls": calls, "ms_per_rhs_call": wall * 1e3 / max(calls, 1),
                                       "finite": finite}
        except Exception as e:  # noqa: BLE001
            res["integrators"][tag] = {"error": repr(e)[:200]}
        print(json.dumps(res), flush=True)   # cumulative: bench.py keeps the last complete line if the time limit hits
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa: BLE001 -- reported, never raised into bench.py
        print(json.dumps({"error": repr(e)[:300]}), flush=True)
