"""Generates the golden fixtures in this directory from the BOUT++ reference itself.

Run in the build container only (needs /root/reference and oracle/_ref/libboutref.so, i.e. the
reference's own sources compiled in place by oracle/Makefile):

    python tests/golden/make_golden.py

Outputs (committed; the GPU box has neither /root/reference nor needs the reference build):
  laplace_benchmark.npz  the reference's golden file tests/integrated/test-laplace/data/benchmark.0.nc
                         (float32, 16 LaplaceCyclic solutions) + the float64 inputs regenerated with the
                         reference's expression parser + the float64 solutions of the reference build.
  ops_small.npz          reference outputs of every operator on the hot path for one small grid with a
                         non-trivial metric (stencils, brackets, Delp2, z-FFT ops, Laplace solves,
                         boundaries, communicate, field algebra).
  ops_shifted.npz        the elm-pb operator vocabulary on a shifted-metric mesh with elm-pb's derivative
                         methods: parallel slices, to/fromFieldAligned, Y operators on slices,
                         b0xGrad_dot_Grad (3D3D, 3D2D, 2D3D), averageY, smooth_x.
  elmpb.npz              examples/elm-pb/elm_pb.cxx (unchanged, shipped [highbeta] switches) on the synthetic
                         shifted-metric tokamak grid of cases.elmpb_ini: post-init equilibrium + metric, one
                         Solver::run_rhs on a seeded (U, P, Psi) state (ddt + phi, Jpar, Jpar2), and a short
                         RK4 trajectory.
  models.npz             unchanged example models (hasegawa-wakatani, hasegawa-wakatani-3d, blob2d) run
                         through the reference's Solver::run_rhs and RK4 / RK3-SSP solvers: initial state,
                         ddt after one RHS, state after a few steps.
  expressions.npz        cases.EXPRESSIONS evaluated by the reference's expression parser / FieldFactory.
  bndry_free.npz         free_o2 / free_o3 boundary conditions on the ops_small grid and field.
  integrators.npz        this repository's test model tests/models/ode_pair.cxx through the reference's own
                         euler, rkgeneric (rkf45, cashkarp, rk4, rkf34; adaptive and fixed) and splitrk solvers
                         (cases.INTEGRATOR_CASES): final state and number of RHS calls.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402

orc = ge.load_oracle()
REF_ROOT = "/root/reference"

LAPLACE_VARS = ["flag0", "flag3", "flag0a", "flag3a", "flag0ac", "flag3ac", "flag0ad", "flag3ad",
                "flagis", "flagos", "flagisa", "flagosa", "flagisac", "flagosac", "flagisad", "flagosad"]

TEST_LAPLACE_INI = """
MZ = 32
MXG = 2
MYG = 2
[mesh]
symmetricGlobalX = false
symmetricGlobalY = false
nx = 12
ny = 16
ixseps1 = 6
ixseps2 = 6
jyseps1_1 = -1
jyseps1_2 = 6
jyseps2_1 = 6
jyseps2_2 = 15
dx = 1
dy = 1
[laplace]
all_terms = false
include_yguards = false
filter = 0.2
"""


def laplace_benchmark():
    """tests/integrated/test-laplace/test_laplace.cxx:19-100 replayed on the reference build."""
    from scipy.io import netcdf_file
    nc = netcdf_file(os.path.join(REF_ROOT, "tests/integrated/test-laplace/data/benchmark.0.nc"), "r", mmap=False)
    bench = {k: nc.variables[k].data.copy() for k in LAPLACE_VARS + ["input", "a", "c", "d", "set_to"]}
    r = orc.Ref(TEST_LAPLACE_INI)
    g = r.grid
    out = {"grid": np.array([g.nx, g.ny, g.nz, g.mxg, g.myg])}
    out["input"] = r.expr3d("(1-gauss(x-0.5,0.2))*gauss(y-pi)*gauss(z-pi)")
    e2 = lambda s: _expr2d(r, s)
    out["a"] = e2("gauss(x) * sin(y)")
    out["c"] = e2("sin(x) * gauss(x-0.5) * gauss(y-pi)")
    out["d"] = e2("y - pi/2")
    out["set_to"] = e2("cos(2*y)*(x - 0.5)")
    for k in ("input",):
        assert np.abs(out[k] - bench[k][:, :, :g.nz]).max() < 1e-7, k
    for k in ("a", "c", "d", "set_to"):
        assert np.abs(out[k] - bench[k]).max() < 1e-6, k
    m = r.metric()
    for n in orc.Metric.NAMES:
        out["metric_" + n] = getattr(m, n)
    set3 = np.ascontiguousarray(np.repeat(out["set_to"][:, :, None], g.nz, axis=2))
    ones, zeros = np.ones(g.shape2), np.zeros(g.shape2)
    DCAC = orc.INVERT_DC_GRAD + orc.INVERT_AC_GRAD
    r.laplace_create()
    ys = slice(g.ystart, g.yend + 1)

    def run(name, A, C, D, inner, outer, x0=None):
        r.laplace_set_coef("A", A)
        r.laplace_set_coef("C", C)
        r.laplace_set_coef("D", D)
        r.laplace_set_flags(0, inner, outer)
        sol = r.laplace_solve(out["input"], x0)
        sol[:, :g.ystart] = 0.0   # include_yguards=false: y guards are not solved
        sol[:, g.yend + 1:] = 0.0
        err = np.abs(sol[:, ys] - bench[name][:, ys, :g.nz]).max()
        assert err < 1e-6, (name, err)   # the reference's own tolerance (runtest:34)
        out["ref64_" + name] = sol
        out["bench32_" + name] = bench[name][:, ys, :g.nz].astype(np.float32)
        print(f"  {name:9s} reference build vs benchmark.0.nc: {err:.2e}")

    for suffix, A, C, D in (("", zeros, ones, ones), ("a", out["a"], ones, ones),
                            ("ac", out["a"], out["c"], ones), ("ad", out["a"], ones, out["d"])):
        run("flag0" + suffix, A, C, D, 0, 0)
        run("flag3" + suffix, A, C, D, DCAC, 0)
        run("flagis" + suffix, A, C, D, orc.INVERT_SET, 0, set3)
        run("flagos" + suffix, A, C, D, 0, orc.INVERT_SET, set3)
    r.close()
    np.savez_compressed(os.path.join(HERE, "laplace_benchmark.npz"), **out)


def _expr2d(r, s):
    import ctypes as C
    out = np.zeros(r.grid.shape2)
    r._check(r.lib.boutref_expr2d(s.encode(), out.ctypes.data_as(C.c_void_p)))
    return out


OPS_INI = orc.make_ini(12, 8, 16, mxg=2, myg=2, dx="0.2*(1+0.1*x)", dy="0.7+0.05*cos(2*pi*x)", dz="2*pi/16",
                       mesh_extra="""
g11 = 1 + 0.1*x
g22 = 1.1 + 0.02*y
g33 = 1.2 + 0.05*x
g12 = 0.01
g13 = 0.05 + 0.02*x
g23 = 0.02
ixseps1 = 12
ixseps2 = 12
""", extra="""
[mesh:ddx]
first = C2
[laplace]
include_yguards = false
""")

INDEX_METHODS = {
    "DDX": ["C2", "C4", "W2", "S2", "W3"], "D2DX2": ["C2", "C4"], "D4DX4": ["C2"],
    "VDDX": ["C2", "C4", "U1", "U2", "U3", "W3"], "FDDX": ["U1", "U2", "C2", "C4"],
}


def smooth_fields(g, n, seed):
    rng = np.random.default_rng(seed)
    X = (np.arange(g.nx)[:, None, None] + 0.5) / g.nx
    Y = 2 * np.pi * np.arange(g.ny)[None, :, None] / g.ny
    Z = 2 * np.pi * np.arange(g.nz)[None, None, :] / g.nz
    return [np.ascontiguousarray(np.sin(2 * np.pi * (1 + i) * X) * np.cos((2 + i) * Z) * np.cos(Y)
                                 + 0.1 * rng.uniform(-1, 1, g.shape3)) for i in range(n)]


def ops_small():
    r = orc.Ref(OPS_INI)
    g = r.grid
    out = {"grid": np.array([g.nx, g.ny, g.nz, g.mxg, g.myg])}
    m = r.metric()
    for n in orc.Metric.NAMES:
        out["metric_" + n] = getattr(m, n)
    f, h, v = smooth_fields(g, 3, 20261017)
    v = v - 0.05  # sign-changing velocity for the upwind branches
    f2 = np.ascontiguousarray(f[:, :, 0] + 1.5)
    out.update(f=f, h=h, v=v, f2=f2)
    for base, methods in INDEX_METHODS.items():
        for d in "XYZ":
            for meth in methods:
                name = "index_" + base.replace("X", d)
                if name in ("index_D4DY4",) and g.myg < 2:
                    continue
                ins = (v, f) if base in ("VDDX", "FDDX") else (f,)
                if base == "FDDX" and d == "Y":
                    pass
                out[f"{name}_{meth}"] = r.op(name, *ins, method=meth)
    out["index_DDZ_FFT"] = r.op("index_DDZ", f, method="FFT")
    out["index_D2DZ2_FFT"] = r.op("index_D2DZ2", f, method="FFT")
    for name in ("DDX", "DDY", "DDZ", "D2DX2", "D2DY2", "D2DZ2", "D2DXDZ"):
        for meth in ("C2", "C4"):
            out[f"{name}_{meth}"] = r.op(name, f, method=meth)
    for name in ("VDDX", "VDDZ"):
        for meth in ("U1", "W3"):
            out[f"{name}_{meth}"] = r.op(name, v, f, method=meth)
    out["bracket_arakawa"] = r.op("bracket", f, h, method="ARAKAWA")
    out["bracket_arakawa_3d2d"] = r.op("bracket_3d2d", f, f2, method="ARAKAWA")
    out["bracket_arakawa_2d3d"] = r.op("bracket_2d3d", f2, h, method="ARAKAWA")
    out["Delp2_fft"] = r.op("Delp2", f)
    out["Delp2_fd"] = r.op("Delp2_fd", f)
    out["filter_3"] = r.op("filter", f, scalar=3)
    out["lowPass_4"] = r.op("lowPass", f, scalar=4)
    out["shiftZ_0.37"] = r.op("shiftZ", f, scalar=0.37)
    out["DC3d"] = r.op("DC3d", f)
    out["averageY_DC"] = r.op("averageY_DC", f)
    out["arith_2a_bc"] = r.op("arith_2a_bc", f, h, v)
    out["mul"] = r.op("mul", f, h)
    out["div_2d"] = r.op("div_2d", f, f2)
    out["s_sub"] = r.op("s_sub", f, scalar=0.25)
    # Laplace solves: flags x coefficient sets (global, inner, outer)
    r.laplace_create()
    X = (np.arange(g.nx)[:, None] + 0.5) / g.nx * np.ones(g.shape2)
    A = np.ascontiguousarray(-0.1 - 0.05 * X)
    D = np.ascontiguousarray(1 + 0.2 * X)
    Cc = np.ascontiguousarray(1 + 0.3 * np.sin(3 * X))
    out.update(lap_A=A, lap_D=D, lap_C=Cc)
    ones, zeros = np.ones(g.shape2), np.zeros(g.shape2)
    flagsets = [(0, 0, 0), (0, 3, 0), (0, 0, 3), (0, 2, 2), (0, 0, 1), (1, 4, 64), (0, 16, 16), (0, 35, 32),
                (4, 0, 0), (0, 128, 0), (0, 19, 32), (0, 1024, 0)]
    out["lap_flagsets"] = np.array(flagsets)
    for ci, (a_, c_, d_) in enumerate(((zeros, ones, ones), (A, Cc, D))):
        r.laplace_set_coef("A", a_)
        r.laplace_set_coef("C", c_)
        r.laplace_set_coef("D", d_)
        for fi, fl in enumerate(flagsets):
            r.laplace_set_flags(*fl)
            sol = r.laplace_solve(f, h)
            sol[:, :g.ystart] = 0.0
            sol[:, g.yend + 1:] = 0.0
            out[f"lap_c{ci}_f{fi}"] = sol
    # boundaries + communicate (ixseps = nx: y periodic on one rank)
    for cond in ("dirichlet", "dirichlet_o2", "neumann", "dirichlet(0.7)", "neumann(0.7)"):
        out["bndry_" + cond] = r.apply_boundary(f, cond)
    out["communicate"] = r.communicate(f)[0]
    r.close()
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **out)
    print("  ops_small:", len(out), "arrays")


from cases import HW_INI, HW3D_INI, BLOB_INI, ELMPB_INI, CONDUCTION_INI  # noqa: E402


MODEL_CASES = {"hw": ("hw", HW_INI, "rk4", 5), "hw_rk3": ("hw", HW_INI, "rk3ssp", 5),
               "hw3d": ("hw3d", HW3D_INI, "rk4", 5), "blob2d": ("blob2d", BLOB_INI, "rk4", 5),
               "conduction": ("conduction", CONDUCTION_INI, "rk4", 5),
               "blob2d_outer": ("blob2d_outer", BLOB_INI, "rk4", 5)}


SHIFTED_INI = orc.make_ini(12, 8, 16, mxg=2, myg=2, dx="0.2*(1+0.1*x)", dy="0.7+0.05*cos(2*pi*x)", dz="2*pi/16",
                           mesh_extra="""
g11 = 1 + 0.1*x
g22 = 1.1 + 0.02*cos(y)
g33 = 1.2 + 0.05*x
g12 = 0.01
g13 = 0.05 + 0.02*x
g23 = 0.02
zShift = 0.4*(1+x)*sin(y) + 0.1*(1+x)*y
ShiftAngle = 0.1*(1+x)*2*pi
ixseps1 = 12
ixseps2 = 12
[mesh:paralleltransform]
type = shifted
""", extra="""
[mesh:ddx]
first = C4
second = C4
upwind = W3
[mesh:ddy]
first = C2
second = C2
upwind = C2
flux = C2
[mesh:ddz]
first = C4
second = C4
upwind = W3
""")


def ops_shifted():
    """elm-pb's operator vocabulary on a shifted-metric mesh with the derivative methods of
    examples/elm-pb/data/BOUT.inp: parallel slices, field-aligned transforms, the Y operators on
    slices (fields communicated first, '@comm'), b0xGrad_dot_Grad in its three operand mixes, averageY,
    smooth_x."""
    r = orc.Ref(SHIFTED_INI)
    g = r.grid
    out = {"grid": np.array([g.nx, g.ny, g.nz, g.mxg, g.myg])}
    m = r.metric()
    for n in orc.Metric.NAMES:
        out["metric_" + n] = getattr(m, n)
    zs, zl = r.zshift()
    out["zShift"] = zs
    out["zlength"] = np.array(zl)
    f, h, v = smooth_fields(g, 3, 20261018)
    v = v - 0.05
    f2 = np.ascontiguousarray(f[:, :, 0] + 1.5)
    # the reference communicates before it uses slices: hand both sides the communicated fields
    f, h, v = r.communicate(f, h, v)
    out.update(f=f, h=h, v=v, f2=f2)
    for off in (1, 2, -1, -2):
        out[f"slice_{off:+d}"] = r.parallel_slice(f, off)
        out[f"slice_h_{off:+d}"] = r.parallel_slice(h, off)
    out["toFieldAligned"] = r.op("toFieldAligned", f)
    out["fromFieldAligned"] = r.op("fromFieldAligned", f)
    for name in ("DDY", "D2DY2", "Grad_par", "Div_par", "Grad2_par2"):
        out[name + "_slices"] = r.op(name + "@comm", f)
    out["VDDY_slices"] = r.op("VDDY@comm", v, f)
    out["Vpar_Grad_par_slices"] = r.op("Vpar_Grad_par@comm", v, f)
    out["b0xGrad_dot_Grad"] = r.op("b0xGrad_dot_Grad@comm", f, h)
    out["b0xGrad_dot_Grad_3d2d"] = r.op("b0xGrad_dot_Grad_3d2d@comm", f, f2)
    out["b0xGrad_dot_Grad_2d3d"] = r.op("b0xGrad_dot_Grad_2d3d@comm", f2, h)
    out["averageY"] = r.op("averageY", f)
    out["smooth_x"] = r.op("smooth_x", f)
    r.close()
    np.savez_compressed(os.path.join(HERE, "ops_shifted.npz"), **out)


def one_model(tag, path):
    """One model case per process: the reference keeps static state that does not survive many
    init/finalise cycles."""
    name, ini, solver, nsteps = MODEL_CASES[tag]
    out = {}
    dt = float([l for l in ini.splitlines() if l.startswith("timestep")][0].split("=")[1])
    # pass 1: Solver::solve with zero outputs = init + one run_rhs at the initial state
    r = orc.Ref(ini)
    g = r.grid
    names = r.model_create(name, solver)
    out[tag + "_grid"] = np.array([g.nx, g.ny, g.nz, g.mxg, g.myg])
    m = r.metric()
    for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13", "G1", "G3", "d1_dx", "g_11"):
        out[f"{tag}_metric_{n}"] = getattr(m, n)
    for v in names:
        out[f"{tag}_init_{v}"] = r.model_get(v)
    r.model_solve(0, dt)
    for v in names:
        out[f"{tag}_ddt_{v}"] = r.model_ddt(v)
    r.close()
    # pass 2: nsteps fixed steps of the reference's own RK solver
    r = orc.Ref(ini)
    r.model_create(name, solver)
    r.model_solve(1, nsteps * dt)
    for v in names:
        out[f"{tag}_final_{v}"] = r.model_get(v)
    out[tag + "_dt_nsteps"] = np.array([dt, nsteps])
    print(f"  model {tag}: vars {names}, max|final| = {[float(np.abs(out[f'{tag}_final_{v}']).max()) for v in names]}")
    np.savez(path, **out)


def bndry_free():
    """free_o2 / free_o3 (BoundaryFree_O2/O3::apply) on the ops_small grid and field: tests/golden/bndry_free.npz."""
    ops = np.load(os.path.join(HERE, "ops_small.npz"))
    r = orc.Ref(OPS_INI)
    out = {"bndry_" + cond: r.apply_boundary(ops["f"], cond) for cond in ("free_o2", "free_o3")}
    r.close()
    np.savez_compressed(os.path.join(HERE, "bndry_free.npz"), **out)
    print("  bndry_free:", {k: float(np.abs(v - ops["f"]).max()) for k, v in out.items()})


def expressions():
    """cases.EXPRESSIONS through the reference's FieldFactory::create3D: tests/golden/expressions.npz."""
    from cases import EXPR_INI, EXPRESSIONS
    r = orc.Ref(EXPR_INI)
    out = {str(i): r.expr3d(e) for i, e in enumerate(EXPRESSIONS)}
    r.close()
    np.savez_compressed(os.path.join(HERE, "expressions.npz"), **out)
    print("  expressions:", len(out))


def options_probe():
    """cases.OPTIONS_PROBE_EXPRS through the reference: tests/golden/options_probe.json."""
    import json
    from cases import OPTIONS_PROBE_EXPRS, OPTIONS_PROBE_INI
    r = orc.Ref(OPTIONS_PROBE_INI)
    res = {}
    for e in OPTIONS_PROBE_EXPRS:
        try:
            res[e] = float(r.expr3d(e)[2, 2, 1])
        except Exception:  # noqa: BLE001 -- the reference rejects the reference
            res[e] = "ERR"
    r.close()
    json.dump(res, open(os.path.join(HERE, "options_probe.json"), "w"), indent=1)
    print("  options_probe:", res)


def one_integrator(tag, path):
    """tests/models/ode_pair.cxx through the reference's own euler / rkgeneric / splitrk solvers."""
    from cases import ELEM_CASES, HW_INTEGRATOR_CASES, INTEGRATOR_CASES, hw_integrator_ini, ode_ini
    if tag in HW_INTEGRATOR_CASES:   # unchanged examples/hasegawa-wakatani/hw.cxx
        solver, opts, nout, step = HW_INTEGRATOR_CASES[tag]
        r = orc.Ref(hw_integrator_ini(opts, nout, step))
        names = r.model_create("hw", solver)
    elif tag == "imex_shipped":     # unchanged examples/IMEX/advection-diffusion/imex.cxx, shipped input
        from cases import IMEX_INI, IMEX_NOUT, IMEX_STEP
        solver, nout, step = "splitrk", IMEX_NOUT, IMEX_STEP
        r = orc.Ref(IMEX_INI)
        names = r.model_create("imex", solver)
    elif tag == "advreact":          # unchanged examples/IMEX/advection-reaction/split_operator.cxx
        from cases import ADVREACT_INI
        solver, nout, step = "splitrk", 1, 0.01
        r = orc.Ref(ADVREACT_INI)
        names = r.model_create("advreact", solver)
    elif tag == "eigenbox_shipped":  # unchanged examples/eigen-box/eigen-box.cxx, shipped input, rk4
        from cases import EIGENBOX_INI
        solver, nout, step = "rk4", 1, 0.2
        r = orc.Ref(EIGENBOX_INI)
        names = r.model_create("eigenbox", solver)
    elif tag in ELEM_CASES:         # tests/models/elem_access.cxx
        solver, opts, nout, step = ELEM_CASES[tag]
        r = orc.Ref(ode_ini(opts, False, nout, step))
        names = r.model_create("elem_access", solver)
    else:
        solver, opts, split, nout, step, *bndry = INTEGRATOR_CASES[tag]
        r = orc.Ref(ode_ini(opts, split, nout, step, *bndry))
        names = r.model_create("ode_pair", solver)
    out = {f"{tag}_init_{v}": r.model_get(v) for v in names}
    r.model_solve(nout, step)
    for v in names:
        out[f"{tag}_final_{v}"] = r.model_get(v)
    out[f"{tag}_rhs_calls"] = np.array(r.model_rhs_calls())
    print(f"  integrator {tag}: rhs calls {int(out[f'{tag}_rhs_calls'])}, max|final| = "
          f"{[float(np.abs(out[f'{tag}_final_{v}']).max()) for v in names]}")
    np.savez(path, **out)


def integrators():
    import subprocess
    import tempfile
    from cases import ELEM_CASES, HW_INTEGRATOR_CASES, INTEGRATOR_CASES
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag in list(INTEGRATOR_CASES) + list(HW_INTEGRATOR_CASES) + list(ELEM_CASES) + ["imex_shipped", "eigenbox_shipped", "advreact"]:
            path = os.path.join(tmp, tag + ".npz")
            subprocess.run([sys.executable, os.path.abspath(__file__), "integrator:" + tag, path], check=True)
            out.update(dict(np.load(path)))
    np.savez_compressed(os.path.join(HERE, "integrators.npz"), **out)


from elmpb_names import ELMPB_FIELDS, ELMPB_SCALARS  # noqa: E402
ELMPB_DT, ELMPB_NSTEPS = 2e-3, 5


def elmpb_states(g, U0):
    """Seeded smooth (U, P, Psi) of elm-pb-like magnitudes for the single-RHS case."""
    f, h, v = smooth_fields(g, 3, 20261019)
    return U0 + 2e-6 * f, 3e-6 * h, 4e-7 * v


def elmpb_case(which, path):
    r = orc.Ref(ELMPB_INI)
    g = r.grid
    out = {}
    names = r.model_create("elmpb", "rk4")
    U0 = r.model_get("U")
    U, P, Psi = elmpb_states(g, U0)
    if which == "rhs":
        out["grid"] = np.array([g.nx, g.ny, g.nz, g.mxg, g.myg])
        m = r.metric()
        for n in orc.Metric.NAMES:
            out["metric_" + n] = getattr(m, n)
        zs, zl = r.zshift()
        out["zShift"], out["zlength"] = zs, np.array(zl)
        for n in ELMPB_FIELDS:
            out["eq_" + n] = r.elmpb_get(n)
        out["scalars"] = np.array([r.elmpb_get(n) for n in ELMPB_SCALARS])
        out["init_U"] = U0
        out["rhs_in_U"], out["rhs_in_P"], out["rhs_in_Psi"] = U, P, Psi
        r.model_set("U", U), r.model_set("P", P), r.model_set("Psi", Psi)
        r.model_rhs(0.0)
        for v in names:
            out["rhs_ddt_" + v] = r.model_ddt(v)
            out["rhs_bc_" + v] = r.model_get(v)      # the state after Solver::pre_rhs boundary conditions
        for n in ("phi", "Jpar", "Jpar2", "Psi_loc"):
            out["rhs_" + n] = r.elmpb_get(n)
    else:
        r.model_set("U", U), r.model_set("P", P), r.model_set("Psi", Psi)
        r.model_solve(1, ELMPB_NSTEPS * ELMPB_DT)
        for v in names:
            out["final_" + v] = r.model_get(v)
        out["dt_nsteps"] = np.array([ELMPB_DT, ELMPB_NSTEPS])
        print("  elmpb trajectory: max|final| =", [float(np.abs(out["final_" + v]).max()) for v in names])
    np.savez(path, **out)


def elmpb():
    import subprocess
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for which in ("rhs", "traj"):
            path = os.path.join(tmp, which + ".npz")
            subprocess.run([sys.executable, os.path.abspath(__file__), "elmpb:" + which, path], check=True)
            out.update(dict(np.load(path)))
    np.savez_compressed(os.path.join(HERE, "elmpb.npz"), **out)


def models():
    import subprocess
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag in MODEL_CASES:
            path = os.path.join(tmp, tag + ".npz")
            subprocess.run([sys.executable, os.path.abspath(__file__), "model:" + tag, path], check=True)
            out.update(dict(np.load(path)))
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["laplace", "ops", "shifted", "models", "elmpb", "integrators", "bndry_free", "expressions", "options_probe"]
    if which[0].startswith("elmpb:"):
        elmpb_case(which[0][6:], which[1])
        sys.stdout.flush()
        os._exit(0)
    if which[0].startswith("integrator:"):
        one_integrator(which[0][11:], which[1])
        sys.stdout.flush()
        os._exit(0)
    if which[0].startswith("model:"):
        one_model(which[0][6:], which[1])
        sys.stdout.flush()
        os._exit(0)
    if "laplace" in which:
        laplace_benchmark()
    if "ops" in which:
        ops_small()
    if "shifted" in which:
        ops_shifted()
    if "models" in which:
        models()
    if "elmpb" in which:
        elmpb()
    if "integrators" in which:
        integrators()
    if "bndry_free" in which:
        bndry_free()
    if "expressions" in which:
        expressions()
    if "options_probe" in which:
        options_probe()
    os._exit(0)  # the reference's static destructors are not exit-safe after repeated init/finalise
