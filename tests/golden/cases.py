"""BOUT.inp texts of the model cases behind tests/golden/models.npz (shared by the generator, which
feeds them to the reference build, and by the GPU tests, which feed them to the B200 host layer)."""


def make_ini(nx, ny, nz, mxg=2, myg=2, dx="0.2", dy="1.0", dz="0.2", periodic_x=False, extra="", mesh_extra=""):
    """BOUT.inp text for a mesh-from-options (reference: src/mesh/data/gridfromoptions.cxx)."""
    return f"""
MXG = {mxg}
MYG = {myg}
periodicX = {'true' if periodic_x else 'false'}
[mesh]
nx = {nx}
ny = {ny}
nz = {nz}
dx = {dx}
dy = {dy}
dz = {dz}
{mesh_extra}
{extra}
"""


HW_INI = make_ini(36, 1, 32, mxg=2, myg=0, periodic_x=True, extra="""
[hw]
alpha = 1.0
kappa = 0.5
Dvort = 1e-4
Dn = 1e-4
modified = true
bracket = 2
[all]
scale = 0.0
bndry_all = dirichlet_o2
[vort]
scale = 0.1
function = mixmode(2*pi*x) * mixmode(z)
[n]
scale = 0.05
function = mixmode(2*pi*x + 1) * mixmode(z + 0.5)
[solver]
timestep = 0.01
[input]
error_on_unused_options = false
""")

HW3D_INI = make_ini(20, 8, 16, mxg=2, myg=2, mesh_extra="ixseps1 = 20\nixseps2 = 20\n", extra="""
[hw]
alpha = 1.0
kappa = 0.5
Dvort = 0.001
Dn = 0.001
[all]
scale = 0.0
bndry_all = dirichlet_o2
[vort]
scale = 0.1
function = mixmode(2*pi*x) * mixmode(z - y)
[n]
scale = 0.05
function = mixmode(2*pi*x + 1) * mixmode(z + y)
[solver]
timestep = 0.01
[input]
error_on_unused_options = false
""")

BLOB_INI = make_ini(36, 1, 32, mxg=2, myg=0, dx="0.3", dz="0.3", extra="""
[mesh:ddx]
first = C2
second = C2
upwind = W3
[mesh:ddz]
first = C2
second = C2
upwind = W3
[model]
Te0 = 5
n0 = 2e18
D_vort = 1e-6
D_n = 1e-6
R_c = 1.5
L_par = 10
B0 = 0.35
compressible = false
boussinesq = true
sheath = true
[phiBoussinesq]
inner_boundary_flags = 0
outer_boundary_flags = 0
[phiSolver]
inner_boundary_flags = 0
[all]
scale = 0.0
bndry_all = neumann
[n]
scale = 1.0
function = 1 + 0.5*exp(-((x-0.25)/0.1)^2 - ((z/(2*pi) - 0.5)/0.1)^2)
[solver]
timestep = 0.05
[input]
error_on_unused_options = false
""")




# examples/conduction/conduction.cxx (unchanged: ddt(T) = Div_par_K_Grad_par(chi, T)) on a periodic-y mesh whose dy and
# g_22 vary along y; the shipped input (MXG = 0, one x point, dirichlet_o4 on open field lines, pvode) is not used.
CONDUCTION_INI = make_ini(8, 16, 4, mxg=2, myg=2, dy="0.2*(1+0.1*cos(y))", mesh_extra="g22 = 1.1 + 0.05*sin(y)\nixseps1 = 8\nixseps2 = 8\n", extra="""
[mesh:ddy]
first = C2
second = C2
[conduction]
chi = 0.7
[all]
scale = 0.0
bndry_all = neumann
[T]
scale = 1.0
function = gauss(y-pi, 0.4) * (1 + 0.1*cos(z))
[solver]
timestep = 0.005
[input]
error_on_unused_options = false
""")


def elmpb_ini(nx, ny, nz, timestep="1e-3", extra_highbeta=""):
    """examples/elm-pb/data/BOUT.inp (methods, solver flags, boundary conditions, [highbeta] switches as
    shipped) on a SYNTHETIC closed-form circular-cross-section tokamak grid given as mesh options (the
    shipped cbm18 grid file is NetCDF, which this build cannot read): all-core (ixseps = nx, y periodic),
    shifted metric, zperiod = 15.  nx includes the 2*MXG guard cells; ny is the interior count."""
    mx = nx - 4
    return f"""
MXG = 2
MYG = 2
zperiod = 15
[mesh]
staggergrids = false
nx = {nx}
ny = {ny}
nz = {nz}
ixseps1 = {nx}
ixseps2 = {nx}
rminor = 0.4 + 0.3*x
qsafe = 1.2 + 2*x^2
Rxy = 3 + mesh:rminor*cos(y)
Btxy = 6/mesh:Rxy
Bpxy = mesh:rminor*mesh:Btxy/(mesh:qsafe*mesh:Rxy)
Bxy = sqrt(mesh:Btxy^2 + mesh:Bpxy^2)
hthe = mesh:rminor*(1 + 0.05*cos(y))
sinty = 1.5*x*(y - pi)
psixy = x
psi_axis = 0
psi_bndry = 1
bmag = 2.0
rmag = 3.0
dx = (0.15/{mx})*(1 + 0.2*x)
dy = 2*pi/{ny}
dz = 2*pi/15/{nz}
zShift = mesh:qsafe*(y - pi + 0.2*mesh:rminor*sin(y))
ShiftAngle = 2*pi*mesh:qsafe
Jpar0 = 8e5*(1 - x^2)*(1 + 0.1*cos(y))
pressure = 2e4*(1 - tanh((x - 0.6)/0.15)) + 500
bxcvx = 0.02*sin(y)*(1 + 0.3*x)
bxcvy = 0.003*cos(y)
bxcvz = 0.05*cos(y) + 0.01*x
[mesh:paralleltransform]
type = shifted
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
[solver]
timestep = {timestep}
[highbeta]
density = 1e+19
evolve_jpar = false
evolve_pressure = true
phi_constraint = false
include_jpar0 = true
include_curvature = true
compress = false
nonlinear = false
diamag = true
diamag_grad_t = false
diamag_phi0 = true
bm_exb_flag = 0
bm_mag_flag = 0
withflow = false
eHall = false
AA = 2.0
noshear = false
relax_j_vac = false
filter_z = true
filter_z_mode = 1
low_pass_z = 16
zonal_flow = false
zonal_field = false
zonal_bkgd = true
smooth_j_x = true
include_rmp = false
vacuum_pressure = 0.02
vacuum_trans = 0.01
vac_lund = 1.e8
core_lund = 1.e8
hyperresist = 1.e-4
damp_width = -1
diffusion_par = -1.0
diffusion_p4 = -1e-05
diffusion_u4 = -1e-05
diffusion_a4 = -1e-05
heating_P = -1
sink_P = -1
sink_Ul = -10.0
sink_Ur = -10.0
viscos_par = -0.1
viscos_perp = -1.0
hyperviscos = -1.0
phi_curv = true
{extra_highbeta}
[phiSolver]
inner_boundary_flags = 4
outer_boundary_flags = 5
[aparSolver]
inner_boundary_flags = 4
outer_boundary_flags = 5
[all]
scale = 0.0
bndry_all = dirichlet_o2
[U]
scale = 1e-05
function = ballooning(gauss(x-0.5,0.1)*gauss(y-pi,0.6*pi)*sin(z),3)
[P]
bndry_core = neumann
[Psi]
bndry_xin = zerolaplace
bndry_xout = zerolaplace
[Psi_loc]
bndry_xin = zerolaplace
bndry_xout = zerolaplace
bndry_yup = free_o3
bndry_ydown = free_o3
[J]
bndry_core = neumann
[phi]
bndry_xin = none
bndry_xout = none
[input]
error_on_unused_options = false
"""


ELMPB_INI = elmpb_ini(20, 8, 16, timestep="2e-3")


# ---- explicit integrators beyond rk4 / rk3ssp on tests/models/ode_pair.cxx (tests/golden/integrators.npz) --------
def ode_ini(solver_opts, split, nout, output_step, bndry=None):
    """12x1x8 periodic-X mesh without boundaries; [solver] options per case; root nout / timestep = the
    arguments later given to Solver::solve, so option defaults that derive from them agree on both sides."""
    return f"nout = {nout}\ntimestep = {output_step}\n" + make_ini(12, 1, 8, mxg=2, myg=0, periodic_x=bndry is None, extra=f"""
[ode]
damp = 0.3
split = {'true' if split else 'false'}
[all]
scale = 0.0
{bndry or ''}
[f]
scale = 1.0
function = 1 + 0.5*sin(2*pi*x)*cos(z)
[g]
scale = 1.0
function = 0.3*cos(2*pi*x) + 0.2*sin(z)
[solver]
{solver_opts}
[input]
error_on_unused_options = false
""")


# tag: (solver type, [solver] options, split-operator model?, nout, output_step[, boundary options of [all]])
# The "bc_" cases have X boundaries (periodicX = false): the comparison then covers the guard cells as well, i.e. the
# boundary-condition names, arguments and the relax() modifier as Solver::pre_rhs applies them.
INTEGRATOR_CASES = {
    "bc_dirichlet_neumann": ("rk4", "timestep = 0.0625", False, 1, 0.25, "bndry_xin = dirichlet(0.5)\nbndry_xout = neumann(0.2)"),
    "bc_free_relax": ("rk3ssp", "timestep = 0.0625", True, 1, 0.25, "bndry_xin = free_o3\nbndry_xout = relax(dirichlet_o2(0.25), 10)"),
    "bc_free_o2_all": ("euler", "timestep = 0.0625", False, 1, 0.25, "bndry_all = free_o2"),
    "rk4_fixed": ("rk4", "timestep = 0.07", False, 2, 0.5),
    "rk4_adaptive": ("rk4", "adaptive = true\ntimestep = 0.25\nmax_timestep = 0.5\natol = 1e-8\nrtol = 1e-7", False, 2, 0.5),
    "rk3ssp_split": ("rk3ssp", "timestep = 0.07", True, 1, 0.5),
    "euler": ("euler", "timestep = 0.01\nmxstep = 1000", False, 2, 0.1),
    "euler_split": ("euler", "timestep = 0.03\nmxstep = 1000", True, 1, 0.1),
    "rkf45_adaptive": ("rkgeneric", "timestep = 0.2\nmax_timestep = 0.5\natol = 1e-8\nrtol = 1e-6", False, 2, 0.5),
    "rkf45_follow_high": ("rkgeneric", "timestep = 0.2\nmax_timestep = 0.5\natol = 1e-8\nrtol = 1e-6\nfollowHighOrder = true",
                          False, 1, 0.5),
    "cashkarp_fixed": ("rkgeneric", "scheme = cashkarp\nadaptive = false\ntimestep = 0.07\nmax_timestep = 0.5", False, 1, 0.5),
    "cashkarp_adaptive": ("rkgeneric", "scheme = cashkarp\ntimestep = 0.25\nmax_timestep = 0.5\natol = 1e-9\nrtol = 1e-7",
                          True, 1, 0.5),
    "rk4simple_adaptive": ("rkgeneric", "scheme = rk4\ntimestep = 0.25\nmax_timestep = 0.5\natol = 1e-8\nrtol = 1e-6\n"
                           "followHighOrder = true", False, 1, 0.5),
    "rkf34_adaptive": ("rkgeneric", "scheme = rkf34\ntimestep = 0.25\nmax_timestep = 0.5\natol = 1e-8\nrtol = 1e-5", False, 1, 0.5),
    "splitrk_fixed": ("splitrk", "adaptive = false\ntimestep = 0.06\nnstages = 5", True, 1, 0.5),
    "splitrk_adaptive": ("splitrk", "timestep = 0.25\nmax_timestep = 0.5\nrtol = 1e-4\natol = 1e-10\nnstages = 6", True, 2, 0.5),
    # fixed-step Adams-Bashforth: binary-exact timesteps, so that no rounding-sized last step (two history times 1e-17
    # apart make the reference's own Lagrange coefficients ill-conditioned) enters the comparison
    "ab_fixed_order3": ("adams-bashforth", "adaptive = false\ntimestep = 0.0625\nmax_timestep = 0.5\norder = 3", False, 2, 0.5),
    "ab_fixed_order8": ("adams-bashforth", "adaptive = false\ntimestep = 0.03125\nmax_timestep = 0.5\norder = 8", True, 1, 0.5),
    "ab_adaptive": ("adams-bashforth", "timestep = 0.05\nmax_timestep = 0.5\nrtol = 1e-6", False, 2, 0.5),
    "ab_adaptive_follow_low": ("adams-bashforth", "timestep = 0.05\nmax_timestep = 0.5\nrtol = 1e-6\norder = 4\n"
                               "followHighOrder = false\nadaptive_order = false", True, 1, 0.5),
    "splitrk_adapt_period3": ("splitrk", "timestep = 0.05\nmax_timestep = 0.5\nrtol = 1e-4\natol = 1e-10\nnstages = 5\n"
                              "adapt_period = 3", True, 2, 0.5),
    "splitrk_nonsplit_model": ("splitrk", "adaptive = false\ntimestep = 0.1\nnstages = 4", False, 1, 0.5),
}


# the same integrators on the UNCHANGED examples/hasegawa-wakatani/hw.cxx (a split-operator model; GPU tests only)
def hw_integrator_ini(solver_opts, nout, output_step):
    assert "[solver]\ntimestep = 0.01" in HW_INI
    return f"nout = {nout}\ntimestep = {output_step}\n" + HW_INI.replace("[solver]\ntimestep = 0.01", "[solver]\n" + solver_opts)


# tag: (solver type, [solver] options, nout, output_step) -- fixed-step settings, so the comparison does not hinge on
# an accept / reject decision taken at rounding level (the controllers are pinned by INTEGRATOR_CASES)
HW_INTEGRATOR_CASES = {
    "hw_euler": ("euler", "timestep = 0.005\nmxstep = 1000", 1, 0.05),
    "hw_cashkarp_fixed": ("rkgeneric", "scheme = cashkarp\nadaptive = false\ntimestep = 0.01\nmax_timestep = 0.05", 1, 0.05),
    "hw_rkf45_fixed": ("rkgeneric", "adaptive = false\ntimestep = 0.01\nmax_timestep = 0.05", 1, 0.05),
    "hw_splitrk_fixed": ("splitrk", "adaptive = false\ntimestep = 0.01\nnstages = 4", 1, 0.05),
}


# host-side element access (tests/models/elem_access.cxx): tag: (solver type, [solver] options, nout, output_step)
ELEM_CASES = {
    "elem_rk4": ("rk4", "timestep = 0.0625", 2, 0.25),
    "elem_rkf45": ("rkgeneric", "timestep = 0.1\nmax_timestep = 0.25\natol = 1e-9\nrtol = 1e-7", 1, 0.25),
}


# examples/IMEX/advection-diffusion (unchanged imex.cxx, a split-operator model) with its SHIPPED data/BOUT.inp
# (solver type = splitrk, nstages = 9, adapt_period = 5, W3 upwinding, C4 second derivatives); the text below is that
# file's content with comments dropped, restated here because /root/reference does not exist on the GPU box.
# Only nout is reduced (50 -> 2).
IMEX_INI = """
nout = 2
periodicX = true
MYG = 0
[mesh]
nx = 132
ny = 1
nz = 128
dx = 1 / nx
dz = 1 / (nz-4)
[mesh:ddx]
first = C4
second = C4
upwind = W3
flux = U1
[mesh:ddy]
first = C4
second = C4
upwind = W3
flux = U1
[mesh:ddz]
first = C4
second = C4
upwind = W3
flux = U1
[solver]
type = splitrk
timestep = 0.0001
nstages = 9
adapt_period = 5
output_step = 0.01
[imex]
Vx = 2.0
Dz = 1.0
[U]
zs = z / (2π)
function = exp(-((x-0.5)/0.1)^2 - ((zs-0.5)/0.1)^2)
[input]
error_on_unused_options = false
"""
IMEX_NOUT, IMEX_STEP = 2, 0.01


# Expressions evaluated by the reference's FieldFactory (tests/golden/expressions.npz) and by the host layer's
# evaluator: operators with the reference's precedences and quirks ("a * -b" is (a*0) - b, "2^-1*x" is 2^0 - 1*x, all
# operators left-associative), implicit product after a number, UTF-8 symbols, every generator of field_factory.cxx
# except ballooning / is_periodic_y.
EXPR_INI = make_ini(12, 6, 8, mxg=2, myg=1, mesh_extra="ixseps1 = 12\nixseps2 = 12\n")
EXPRESSIONS = [
    "sin(2*pi*x) + cos(z - y)", "2π*x", "3x + 2 z", "exp(-((x-0.5)/0.1)^2)", "gauss(x-0.5, 0.1)", "gauss(z-pi)",
    "mixmode(2*pi*x)*mixmode(z)", "mixmode(y, 3.5)", "H(x-0.5)", "tanh(3*(x-0.5))", "abs(x-0.5)^1.5", "sqrt(x+1)",
    "min(x, 0.3, z/10)", "max(x,y)", "power(x+1, 2.5)", "log(x+1)", "sinh(x) + cosh(z)", "erf(2*x-1)", "round(3.6*x)",
    "-x^2", "2^-1 * x", "1e-3*x + 1.5E2", "atan(x)", "atan(x-0.5, z+1)", "x - -y", "tan(x)", "asin(x/2) + acos(x/2)",
    "1/2π", "2x^2", "fmod(z, 1.5)", "tanhhat(x, 0.3, 0.5, 0.1)", "clamp(x, 0.2, 0.7)", "where(x - 0.5, y, z)", "x * -y",
    "2^3^2", "x < 0.5 | z > 3", "!(x > 0.5) & true", "2 - 3 - 4 * 2 / 4 ^ 2", "(x + 1)^2 * 3 - 1", "-(x+1)", "1 - -(y)",
]


# Options-file syntax seen through option references in expressions (tests/golden/options_probe.json, values from the
# reference's OptionINI reader + FieldFactory; "ERR" = the reference throws): comments, blanks, multi-line values with
# open brackets, bare flags, sections, case-sensitive names, parent-section lookup.
OPTIONS_PROBE_INI = make_ini(8, 4, 4, mxg=1, myg=1, mesh_extra="ixseps1 = 8\nixseps2 = 8\n", extra="""
[sec]
a = 1.5   # comment
b = 2*a
k = 1e-2
neg = -a
flag
[other]
g = sec:b^2
h = 2 * (sec:a
   + 1)
hash = 7 # trailing
semi = 8 ; also
[CaseSec]
Val = 9
[ws]
  spaced   =    10
tabbed	=	11
""")
OPTIONS_PROBE_EXPRS = ["sec:a", "sec:b", "sec:k", "sec:neg", "sec:flag", "other:g", "other:h", "other:hash", "other:semi",
                       "CaseSec:Val", "ws:spaced", "ws:tabbed", "casesec:val", "sec:A", "sec:missing"]


# The mesh of tests/golden/ops_small.npz (make_golden.OPS_INI: non-uniform dx, x-dependent dy, full non-orthogonal,
# x- and y-dependent metric): the host layer's Coordinates (covariant metric, J, Bxy, Christoffel contractions G1..G3,
# non-uniform corrections, guard-cell treatment) is compared with the reference's on it.
OPS_MESH_INI = make_ini(12, 8, 16, mxg=2, myg=2, dx="0.2*(1+0.1*x)", dy="0.7+0.05*cos(2*pi*x)", dz="2*pi/16", mesh_extra="""
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
[mesh:ddy]
first = C2
""")


# examples/eigen-box (unchanged eigen-box.cxx: a wave equation, D2DX2 + communicate) with its shipped data/BOUT.inp
# restated without comments; the shipped solver type is slepc (an eigenvalue solver, out of scope), so the run uses
# rk4 with the input's own timestep / output_step / nout.  nz = 1, MXG = 1, mesh sizes given as expressions of MXG.
EIGENBOX_INI = """
MXG = 1
MYG = 0
dump_on_restart = false
[mesh]
length = 1
nx = 32 + 2*MXG
ny = 1
nz = 1
dx = length / (nx - 2*MXG)
dy = 1
[solver]
type = "slepc"
timestep = 0.01
output_step = 0.2
nout = 1
[f]
bndry_all = dirichlet_o2
scale = 1
function = sin(2*pi*x)
[g]
bndry_all = none
scale = 0.0
[input]
error_on_unused_options = false
"""


# examples/IMEX/advection-reaction (unchanged split_operator.cxx: bracket(phi, U, BRACKET_SIMPLE) advection + a reaction
# term, split operator, SAVE_ONCE, initial_profile).  Its shipped input reads the mesh from simple_xz.nc (NetCDF,
# unreadable in this image on either side), so the mesh is given as options here; derivative methods, solver type and
# profiles are the shipped ones; rate is non-zero so that the diffusive half does something; nout reduced.
ADVREACT_INI = """
MZ = 64
periodicX = true
MYG = 0
[mesh]
nx = 68
ny = 1
dx = 1 / 64
dz = 2*pi / 64
[mesh:ddx]
first = C4
second = C4
upwind = W3
flux = U1
[mesh:ddz]
first = C4
second = C4
upwind = W3
flux = U1
[solver]
type = splitrk
timestep = 0.0001
output_step = 0.01
[split]
rate = 0.5
[U]
scale = 1.0
function = sin(4*pi*x) * sin(3*z)
[phi]
scale = 1.0
function = sin(2*pi*x) * sin(z)
[input]
error_on_unused_options = false
"""


# Laplacian types pcr / pcr_thomas, LaplaceXZ cyclic, InvertPar cyclic (tests/golden/make_golden_invert.py).
# 16 interior x points (the reference's pcr needs a power of two), x- and y-dependent metric with g13 = 0 (LaplaceXZ
# cyclic asserts it), y periodic.
INVERT_GRID = (20, 8, 16)
INVERT_LAP_FLAGS = [(0, 0, 0), (0, 3, 0), (0, 0, 3), (0, 16, 16), (0, 32, 32), (1, 0, 0), (0, 1, 2)]
INVERT_XZ_FLAGS = [(0, 0), (3, 0), (0, 3), (16, 16), (32, 32), (1, 2)]
INVERT_LAP_INI = make_ini(20, 8, 16, mxg=2, myg=2, dx="0.2*(1+0.1*x)", dy="0.7+0.05*cos(2*pi*x)", dz="2*pi/16", mesh_extra="""
g11 = 1 + 0.1*x
g22 = 1.1 + 0.02*cos(y)
g33 = 1.2 + 0.05*x
ixseps1 = 20
ixseps2 = 20
""", extra="""
[mesh:ddx]
first = C2
[mesh:ddy]
first = C2
[lap_cyclic]
type = cyclic
[lap_pcr]
type = pcr
[lap_pcr_thomas]
type = pcr_thomas
[parderiv]
type = cyclic
""" + "".join(f"[xz{i}]\ntype = cyclic\ninner_boundary_flags = {a}\nouter_boundary_flags = {b}\n"
              for i, (a, b) in enumerate(INVERT_XZ_FLAGS)))
# the same mesh with open field lines everywhere (y boundaries): InvertPar's boundary rows in the y guard cells
INVERT_OPEN_INI = INVERT_LAP_INI.replace("ixseps1 = 20", "ixseps1 = 0").replace("ixseps2 = 20", "ixseps2 = 0")
