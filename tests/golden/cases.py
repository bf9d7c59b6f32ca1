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


