"""Unchanged blob2d (1024 x 1024) / hasegawa-wakatani (256 x 256) through the C++ host layer: wall us per RHS call over
two step counts, launches per RHS.  Run under an ncu launch list to see the GPU time of the same launches.
Usage: python tools/time_host_2d.py [blob2d|hw] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
tag = sys.argv[1] if len(sys.argv) > 1 else "blob2d"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ini, dims = bench.model_2d_ini(tag)
dt = float([ln for ln in ini.splitlines() if ln.startswith("timestep")][0].split("=")[1])
res = []
for n in (steps // 4, steps):      # one Solver::solve per Host (a solver initialises once)
    h = pkg.host.Host(ini)
    h.model_create("blob2d" if tag == "blob2d" else "hw", "rk4")
    h.model_rhs(0.0)
    h.stream_sync()
    c0, l0, w0 = h.rhs_calls(), pkg.capi.launch_count(), h.cow_copies()
    t0 = time.perf_counter()
    h.model_solve(1, n * dt)
    h.stream_sync()
    res.append((time.perf_counter() - t0, h.rhs_calls() - c0, pkg.capi.launch_count() - l0))
    cow = (h.cow_copies() - w0) / max(h.rhs_calls() - c0, 1)
    allocs = h.pool_allocations()
    h.close()
(t1, c1, l1), (t2, c2, l2) = res
print(f"{tag} {dims}: {(t2 - t1) / (c2 - c1) * 1e6:.1f} us per RHS call (slope), {t2 / c2 * 1e6:.1f} us (long run), "
      f"{(l2 - l1) / (c2 - c1):.1f} launches per RHS call, {cow:.2f} copy-on-write block copies per RHS call, pool allocations {allocs}")
