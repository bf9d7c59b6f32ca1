"""A few RK4 steps of the unchanged hasegawa-wakatani-3d/hw.cxx through the C++ host layer (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
h = pkg.host.Host(bench.hw3d_host_ini())
h.model_create("hw3d", "rk4")
h.model_solve(1, (int(sys.argv[1]) if len(sys.argv) > 1 else 3) * bench.DT)
h.stream_sync()
print("rhs calls", h.rhs_calls(), "pool allocations", h.pool_allocations(), "copy-on-write copies", h.cow_copies())
h.close()
