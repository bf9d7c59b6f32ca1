"""Unchanged examples/hasegawa-wakatani-3d/hw.cxx through the C++ host layer at the headline size: us per RHS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
t0 = time.time()
r = bench.host_layer_models(pkg)
print({k: v for k, v in r.items() if "3d" in k}, f"wall {time.time()-t0:.1f} s")
