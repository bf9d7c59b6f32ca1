#!/usr/bin/env python
"""bench.py -- RHS grid-points/s (fp64) of the BOUT++ physics-model RHS hot path on B200.

Workload (BASELINE.json configs[3], the config the metric "RHS grid-pts/s fp64 at 1/2/4/8 B200" is
quoted on): examples/hasegawa-wakatani-3d at 512x128x512 interior points PER GPU (516x132x512 with
guard cells, MXG = MYG = 2, y periodic, bndry_all = dirichlet_o2).  At N GPUs the BoutMesh decomposition is
NYPE = N by default (weak scaling: global ny = 128*N; --decomp x gives NXPE = N, global nx = 512*N + 4, with the
X-decomposed LaplaceCyclic; --decomp xy gives NXPE = 2).  One "step" is one RK4 time step = 4 RHS evaluations, each
  Solver::pre_rhs boundary conditions -> LaplaceCyclic solve (vort -> phi) -> Mesh::communicate(n, vort, phi)
  -> fused RHS kernel (2 Arakawa brackets, Div_par_Grad_par, DDZ, 2 Delp2) with the RK stage update in
  its epilogue;
value = interior points * 4 * K / time.  Fields (279 MB each) are far larger than the 126 MB L2, so no
L2 flush is needed between iterations (stated in config.l2).  Before anything is timed the arm must pass a
full-size parity gate: one RHS on the reference arm's own 512x16x512 slab state against the reference's output
(<= 1e-10, N = 1), and at N > 1 every rank's slab of a decomposed 512x16Nx512 (or 512Nx16x512) problem against
the same problem computed on one rank (<= 1e-12).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one process per GPU under torchrun)
  python bench.py --workload elmpb [...]                        BASELINE configs[4]: elm-pb 3-field RHS, 512x64x256
                                                                interior points per GPU (shifted metric), RK4
  python bench.py --workload blob2d [...]                       BASELINE configs[2]: blob2d 1024x1x1024 in total,
                                                                X-decomposed over the GPUs (NXPE = N, strong scaling)
The default (hw3d) line also carries "other": elm-pb RHS points/s at 1 GPU, LaplaceCyclic solves/s, the
examples/performance microbenchmarks (BASELINE configs[1], 256x64x256; timed both after an L2 flush and in steady
state over operand sets larger than L2) and the UNCHANGED reference sources (hasegawa-wakatani, blob2d,
hasegawa-wakatani-3d) run through the C++ host layer -- short extra measurements, N = 1 only.
  python bench.py --impl reference [...]                        the reference's own CPU implementation of the
                                                                same RHS (oracle/_ref, host cores, bounded slab)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MXSUB, MYSUB, MZ, MXG, MYG = 512, 128, 512, 2, 2
HW = dict(alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3)
# Explicit RK4 on HW3D is limited by the zonal (kz = 0) parallel-coupling mode: lambda ~ alpha * (4/dy^2) * (Lx/pi)^2,
# i.e. 2.7e5 at 8 GPUs (Lx = 8 * 102.4): dt must stay below 2.8/lambda ~ 1e-5 (tools/dbg_stability.py shows the
# blow-up at dt = 1e-3 for nx >= 768 on a single rank).  The arithmetic per step does not depend on dt.
DT = 2e-6
METRIC = "rhs_grid_points_per_s"
UNIT = "grid-points/s"
WORKLOAD = "hasegawa-wakatani-3d RHS, 512x128x512 interior points per GPU (516x132x512 with guards), RK4"
# reference arm / cpu_baseline: same model and x-z extent, a y-slab of the workload
REF_MY = 16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def profile_traffic(key="k_hw3d_rhs_bytes_per_launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    ncu --set full capture (profiles/*.json), or None."""
    p = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True,
                                         bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        import datetime
        for line in self.proc.stdout:
            line = line.strip()
            try:   # nvidia-smi's own sample time (the pipe may deliver lines late)
                t = datetime.datetime.strptime(line.split(",")[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                t = time.time()
            self.lines.append((t, line))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)
        self.proc.terminate()
        try:
            self.thread.join(timeout=2.0)
        except Exception:
            pass
        sm, smax, reasons, power = [], None, set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")][1:]
            if len(f) < 7 or not (t0 - 0.03 <= t <= t1 + 0.03):
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


def pipelined_e2e(torch, model, names, host_in, host_out, steps, step_fn):
    """End-to-end loop through the public call with HOST buffers: every step uploads its input state from
    pinned memory and downloads its result.  The three legs run on three streams over two device buffer sets, so
    the upload of step k+1 and the download of step k-1 overlap the computation of step k (PCIe is full duplex);
    the bytes moved per step are unchanged.  Returns seconds (device clock, all streams drained)."""
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    bufs = [[torch.empty_like(getattr(model, n)) for n in names] for _ in range(2)]
    keep = [getattr(model, n) for n in names]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [None, None]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def upload(b):
        with torch.cuda.stream(s_in):
            if ev_out[b] is not None:
                s_in.wait_event(ev_out[b])          # the previous result in this buffer set has left the device
            for d, h in zip(bufs[b], host_in):
                d.copy_(h, non_blocking=True)
            ev_in[b].record(s_in)

    torch.cuda.synchronize()
    e0.record(main)
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    upload(0)
    for k in range(steps):
        b = k % 2
        if k + 1 < steps:
            upload((k + 1) % 2)
        main.wait_event(ev_in[b])
        for n, d in zip(names, bufs[b]):
            setattr(model, n, d)
        step_fn()
        ev_done[b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            for h, d in zip(host_out, bufs[b]):
                h.copy_(d, non_blocking=True)
            ev_out[b] = torch.cuda.Event()
            ev_out[b].record(s_out)
    main.wait_stream(s_out)
    main.wait_stream(s_in)
    e1.record(main)
    torch.cuda.synchronize()
    for n, t in zip(names, keep):
        setattr(model, n, t)
    return e0.elapsed_time(e1) * 1e-3


def mixmode(a, seed=0.5):
    """FieldMixmode (src/field/fieldgenerators.cxx:110-167): sum_i cos(i a + phase_i) / (1 + |i-4|)^2."""
    import numpy as np

    def gen_rand(s):
        niter = 11 + (23 + int(round(s))) % 79
        A, B = 0.01, 1.23456789
        x = (A + s % B) / (B + 2 * A)
        for _ in range(niter):
            x = 3.99 * x * (1 - x)
        return x
    out = np.zeros_like(a)
    for i in range(14):
        out += np.cos(i * a + np.pi * (2 * gen_rand(seed + i) - 1)) / (1 + abs(i - 4)) ** 2
    return out


def initial_state(pe_xind, nxpe, pe_yind, nype, torch, device, MXSUB=MXSUB, MYSUB=MYSUB):
    """The shipped initial condition (examples/hasegawa-wakatani-3d/data/BOUT.inp): n = 0,
    vort = 0.1 mixmode(2 pi x) mixmode(z - y), on this rank's slab of the global grid.  Synthetic,
    generated on the host once (setup, not timed)."""
    import numpy as np
    gnx = MXSUB * nxpe
    gx = np.arange(MXSUB + 2 * MXG) + pe_xind * MXSUB       # global x index incl. guards
    X = (0.5 + gx - MXG) / gnx                              # symmetricGlobalX (boutmesh.cxx:3159-3165)
    gy = np.arange(MYSUB + 2 * MYG) + pe_yind * MYSUB       # global y index incl. guards
    Y = 2 * np.pi * (gy - MYG + 0.5) / (MYSUB * nype)
    Z = 2 * np.pi * np.arange(MZ) / MZ
    fx = mixmode(2 * np.pi * X)
    zy = mixmode((Z[None, :] - Y[:, None]))
    vort = 0.1 * fx[:, None, None] * zy[None, :, :]
    return torch.zeros(vort.shape, dtype=torch.float64), torch.from_numpy(np.ascontiguousarray(vort))


def hw3d_metric(torch, g, dev):
    one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device=dev)
    return dict(dx=one(0.2), dy=one(1.0), dz=one(0.2), J=one(1.0), g_22=one(1.0), g11=one(1.0), g33=one(1.0),
                g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0))


def rel_err_interior(torch, got, want, sl):
    return float(((got[sl] - want[sl]).abs().max() / want[sl].abs().max().clamp_min(1e-300)).item())


def parity_gate_reference(torch, pkg, dev):
    """N = 1, before anything is timed: the reference arm's own 512x16x512 slab state through ONE RHS of our arm
    against the reference's Solver::run_rhs output for that state (the unchanged hw.cxx; <= 1e-10 relative)."""
    import tempfile
    import numpy as np
    path = os.path.join(tempfile.mkdtemp(prefix="b200_gate_"), "slab.npz")
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--dump-rhs", path],
                           capture_output=True, text=True, timeout=900,
                           env={k: v for k, v in os.environ.items()
                                if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS", "OMP_PROC_BIND")})
    except Exception as e:  # noqa: BLE001
        return {"ok": None, "skipped": repr(e)[:200]}
    if not os.path.exists(path):
        return {"ok": None, "skipped": ("reference arm gave no slab: " + (p.stdout + p.stderr)[-200:])}
    d = np.load(path)
    capi = pkg.capi
    g = capi.Grid(MXSUB + 2 * MXG, REF_MY + 2 * MYG, MZ, MXG, MYG)
    m = pkg.models.HW3D(g, hw3d_metric(torch, g, dev), device=dev, **HW)
    m.set_state(torch.from_numpy(d["n"]).to(dev), torch.from_numpy(d["vort"]).to(dev))
    m.rhs()
    torch.cuda.synchronize()
    errs = {nm: rel_err_interior(torch, f, torch.from_numpy(d["ddt_" + nm]).to(dev), g.interior)
            for nm, f in (("n", m.ddt_n), ("vort", m.ddt_vort))}
    tol = 1e-10
    return {"ok": bool(max(errs.values()) <= tol), "tolerance": tol, "rel_err": errs,
            "what": f"one RHS on the reference arm's 512x{REF_MY}x512 slab state vs the reference's Solver::run_rhs"}


def parity_gate_decomposed(torch, pkg, dev, rank, nxpe, nype):
    """N > 1, before anything is timed: this rank's slab of ONE RHS of a decomposed 512*NXPE x 16*NYPE x 512
    problem (same kernels, exchanges and X-decomposed solve as the workload) against the same problem computed on
    one rank by the same GPU (<= 1e-12 relative)."""
    capi, models = pkg.capi, pkg.models
    my = REF_MY
    mesh = pkg.mesh.BoutMesh(nxpe=nxpe, nype=nype, rank=rank, mxsub=MXSUB, mysub=my, mz=MZ, mxg=MXG, myg=MYG,
                             periodic_x=False, periodic_y=True, device=dev)
    pe_x, pe_y = rank % nxpe, rank // nxpe
    g = capi.Grid(MXSUB + 2 * MXG, my + 2 * MYG, MZ, MXG, MYG)
    md = models.HW3D(g, hw3d_metric(torch, g, dev), device=dev, mesh=mesh, **HW)
    n0, v0 = initial_state(pe_x, nxpe, pe_y, nype, torch, dev, MYSUB=my)
    md.set_state(n0.to(dev), v0.to(dev))
    md.rhs()
    G = capi.Grid(MXSUB * nxpe + 2 * MXG, my * nype + 2 * MYG, MZ, MXG, MYG)
    mg = models.HW3D(G, hw3d_metric(torch, G, dev), device=dev, **HW)
    N0, V0 = initial_state(0, 1, 0, 1, torch, dev, MXSUB=MXSUB * nxpe, MYSUB=my * nype)
    mg.set_state(N0.to(dev), V0.to(dev))
    mg.rhs()
    torch.cuda.synchronize()
    sl = (slice(MXG + pe_x * MXSUB, MXG + (pe_x + 1) * MXSUB), slice(MYG + pe_y * my, MYG + (pe_y + 1) * my), slice(None))
    errs = {}
    for nm, fd, fg in (("n", md.ddt_n, mg.ddt_n), ("vort", md.ddt_vort, mg.ddt_vort)):
        want = fg[sl]
        errs[nm] = float(((fd[g.interior] - want).abs().max() / want.abs().max().clamp_min(1e-300)).item())
    worst = torch.tensor([max(errs.values())], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(worst, op=torch.distributed.ReduceOp.MAX)
    tol = 1e-12
    del md, mg
    torch.cuda.empty_cache()
    return {"ok": bool(worst.item() <= tol), "tolerance": tol, "rel_err_max_over_ranks": float(worst.item()),
            "what": f"one RHS of a {MXSUB * nxpe}x{my * nype}x{MZ} problem decomposed NXPE={nxpe} NYPE={nype} vs the same "
                    f"problem on one rank"}



def bind_to_gpu_numa_node(torch, local):
    """Best effort, N > 1 only: run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host
    buffer is allocated (first touch then places the e2e staging buffers next to that GPU's PCIe root instead of on
    whichever node the launcher started the process).  Returns what was done, for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return {"gpu": bdf, "numa_node": node, "nodes": len(nodes), "bound": False}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu": bdf, "numa_node": node, "nodes": len(nodes), "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "numa_node": node, "nodes": len(nodes), "bound": True, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001 -- sysfs not visible in this container, old torch, ...
        return {"bound": False, "why": repr(e)[:120]}


def run_ours(args):
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun as the driver would
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__),
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--decomp", args.decomp]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bout-dev_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if (world > 1 and os.environ.get("BOUT_B200_NUMA_BIND", "1") != "0") else None
    pkg = ge.load_package()
    capi, models = pkg.capi, pkg.models
    mesh = None
    # BoutMesh decomposition of the weak-scaled global grid (mesh:NXPE / mesh:NYPE of BOUT.inp): "y" splits
    # the parallel direction (the tridiagonal solves stay rank-local), "x" the radial one (X-decomposed
    # LaplaceCyclic: reduce + all-gather + finish), "xy" uses NXPE = 2.
    nxpe, nype = {"x": (world, 1), "y": (1, world), "xy": (2 if world % 2 == 0 else 1, world // 2 if world % 2 == 0 else world)}[args.decomp]
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        mesh = pkg.mesh.BoutMesh(nxpe=nxpe, nype=nype, rank=rank, mxsub=MXSUB, mysub=MYSUB, mz=MZ, mxg=MXG, myg=MYG,
                                 periodic_x=False, periodic_y=True, device=dev)
    pe_xind, pe_yind = rank % nxpe, rank // nxpe

    # ---- parity gate (nothing is timed before it passes) ------------------------------------------
    if args.no_parity_gate:
        gate = {"ok": None, "skipped": "--no-parity-gate"}
    elif world == 1:
        gate = parity_gate_reference(torch, pkg, dev)
    else:
        gate = parity_gate_decomposed(torch, pkg, dev, rank, nxpe, nype)
    if gate.get("ok") is False:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "error": "parity gate failed", "parity_gate": gate}), flush=True)
        raise SystemExit(3)

    g = capi.Grid(MXSUB + 2 * MXG, MYSUB + 2 * MYG, MZ, MXG, MYG)
    metric = hw3d_metric(torch, g, dev)
    model = models.HW3D(g, metric, device=dev, mesh=mesh, **HW)
    n0, v0 = initial_state(pe_xind, nxpe, pe_yind, nype, torch, dev)
    hn, hv = n0.pin_memory(), v0.pin_memory()
    hn_out, hv_out = torch.empty_like(hn).pin_memory(), torch.empty_like(hv).pin_memory()
    model.set_state(hn.to(dev), hv.to(dev))
    npts = MXSUB * MYSUB * MZ

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        model.rk4_step(DT)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    model.kernel_events = []
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        model.rk4_step(DT)
    e1.record()
    barrier()
    w1 = time.time()
    launches = capi.launch_count() - l0
    secs = reduce_max(e0.elapsed_time(e1) * 1e-3)
    kev = model.kernel_events
    model.kernel_events = None
    k_secs = sum(a.elapsed_time(b) for a, b in kev) * 1e-3 / len(kev)
    finite = bool(torch.isfinite(model.n).all().item() and torch.isfinite(model.vort).all().item())

    # ---- end to end through the public call with HOST buffers ------------------------------------
    e2e_steps = args.steps
    model.set_state(hn.to(dev), hv.to(dev))
    barrier()
    e2e_local = pipelined_e2e(torch, model, ("n", "vort"), (hn, hv), (hn_out, hv_out), e2e_steps,
                              lambda: model.rk4_step(DT))
    barrier()
    e2e_secs = reduce_max(e2e_local)
    w2 = time.time()
    # the downloaded result of the last e2e step must be what one device-resident step gives from the same input
    model.set_state(hn.to(dev), hv.to(dev))
    model.rk4_step(DT)
    torch.cuda.synchronize()
    # (interior points: the guard cells of an output state are not part of it -- the next right-hand side's boundary
    # conditions / exchange set them, and whether the last kernel has already written them depends on where the
    # output buffer lives)
    I = g.interior
    e2e_ok = bool(torch.equal(model.n.cpu()[I], hn_out[I]) and torch.equal(model.vort.cpu()[I], hv_out[I]))
    clocks = sampler.stop(w0, w2) if rank == 0 else None   # samples over both timed regions
    field_bytes = hn.numel() * 8

    # ---- per-phase breakdown of one RHS (separate short run; CUDA events between phases) -------------
    model.phase_events = []
    for _ in range(3):
        model.rk4_step(DT)
    torch.cuda.synchronize()
    phases = model.phase_summary()
    model.phase_events = None
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        return
    peak, peak_src = peaks()
    # fused kernel, compulsory bytes per interior point per launch (DESIGN.md "k_hw3d_rhs"):
    #   reads n, vort, phi (24) + start n, vort (16; in stage 1 they ARE the input rows: 0) + running combination
    #   in (16, stages 2-4); writes next stage state (16, stages 1-3) + running combination out (16) -> 56, 88, 88, 72
    alg_bytes = (56 + 88 + 88 + 72) / 4.0 * npts
    achieved = alg_bytes / k_secs / 1e9
    value = npts * 4 * args.steps * world / secs
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (shipped closed-form initial condition, mixmode)",
        "config": {"workload": WORKLOAD, "decomposition": f"NXPE={nxpe} NYPE={nype}", "rhs_per_step": 4,
                   "points_per_gpu": npts, "dt": DT,
                   "l2": "inputs larger than L2 (each field 279 MB vs 126 MB L2); no flush needed"},
        "gpu_launches": int(launches),
        "finite": finite,
        "e2e": {"value": npts * 4 * e2e_steps * world / e2e_secs, "unit": UNIT,
                "h2d_bytes_per_step": 2 * field_bytes, "d2h_bytes_per_step": 2 * field_bytes,
                "ms_per_step": e2e_secs / e2e_steps * 1e3, "result_equals_device_resident_step": e2e_ok,
                "host_numa_binding": numa, "note": "state (n, vort) copied host->device from pinned memory and the result device->host every RK4 step; "
                        "uploads / computation / downloads of neighbouring steps overlap on three streams (double-buffered)"},
        "roofline": {"bound": "hbm", "kernel": "k_hw3d_rhs<2> (fused RHS + RK stage update)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": profile_traffic(),
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_secs * 1e3,
                     "kernel_share_of_step": 4 * k_secs / (secs / args.steps),
                     "rhs_algorithmic_GBps": (56.0 + 64.0) * npts * 4 * args.steps / secs / 1e9,
                     "rhs_bytes_per_point": "56 (solve 16 + fused 40, SURVEY 8d) + 64 RK stage update",
                     "rhs_frac_of_hbm_peak": (56.0 + 64.0) * npts * 4 * args.steps / secs / 1e9 / peak,
                     "rhs_frac_of_survey_ceiling": 56.0 * npts * 4 * args.steps / secs / 1e9 / peak,
                     "survey_ceiling_points_per_s": peak * 1e9 / 56.0},
        "parity_gate": gate,
        "clocks": clocks,
        "rhs_phase_ms": {k: round(v, 4) for k, v in phases.items()},
    }
    if world == 1 and not args.no_other:
        del model
        torch.cuda.empty_cache()
        try:
            out["other"] = other_measurements(torch, pkg, dev, peak, with_cpu=not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001 -- the extras never invalidate the main line
            out["other"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(steps=3, warmup=1)
    print(json.dumps(out), flush=True)



ELM = dict(nx=516, ny=64, nz=256)   # BASELINE configs[4]: 516 x 64 x 256 (nx with guard cells)
ELM_DT = 1e-3
ELM_WORKLOAD = ("elm-pb linear 3-field reduced-MHD RHS (shipped [highbeta] switches), 512x64x256 interior points per "
                "GPU (516x68x256 with guards), synthetic shifted-metric tokamak grid, zperiod 15, RK4")
# Algorithmic bytes per interior point of the fused kernel k_elmpb_rhs: reads U, P, Psi, phi, Jpar
# (5 x 8) + the parallel slices of Psi and phi (32), writes the three time derivatives (24)
ELM_FUSED_BPP = 96.0
# inside an RK4 step three of the four launches also read the step's start state and write the next stage state
# (3 fields x 16 B): (3 * 144 + 96) / 4
ELM_FUSED_BPP_RK4 = (3 * 144.0 + 96.0) / 4
# Whole elm-pb RHS pass list: slices of Psi, phi 2 x 24, row modes of U, P, Jpar 3 x 8,
# solve 16, phi update 24, Delp2(Psi) 16, smooth_x 16, fused 96
ELM_RHS_BPP = 2 * 24 + 3 * 8 + 16 + 24 + 16 + 16 + 96   # = 240; + 36 per RHS of RK stage traffic in a step


def event_time(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def elmpb_model(pkg, dev, mesh=None, nype=1):
    """Weak scaling: the global grid has 64 * NYPE y points, each rank keeps a 512 x 64 x 256 slab."""
    m = pkg.models.ELMpb.from_setup(pkg.elm_setup.elmpb_setup(ELM["nx"], ELM["ny"] * nype, ELM["nz"]), device=dev,
                                    mesh=mesh)
    m.set_state(*pkg.models.elmpb_initial_state(m.g, device=dev))
    return m


def other_measurements(torch, pkg, dev, peak, with_cpu=False):
    """Short N = 1 extras for the default line: elm-pb RHS, LaplaceCyclic solves/s, configs[1] microbenchmarks."""
    capi = pkg.capi
    out = {}
    # -- elm-pb RHS (configs[4])
    m = elmpb_model(pkg, dev)
    t = event_time(torch, m.rhs, reps=10)
    m.kernel_events = []
    for _ in range(5):
        m.rhs()
    torch.cuda.synchronize()
    kt = sum(a.elapsed_time(b) for a, b in m.kernel_events) * 1e-3 / len(m.kernel_events)
    m.kernel_events = None
    npts = m.g.n_interior
    out["elmpb_rhs"] = {"workload": ELM_WORKLOAD, "grid_points_per_s": npts / t, "ms_per_rhs": t * 1e3,
                        "rhs_algorithmic_GBps": ELM_RHS_BPP * npts / t / 1e9,
                        "rhs_frac_of_hbm_peak": ELM_RHS_BPP * npts / t / 1e9 / peak,
                        "fused_kernel": {"name": "k_elmpb_rhs", "ms": kt * 1e3,
                                         "achieved_GBps": ELM_FUSED_BPP * npts / kt / 1e9,
                                         "frac": ELM_FUSED_BPP * npts / kt / 1e9 / peak},
                        "finite": bool(all(torch.isfinite(f).all().item() for f in (m.ddt_U, m.ddt_P, m.ddt_Psi)))}
    del m
    # -- LaplaceCyclic solves/s (metric: "LaplaceCyclic solves/s"), 16 B/pt algorithmic
    lap = {}
    for name, dims in (("512x128x512", (516, 132, 512, 2, 2)), ("512x64x256", (516, 68, 256, 2, 2)),
                       ("1024x1x1024", (1028, 1, 1024, 2, 0))):
        g = capi.Grid(*dims)
        one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device=dev)
        md = dict(dx=one(0.2), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0),
                  g_22=one(1.0))
        f = torch.rand(g.shape3, dtype=torch.float64, device=dev) - 0.5
        o = torch.zeros_like(f)
        L = capi.Laplace(g, md, 0.2 * g.nz)
        t = event_time(torch, lambda: L.solve(f, out=o), reps=20)
        L.status()
        lap[name] = {"solves_per_s": 1.0 / t, "us_per_solve": t * 1e6, "algorithmic_GBps": 16.0 * g.n_interior / t / 1e9,
                     "frac_of_hbm_peak": 16.0 * g.n_interior / t / 1e9 / peak}
        del L, f, o
    out["laplace_cyclic"] = lap
    # -- examples/performance microbenchmarks on a 256x64x256 Field3D (configs[1]), timed two ways:
    #    "us"         one launch between two CUDA events after flush.zero_() of a 320 MB buffer (L2 flushed by WRITING:
    #                 the write-back of the flush's dirty lines and the launch / event latency land inside the ~20 us
    #                 timed region -- torch's own copy_ of one field reaches 60 % of the peak this way);
    #    "us_steady"  8 distinct operand sets (8 x 4 fields x 36 MB = 1.2 GB, inputs larger than L2) used round-robin,
    #                 80 launches back to back between two events: the operator's throughput as it runs inside an RHS
    #                 (torch copy_: 77 %).
    g = capi.Grid(260, 68, 256, 2, 2)
    nset, nrep = 8, 80
    sets = [tuple(torch.rand(g.shape3, dtype=torch.float64, device=dev) for _ in range(3))
            + (torch.zeros(g.shape3, dtype=torch.float64, device=dev),) for _ in range(nset)]
    dx = torch.full(g.shape2, 0.2, dtype=torch.float64, device=dev)
    flush = torch.empty(40 * 1024 * 1024, dtype=torch.float64, device=dev)   # 320 MB > L2

    def flushed(fn, reps=10):
        fn(*sets[0])
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*sets[0])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    def steady(fn):
        for st in sets:
            fn(*st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nrep):
            fn(*sets[i % nset])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / nrep
    micro = {}
    full = 16.0 * g.nx * g.ny / ((g.nx - 4) * (g.ny - 4))    # copy_ moves the guard cells too
    for name, fn, bpp in (
            ("DDX_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "X", a, scale1=dx, out=o), 16),
            ("DDX_C4", lambda a, b, c, o: capi.stencil(g, "DDX_C4", "X", a, scale1=dx, out=o), 16),
            ("DDX_W3", lambda a, b, c, o: capi.stencil(g, "DDX_W3", "X", a, scale1=dx, out=o), 16),
            ("DDY_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "Y", a, scale1=dx, out=o), 16),
            ("DDZ_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "Z", a, scale1=dx, out=o), 16),
            ("DDZ_FFT", lambda a, b, c, o: capi.ddz_fft(g, a, out=o), 16),
            ("VDDX_U1", lambda a, b, c, o: capi.stencil(g, "VDDX_U1", "X", a, v=b, scale1=dx, out=o), 24),
            ("bracket_arakawa", lambda a, b, c, o: capi.bracket_arakawa(g, a, b, dx, dx, out=o), 24),
            ("arithmetic_2a_plus_bc", lambda a, b, c, o: capi.elementwise(g, [(2.0, a), (1.0, b, c)], out=o), 32),
            ("calibration_torch_copy", lambda a, b, c, o: o.copy_(a), full)):
        t, ts = flushed(fn), steady(fn)
        micro[name] = {"us": t * 1e6, "grid_points_per_s": g.n_interior / t,
                       "algorithmic_GBps": bpp * g.n_interior / t / 1e9, "frac_of_hbm_peak": bpp * g.n_interior / t / 1e9 / peak,
                       "us_steady": ts * 1e6, "frac_of_hbm_peak_steady": bpp * g.n_interior / ts / 1e9 / peak}
    out["microbench_256x64x256"] = micro
    del sets, flush
    torch.cuda.empty_cache()
    # -- UNCHANGED reference model sources through the C++ host layer (bout-dev_b200/host: BOUT++'s own API on the
    #    C-ABI): examples/blob2d scaled to 1024 x 1024 (configs[2]) and examples/hasegawa-wakatani as shipped
    #    (configs[0], 256 x 256, periodic X); wall clock around Solver::solve, all host overheads included
    try:
        out["host_layer_models"] = host_layer_models(pkg, with_cpu=with_cpu)
    except Exception as e:  # noqa: BLE001
        out["host_layer_models"] = {"error": repr(e)[:300]}
    # -- SURVEY 8(f)2/(f)3 through the host layer: the other explicit integrators and the output step on blob2d
    #    1024 x 1024, in a child process with a time limit (tools/bench_host_extras.py)
    try:
        import subprocess
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_host_extras.py")], capture_output=True,
                               text=True, timeout=120)
            stdout, stderr = p.stdout, p.stderr
        except subprocess.TimeoutExpired as te:   # keep what the child had finished
            stdout = te.stdout.decode() if isinstance(te.stdout, bytes) else (te.stdout or "")
            stderr = "time limit"
        lines = [ln for ln in stdout.splitlines() if ln.startswith("{") and ln.endswith("}")]
        out["host_layer_extras"] = json.loads(lines[-1]) if lines else {"error": (stderr or "no output")[-300:]}
    except Exception as e:  # noqa: BLE001
        out["host_layer_extras"] = {"error": repr(e)[:300]}
    return out


def host_layer_models(pkg, with_cpu=False):
    """Wall time of Solver::solve for two step counts (after a discarded warm-up instance): the slope is the
    per-RHS cost with every host overhead of the C++ layer in it; the intercept is start-up."""
    import numpy as np
    res = {}
    for tag, model, ini, steps in (
            ("blob2d_1024x1x1024", "blob2d", model_2d_ini("blob2d")[0], (5, 20, 60)),
            ("hasegawa-wakatani_256x1x256", "hw", model_2d_ini("hw")[0], (5, 50, 150)),
            # the headline workload as a drop-in: UNCHANGED examples/hasegawa-wakatani-3d/hw.cxx, its BOUT_FOR_RAJA
            # single-index loop recorded by the host layer and dispatched to bout_b200_hw3d_rhs (the RK4 stage updates
            # are the C++ solver's own kernels here, not the fused epilogue of the harness the headline number uses)
            ("hasegawa-wakatani-3d_512x128x512", "hw3d", hw3d_host_ini(), (2, 4, 12))):
        dt = float([ln for ln in ini.splitlines() if ln.startswith("timestep")][0].split("=")[1])
        pts = []
        for nsteps in steps:
            h = pkg.host.Host(ini)
            try:
                names = h.model_create(model, "rk4")
                h.model_rhs(0.0)
                h.model_get(names[0])
                c0 = h.rhs_calls()
                t0 = time.time()
                h.model_solve(1, nsteps * dt)
                h.stream_sync()                           # the solve has completed on the device
                t1 = time.time()
                final = h.model_get(names[0])
                pts.append((h.rhs_calls() - c0, t1 - t0, bool(np.isfinite(final).all()), h.pool_allocations()))
                nx, ny, nz = h.shape3
            finally:
                h.close()
        # (the first instance is a discarded warm-up: kernel modules and the device pool fill during it)
        (n1, t1, _, _), (n2, t2, fin, pool) = pts[-2], pts[-1]
        per = (t2 - t1) / (n2 - n1)
        npts = (nx - 4) * (ny - 4 if model == "hw3d" else max(ny, 1)) * nz
        res[tag] = {"us_per_rhs": per * 1e6, "grid_points_per_s": npts / per, "rhs_calls": [n1, n2],
                    "wall_ms": [t1 * 1e3, t2 * 1e3], "finite": fin, "steady_state_pool_allocations": pool}
        if with_cpu:   # the reference's own CPU run of the same unchanged source, beside it (cpu_baseline leg)
            res[tag]["cpu_baseline"] = cpu_baseline(steps=2, warmup=0, workload=model)
    return res


def run_elmpb(args):
    """--workload elmpb: BASELINE configs[4].  N > 1 splits the parallel (y) direction over the GPUs (NXPE = 1,
    NYPE = N): guard rows travel by NCCL send/recv, everything else stays rank-local."""
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__), "--gpus",
                   str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", "elmpb"]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bout-dev_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg = ge.load_package()
    capi = pkg.capi
    mesh = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        mesh = pkg.mesh.BoutMesh(nxpe=1, nype=world, rank=rank, mxsub=ELM["nx"] - 4, mysub=ELM["ny"], mz=ELM["nz"],
                                 mxg=2, myg=2, periodic_x=False, periodic_y=True, device=dev)
    m = elmpb_model(pkg, dev, mesh=mesh, nype=world)
    host = [f.cpu().pin_memory() for f in (m.U, m.P, m.Psi)]
    host_out = [torch.empty_like(h).pin_memory() for h in host]
    npts = m.g.n_interior

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        m.rk4_step(ELM_DT)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    m.kernel_events = []
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        m.rk4_step(ELM_DT)
    e1.record()
    barrier()
    launches = capi.launch_count() - l0
    secs = reduce_max(e0.elapsed_time(e1) * 1e-3)
    kev = m.kernel_events
    m.kernel_events = None
    k_secs = sum(a.elapsed_time(b) for a, b in kev) * 1e-3 / len(kev)
    finite = bool(all(torch.isfinite(f).all().item() for f in (m.U, m.P, m.Psi)))
    # end to end: state from pinned host memory every step, result back
    barrier()
    e2e_local = pipelined_e2e(torch, m, ("U", "P", "Psi"), host, host_out, args.steps, lambda: m.rk4_step(ELM_DT))
    barrier()
    e2e_secs = reduce_max(e2e_local)
    w2 = time.time()
    m.set_state(*[h.to(dev) for h in host])
    m.rk4_step(ELM_DT)
    torch.cuda.synchronize()
    e2e_ok = bool(all(torch.equal(f.cpu(), h) for f, h in zip((m.U, m.P, m.Psi), host_out)))
    clocks = sampler.stop(w0, w2) if rank == 0 else None
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        return
    peak, peak_src = peaks()
    fb = host[0].numel() * 8
    achieved = ELM_FUSED_BPP_RK4 * npts / k_secs / 1e9
    out = {
        "metric": METRIC, "value": npts * 4 * args.steps * world / secs, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (closed-form tokamak grid and perturbation)",
        "config": {"workload": ELM_WORKLOAD, "decomposition": f"NXPE=1 NYPE={world}",
                   "rhs_per_step": 4, "points_per_gpu": npts, "dt": ELM_DT,
                   "l2": "the RHS streams ~25 fields of 72 MB each (1.8 GB per RHS) through the 126 MB L2; no flush needed"},
        "gpu_launches": int(launches), "finite": finite,
        "e2e": {"value": npts * 4 * args.steps * world / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": 3 * fb,
                "d2h_bytes_per_step": 3 * fb, "ms_per_step": e2e_secs / args.steps * 1e3,
                "result_equals_device_resident_step": e2e_ok,
                "note": "state (U, P, Psi) copied host->device from pinned memory and the result device->host every RK4 step; "
                        "uploads / computation / downloads of neighbouring steps overlap on three streams (double-buffered)"},
        "roofline": {"bound": "hbm", "kernel": "k_elmpb_rhs (fused 3-equation RHS + z filter + RK stage update)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "traffic": profile_traffic("k_elmpb_rhs_bytes_per_launch"),
                     "algorithmic_bytes_per_launch": ELM_FUSED_BPP_RK4 * npts, "kernel_ms": k_secs * 1e3,
                     "kernel_share_of_step": 4 * k_secs / (secs / args.steps),
                     "rhs_algorithmic_GBps": (ELM_RHS_BPP + 36.0 + 36.0) * npts * 4 * args.steps / secs / 1e9,
                     "rhs_bytes_per_point": f"{ELM_RHS_BPP:.0f} (pass list in DESIGN.md section 3) + 36 RK stage states "
                                            f"+ 36 final combination (3 fields x 48 B per step / 4)"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(steps=2, warmup=1, workload="elmpb")
    print(json.dumps(out), flush=True)


BLOB_WORKLOAD = ("examples/blob2d RHS (shipped switches: boussinesq, sheath; Arakawa brackets, Delp2 through FFTs, LaplaceCyclic), "
                 "1024x1x1024 interior points in total (1028x1x1024 with guards), X-decomposed over the GPUs (NXPE = N), RK4")
BLOB = {"mx": 1024, "nz": 1024, "dt": 0.05}


def run_blob2d(args):
    """--workload blob2d: BASELINE configs[2], strong scaling -- the 1024 x 1 x 1024 grid is split in X over the N
    GPUs (NXPE = N): X guard planes of n, omega and phi through Mesh::communicate (NCCL send/recv in place), the
    tridiagonal solves through reduce -> all-gather of the interface rows -> finish.  The whole problem is 8 MB per
    field: L2-resident and launch-latency-bound at every N, so no HBM fraction is claimed."""
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.abspath(__file__), "--gpus",
                   str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", "blob2d"]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bout-dev_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg = ge.load_package()
    capi = pkg.capi
    mxsub = BLOB["mx"] // world
    mesh = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        mesh = pkg.mesh.BoutMesh(nxpe=world, nype=1, rank=rank, mxsub=mxsub, mysub=1, mz=BLOB["nz"], mxg=2, myg=0,
                                 periodic_x=False, periodic_y=True, device=dev)
    g = mesh.grid if mesh is not None else capi.Grid(mxsub + 4, 1, BLOB["nz"], 2, 0)
    one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device=dev)
    md = dict(dx=one(0.3), dy=one(1.0), dz=one(0.3), J=one(1.0), g_22=one(1.0), g11=one(1.0), g33=one(1.0), g13=one(0.0),
              G1=one(0.0), G3=one(0.0), d1_dx=one(0.0), g_11=one(1.0))
    m = pkg.models.Blob2D(g, md, Te0=5.0, B0=0.35, D_n=1e-6, D_vort=1e-6, R_c=1.5, L_par=10.0, device=dev, mesh=mesh)
    m.set_state(*pkg.models.blob2d_initial_state(g, mesh=mesh, device=dev, height=0.5, width=0.1))
    host = [f.cpu().pin_memory() for f in (m.n, m.omega)]
    host_out = [torch.empty_like(h).pin_memory() for h in host]
    npts = BLOB["mx"] * BLOB["nz"]          # whole problem (strong scaling)
    dt = BLOB["dt"]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        m.rk4_step(dt)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        m.rk4_step(dt)
    e1.record()
    barrier()
    launches = capi.launch_count() - l0
    secs = reduce_max(e0.elapsed_time(e1) * 1e-3)
    finite = bool(all(torch.isfinite(f).all().item() for f in (m.n, m.omega)))
    barrier()
    e2e_local = pipelined_e2e(torch, m, ("n", "omega"), host, host_out, args.steps, lambda: m.rk4_step(dt))
    barrier()
    e2e_secs = reduce_max(e2e_local)
    w2 = time.time()
    clocks = sampler.stop(w0, w2) if rank == 0 else None
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        return
    fb = host[0].numel() * 8
    out = {
        "metric": METRIC, "value": npts * 4 * args.steps / secs, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (shipped closed-form initial condition: a gaussian blob)",
        "config": {"workload": BLOB_WORKLOAD, "decomposition": f"NXPE={world} NYPE=1", "rhs_per_step": 4,
                   "points_total": npts, "points_per_gpu": npts // world, "dt": dt, "us_per_rhs": secs / args.steps / 4 * 1e6,
                   "l2": "the whole problem (8 MB per field) is L2-resident: latency-bound, no HBM fraction claimed"},
        "gpu_launches": int(launches), "finite": finite,
        "e2e": {"value": npts * 4 * args.steps / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": 2 * fb,
                "d2h_bytes_per_step": 2 * fb, "ms_per_step": e2e_secs / args.steps * 1e3,
                "note": "this rank's slab of (n, omega) copied host->device from pinned memory and the result device->host "
                        "every RK4 step (double-buffered, three streams); bytes are per rank"},
        "roofline": None,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(steps=2, warmup=1, workload="blob2d")
    print(json.dumps(out), flush=True)


# ---- reference arm: the reference's own CPU implementation (oracle/_ref/libboutref.so) ---------------
def ref_ini():
    import __graft_entry__ as ge
    orc = ge.load_oracle()
    nx = MXSUB + 2 * MXG
    return orc, orc.make_ini(nx, REF_MY, MZ, mxg=MXG, myg=MYG, mesh_extra=f"ixseps1 = {nx}\nixseps2 = {nx}\n", extra=f"""
[hw]
alpha = {HW['alpha']}
kappa = {HW['kappa']}
Dvort = {HW['Dvort']}
Dn = {HW['Dn']}
[all]
scale = 0.0
bndry_all = dirichlet_o2
[vort]
scale = 0.1
function = mixmode(2*pi*x) * mixmode(z - y)
[solver]
timestep = {DT}
[input]
error_on_unused_options = false
""")


def run_reference_elmpb(args, cores):
    """Reference arm of --workload elmpb: the UNCHANGED examples/elm-pb/elm_pb.cxx through the reference's
    Solver::run_rhs on a y-slab (512 x 8 x 256) of the workload's synthetic grid."""
    import __graft_entry__ as ge
    orc = ge.load_oracle()
    if not os.path.exists(orc.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libboutref.so not built"}))
        return
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from cases import elmpb_ini
    my = 8
    r = orc.Ref(elmpb_ini(ELM["nx"], my, ELM["nz"], timestep=str(ELM_DT)))
    r.model_create("elmpb", "rk4")
    r.model_rhs(0.0)
    npts = (ELM["nx"] - 4) * my * ELM["nz"]
    if args.warmup > 0:
        r.model_time_rhs(reps=args.warmup)
    reps = max(4 * args.steps, 4)
    secs = r.model_time_rhs(reps=reps) * reps
    value = npts * reps / secs
    threads = r.omp_threads()
    sample = (f"unchanged examples/elm-pb/elm_pb.cxx through the reference's Solver::run_rhs, y-slab 512x{my}x256 of the "
              f"512x64x256 workload, {reps} run_rhs, OpenMP {threads} threads (omp_get_max_threads; {cores} host cores), "
              f"single MPI rank (shim)")
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample, "same_config": False,
          "host_cores": cores}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / reps * 4 * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (closed-form tokamak grid and perturbation)",
           "config": {"workload": ELM_WORKLOAD, "rhs_per_step": 4, "sample_points": npts},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0)


def model_2d_ini(tag):
    """BOUT.inp of the two 2-D model cases timed through the host layer (and through the reference beside them)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    if tag == "blob2d":
        extra = cases.BLOB_INI[cases.BLOB_INI.index("[mesh:ddx]"):]
        return cases.make_ini(1028, 1, 1024, mxg=2, myg=0, dx="0.3", dz="0.3", extra=extra), (1024, 1, 1024)
    extra = cases.HW_INI[cases.HW_INI.index("[hw]"):]
    return cases.make_ini(260, 1, 256, mxg=2, myg=0, periodic_x=True, extra=extra), (256, 1, 256)


def hw3d_host_ini():
    """BOUT.inp of the headline workload for the unchanged hw.cxx through the host layer (shipped options of
    examples/hasegawa-wakatani-3d/data/BOUT.inp, grid of BASELINE configs[3])."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    extra = cases.HW3D_INI[cases.HW3D_INI.index("[hw]"):].replace("timestep = 0.01", f"timestep = {DT}")
    return cases.make_ini(MXSUB + 2 * MXG, MYSUB, MZ, mxg=MXG, myg=MYG,
                          mesh_extra=f"ixseps1 = {MXSUB + 2 * MXG}\nixseps2 = {MXSUB + 2 * MXG}\n", extra=extra)


def run_reference_2d(args, cores):
    """Reference arm for the 2-D model cases of the default line's extras: the unchanged example source through
    the reference's Solver::run_rhs at the full grid (blob2d 1024 x 1024, hasegawa-wakatani 256 x 256)."""
    import __graft_entry__ as ge
    orc = ge.load_oracle()
    if not os.path.exists(orc.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libboutref.so not built"}))
        return
    ini, (mx, my, mz) = model_2d_ini(args.workload)
    r = orc.Ref(ini)
    r.model_create("blob2d" if args.workload == "blob2d" else "hw", "rk4")
    dt = float([ln for ln in ini.splitlines() if ln.startswith("timestep")][0].split("=")[1])
    r.model_solve(0, dt)     # Solver::init + first run_rhs (the reference's own start-up)
    reps = max(4 * args.steps, 4)
    secs = r.model_time_rhs(reps=reps)
    value = mx * my * mz / secs
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "us_per_rhs": secs * 1e6,
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": r.omp_threads(), "kind": "reference",
                                       "sample": f"unchanged examples/{args.workload} source, {mx}x{my}x{mz}, {reps} run_rhs, "
                                                 f"OpenMP {r.omp_threads()} threads"}}), flush=True)
    sys.stdout.flush()
    os._exit(0)


def run_reference(args, as_baseline=False):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that.  It runs on rank 0 only
    # (one process, every host core of the box, whatever N is) -- the reference build here is single-rank (MPI shim).
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_PROC_BIND", None)
    if args.workload == "elmpb":
        return run_reference_elmpb(args, cores)
    if args.workload in ("blob2d", "hw"):
        return run_reference_2d(args, cores)
    orc, ini = ref_ini()
    kind = "reference" if os.path.exists(orc.REF_SO) else None
    if kind is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libboutref.so not built"}))
        return
    r = orc.Ref(ini)
    names = r.model_create("hw3d", "rk4")
    r.model_solve(0, DT)     # Solver::init + first run_rhs (the reference's own start-up)
    threads = r.omp_threads()
    if args.dump_rhs:        # state + one Solver::run_rhs of it, for the parity gate of our arm
        import numpy as np
        state = {nm: r.model_get(nm) for nm in names}
        r.model_rhs(0.0)
        np.savez(args.dump_rhs, **state, **{"ddt_" + nm: r.model_ddt(nm) for nm in names})
        print(json.dumps({"impl": "reference", "dumped": args.dump_rhs, "vars": names}), flush=True)
        sys.stdout.flush()
        os._exit(0)
    npts = MXSUB * REF_MY * MZ
    if args.warmup > 0:
        r.model_time_rhs(reps=4 * args.warmup)
    secs = r.model_time_rhs(reps=4 * args.steps) * 4 * args.steps
    value = npts * 4 * args.steps / secs
    sample = (f"unchanged examples/hasegawa-wakatani-3d/hw.cxx through the reference's Solver::run_rhs, "
              f"y-slab 512x{REF_MY}x512 of the 512x128x512 workload, 4 run_rhs per step, OpenMP {threads} threads "
              f"(omp_get_max_threads; {cores} host cores), single MPI rank (shim); one CPU process whatever --gpus is")
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
          "same_config": False, "host_cores": cores}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (shipped closed-form initial condition, mixmode)",
           "config": {"workload": WORKLOAD, "rhs_per_step": 4, "sample_points": npts},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0)  # the reference's static destructors are not exit-safe


def cpu_baseline(steps, warmup, workload="hw3d"):
    """Runs the reference arm in a child process (it owns global state) and returns its cpu_baseline."""
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps),
                            "--warmup", str(warmup), "--workload", workload], capture_output=True, text=True, timeout=600,
                           env={k: v for k, v in os.environ.items()
                                if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS", "OMP_PROC_BIND")})
        for line in reversed(p.stdout.strip().splitlines()):
            if line.startswith("{"):
                d = json.loads(line)
                return d.get("cpu_baseline", {"unavailable": d.get("unavailable")})
        return {"unavailable": (p.stderr or "no output")[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50, help="RK4 steps timed (4 RHS each: the default is 200 RHS)")
    ap.add_argument("--warmup", type=int, default=5, help="untimed RK4 steps first (default: 20 RHS)")
    ap.add_argument("--dump-rhs", default=None, help="reference arm: write the slab state and its RHS to this .npz "
                                                     "(the parity gate of our arm reads it)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the full-size parity gate (profiling runs)")
    ap.add_argument("--no-other", action="store_true", help="skip the short extra measurements of the default line")
    ap.add_argument("--workload", default="hw3d", choices=["hw3d", "elmpb", "blob2d", "hw"],
                    help="blob2d: configs[2], X-decomposed over --gpus (strong scaling); hw: reference arm only (the CPU number "
                         "beside the host-layer extras)")
    ap.add_argument("--decomp", default=os.environ.get("BOUT_B200_DECOMP", "y"), choices=["x", "y", "xy"],
                    help="BoutMesh decomposition at N > 1: NYPE=N (default), NXPE=N, or NXPE=2 x NYPE=N/2")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "elmpb":
        run_elmpb(args)
    elif args.workload == "blob2d":
        run_blob2d(args)
    elif args.workload != "hw3d":
        raise SystemExit("--workload hw exists for --impl reference only; our arm reports it in the default "
                         "line's other.host_layer_models")
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
