"""Cost of one Mesh::communicate of nf fields at the HW3D bench slab (516x132x512), Y-decomposed over the ranks:
torchrun --nproc-per-node N tools/time_halo.py.  Compares BOUT_B200_HALO=peer (NVLink peer memory) and nccl."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg = ge.load_package()
for mode in ("peer", "nccl"):
    os.environ["BOUT_B200_HALO"] = mode
    mesh = pkg.mesh.BoutMesh(1, world, rank, 512, 128, 512, 2, 2, periodic_x=False, periodic_y=True, device=dev)
    fields = [torch.rand(mesh.grid.shape3, dtype=torch.float64, device=dev) for _ in range(3)]
    for nf in (1, 3):
        fs = fields[:nf]
        for _ in range(5):
            mesh.communicate(fs)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            mesh.communicate(fs)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        mb = nf * 516 * 2 * 512 * 8 / 1e6
        if rank == 0:
            print(f"{mode:5s} nf={nf}: {us:7.1f} us per communicate ({mb:.1f} MB each way, {2 * mb / us * 1e-3:.1f} GB/s in+out per GPU)"
                  f"  timed_out={mesh.halo_timed_out()}", flush=True)
dist.barrier()
dist.destroy_process_group()
