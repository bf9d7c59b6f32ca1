for ch in 64 57 52 32 29; do
BOUT_B200_HW3D_CHUNK=$ch python bench.py --no-other --no-cpu-baseline --no-parity-gate --steps 20 > gpurun_out/r02x_bench_$ch.json 2> gpurun_out/r02x_bench_$ch.err; tail -3 gpurun_out/r02x_bench_$ch.err
python -c "
import json; d=json.loads(open('gpurun_out/r02x_bench_$ch.json').read().strip().splitlines()[-1]); print($ch, d['ms_per_step'], d['value'], d['rhs_phase_ms'], d['e2e']['result_equals_device_resident_step'], d['gpu_launches'], d['roofline']['kernel_ms'])"
done
