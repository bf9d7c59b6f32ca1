for gfill in 0 1; do
BOUT_B200_GUARD_FILL=$gfill python bench.py --no-other --no-cpu-baseline --no-parity-gate --steps 20 > gpurun_out/r02v_bench_$gfill.json 2> gpurun_out/r02v_bench_$gfill.err; tail -3 gpurun_out/r02v_bench_$gfill.err
python -c "
import json; d=json.loads(open('gpurun_out/r02v_bench_$gfill.json').read().strip().splitlines()[-1]); print($gfill, d['ms_per_step'], d['value'], d['rhs_phase_ms'], d['e2e']['result_equals_device_resident_step'], d['gpu_launches'], d['roofline']['kernel_ms'])"
done
