for gfill in 0 1; do
BOUT_B200_GUARD_FILL=$gfill ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_hw3d -c 12 --csv --log-file gpurun_out/r02w_fused_$gfill.csv python bench.py --no-other --no-cpu-baseline --no-parity-gate --steps 2 --warmup 1 > /dev/null 2>&1
done
