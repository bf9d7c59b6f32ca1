# usage: bash tools/profile_round.sh <tag>   (one B200; writes gpurun_out/<tag>_*)
tag=$1
python bench.py > gpurun_out/${tag}_bench_n1_hw3d.json 2> gpurun_out/${tag}_bench_n1_hw3d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_hw3d.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other --no-parity-gate > /dev/null 2> gpurun_out/${tag}_ncu.log
ncu --set full --clock-control none --import-source on -k regex:'k_hw3d_rhs_tma|k_lapb_rows|k_lapb_reduced' -s 12 -c 8 \
    -o gpurun_out/${tag}_prof_hw3d python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other --no-parity-gate > /dev/null 2>> gpurun_out/${tag}_ncu.log
ncu -i gpurun_out/${tag}_prof_hw3d.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_raw_hw3d_kernels.csv 2>/dev/null
ncu -i gpurun_out/${tag}_prof_hw3d.ncu-rep --page details --csv > gpurun_out/${tag}_ncu_full_details_hw3d_kernels.csv 2>/dev/null
