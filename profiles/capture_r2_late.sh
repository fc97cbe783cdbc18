#!/bin/bash
# Late round-2 re-capture after the axis-sum / SpMV kernel changes: launch list of the bench step, the config-2 axis
# sums, the config-4 row sums and SpMV.  Reports are turned into their text summaries on the box (the .ncu-rep
# files together exceed what gpurun copies back).  Usage: bash profiles/capture_r2_late.sh OUTDIR
out=${1:-gpurun_out}
set -x
python bench.py --steps 3 --warmup 3 --no-ops > $out/r2_bench_plain.json 2> $out/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size \
    --clock-control none -c 400 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-ops > $out/ncu_bench.log 2>&1
python profiles/run_kernels.py sum_last sum_mid > $out/plain_kernels.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'slab_|axis_accumulate|touched' -c 8 \
    -f -o /tmp/r2_axis_sums python profiles/run_kernels.py sum_last sum_mid > $out/ncu_axis.log 2>&1
ncu -i /tmp/r2_axis_sums.ncu-rep --page raw --csv > $out/r2_axis_sums_ncu_raw.csv
python profiles/summarize_rep.py /tmp/r2_axis_sums.ncu-rep > $out/r2_axis_sums_ncu_summary.txt
python profiles/run_kernels.py spmv sum_c4 > $out/plain_kernels_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'spmv_|axis_accumulate|touched' -c 5 -f \
    -o /tmp/r2_c4 python profiles/run_kernels.py spmv sum_c4 > $out/ncu_c4.log 2>&1
ncu -i /tmp/r2_c4.ncu-rep --page raw --csv > $out/r2_c4_ncu_raw.csv
python profiles/summarize_rep.py /tmp/r2_c4.ncu-rep > $out/r2_c4_ncu_summary.txt
echo done
