#!/bin/bash
# Round-2 evidence captures (each program has already exited 0 without ncu).  Usage: bash profiles/capture_r2.sh OUTDIR
out=${1:-gpurun_out}
set -x
python bench.py --steps 3 --warmup 3 --no-ops > $out/r2_bench_plain.json 2> $out/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size \
    --clock-control none -c 400 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-ops > $out/ncu_bench.log 2>&1
python profiles/run_kernels.py intersect union sum_last sum_mid transpose > $out/plain_kernels.log 2>&1 || exit 1
# the three launches of one fused intersection call, the union kernel, the slab axis-sum launches, the radix launches
ncu --set full --clock-control none --import-source on -k regex:'merge_partition|setop_kernel|merge_reorder' -c 5 \
    -f -o $out/r2_setops python profiles/run_kernels.py intersect union > $out/ncu_setops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'slab_|axis_accumulate|touched' -c 8 \
    -f -o $out/r2_axis_sums python profiles/run_kernels.py sum_last sum_mid > $out/ncu_axis.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'radix_' -c 5 \
    -f -o $out/r2_radix python profiles/run_kernels.py transpose > $out/ncu_radix.log 2>&1
# config-4 operand (1e6 x 1e6 float64, 100 M stored elements): SpMV and the row sums
python profiles/run_kernels.py spmv sum_c4 > $out/plain_kernels_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'spmv_|axis_accumulate|touched' -c 5 -f \
    -o $out/r2_c4 python profiles/run_kernels.py spmv sum_c4 > $out/ncu_c4.log 2>&1
echo done
