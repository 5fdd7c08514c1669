#!/bin/bash
# One gpurun call (round 2, fourth pass): the GPU test suite on the final tree, the helper-kernel event timings and one light ncu
# launch of the source-centric up-sampling kernel.  Only compact text leaves the box.
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q > $O/r02d_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/r02d_gputests.log
timeout 60 python tools/hbm_kernels_bench.py $O/r02d_hbm_kernels.json > $O/r02d_hbm_kernels.txt 2>&1; echo "hbm bench rc=$?"; tail -n 3 $O/r02d_hbm_kernels.txt
LIGHT="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size"
timeout 90 ncu $LIGHT --clock-control none --kernel-name-base demangled -k "regex:upsample_d" -c 1 -o $O/tmp_r02d -f python tools/misc_ops_once.py > $O/ncu_r02d.log 2>&1
echo "ncu rc=$?"
ncu -i $O/tmp_r02d.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py upsample_r02d > $O/r02d_ncu_upsample.txt
rm -f $O/tmp_r02d.ncu-rep
cat $O/r02d_ncu_upsample.txt
