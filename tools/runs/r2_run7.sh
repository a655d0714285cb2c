#!/bin/bash
# round 2, ncu call: launch list with DRAM bytes of the bench step, --set full of the fused encoder kernel (report kept)
# and of one launch of each other main kernel (exported to CSV on the box; the reports are too large to bring back)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-modes"
$CMD > gpurun_out/r2j_plain.log 2> gpurun_out/r2j_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r2j_launches.csv $CMD > gpurun_out/r2j_ncu1.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv01_kernel -s 3 -c 1 -o gpurun_out/r2j_conv01_full $CMD > gpurun_out/r2j_ncu2.log 2>&1
echo "ncu full conv01 exit $?"
ncu -i gpurun_out/r2j_conv01_full.ncu-rep --page raw --csv > gpurun_out/r2j_conv01_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'attention_tc|rnn_tc|gemm_lin|ffn_fused|gemm_2sm|probs_kernel' -s 70 -c 14 -o /tmp/r2j_others_full $CMD > gpurun_out/r2j_ncu3.log 2>&1
echo "ncu full others exit $?"
ncu -i /tmp/r2j_others_full.ncu-rep --page raw --csv > gpurun_out/r2j_others_full_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2j
du -sh gpurun_out
