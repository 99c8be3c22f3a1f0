#!/bin/bash
# Round-2 evidence run (one GPU): tests, smoke, bench lines, kernel timings, learning check, ncu launch lists + full captures.
# Everything lands in gpurun_out/ (copied into profiles/ afterwards).  ncu passes only after the same command exited 0 without ncu.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.txt 2>&1; tail -3 $O/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; tail -2 $O/r2_smoke.txt
python bench.py > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; tail -c 400 $O/r2_bench_c3.json
python bench.py --impl reference > $O/r2_bench_ref.json 2>> $O/r2_bench_c3.err
python bench.py --workload c2 --no-ppo > $O/r2_bench_c2.json 2>> $O/r2_bench_c3.err
python bench.py --workload c5 --no-ppo > $O/r2_bench_c5.json 2>> $O/r2_bench_c3.err
python bench.py --workload c4 > $O/r2_bench_c4_ppo.json 2>> $O/r2_bench_c3.err
python profiles/measure_aux_kernels.py > $O/r2_aux_kernels.jsonl 2>> $O/r2_bench_c3.err
python profiles/measure_policy.py > $O/r2_policy_forward.jsonl 2>> $O/r2_bench_c3.err
python profiles/measure_update.py 131072 fused --timeline > $O/r2_ppo_update_step.txt 2>&1
python profiles/learning_check.py > $O/r2_learning_check.txt 2>&1; tail -3 $O/r2_learning_check.txt
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ppo"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 400 --csv --log-file $O/r2_launches_bench_c3.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 310 -c 3 -f -o $O/r2_step_kernel_c3 $CMD > $O/ncu2.log 2>&1
POL="python profiles/measure_policy.py 16384"
$POL > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 2 -f -o $O/r2_fused_block $POL > $O/ncu3.log 2>&1
$POL > $O/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file $O/r2_launches_policy_forward_B16384.csv $POL > $O/ncu3b.log 2>&1
UPD="python profiles/measure_update.py 131072 fused"
$UPD > $O/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --launch-skip 500 -c 220 --csv --log-file $O/r2_launches_ppo_update_step.csv $UPD > $O/ncu4.log 2>&1
ls -la $O | tail -5
