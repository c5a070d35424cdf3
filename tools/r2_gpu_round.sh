#!/bin/bash
# One GPU session: parity suite, bench line, launch list, ncu --set full of the two tensor-core kernels, DRAM bytes of a full attack.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/rc.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/rc.log
# the persistent attack kernel: 20 iterations are enough for the counters (one launch = the whole attack) ...
ncu --set full --clock-control none --import-source on -k k_tc -s 1 -c 1 -o gpurun_out/k_tc_run_r2 -f \
    python tools/run_dbg.py lotka_volterra_l2pgd_rkl 10000 > gpurun_out/ncu_tc.log 2>&1; echo "ncu k_tc rc=$?" >> gpurun_out/rc.log
# ... and the DRAM bytes of a full 200-iteration launch
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k k_tc -s 1 -c 1 \
    python tools/run_dbg.py lotka_volterra_l2pgd_rkl 10000 200 > gpurun_out/ncu_tc_dram200.log 2>&1; echo "ncu dram200 rc=$?" >> gpurun_out/rc.log
ncu --set full --clock-control none --import-source on -k k_tcd -s 4 -c 1 -o gpurun_out/k_tcd_r2 -f \
    python tools/tc_prof.py pyloric_ctx_l2pgd_rkl 12500 6 > gpurun_out/ncu_tcd.log 2>&1; echo "ncu k_tcd rc=$?" >> gpurun_out/rc.log
python tools/trades_dbg.py nll,fim,trades > gpurun_out/train_steps.log 2>&1; echo "train steps rc=$?" >> gpurun_out/rc.log
cat gpurun_out/rc.log; tail -3 gpurun_out/gputest.log; grep "ms$" gpurun_out/train_steps.log; grep -E "dram__|gpu__time" gpurun_out/ncu_tc_dram200.log
