#!/bin/bash
# Round-end measurement session (run under gpurun from the repo root): parity suite, smoke, bench (both arms),
# ncu launch list and DRAM traffic of the dominant kernel.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 60 python scripts/gpu_dbg.py 2 || exit 1
timeout 700 python -m pytest tests -m gpu -x -q --timeout 200 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "bench rc=$?"; tail -c 300 $O/bench_default.err; tail -3 $O/bench_default.time
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 0 --sample-seconds 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:k_opt_blocks -c 1 --csv --log-file $O/traffic.csv python bench.py --steps 1 --warmup 0 --sample-seconds 1 > $O/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
