#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list, one --set full capture of the fused kernel
# (each ncu pass only after the plain command exited 0).  Usage: gpurun -- 'bash tools/gpu_round.sh <tag>'
set -u
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>$O/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python bench.py > $O/${TAG}_bench.log 2>$O/${TAG}_bench.err; rc=$?; echo "bench rc=$rc"; tail -c 600 $O/${TAG}_bench.log
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
fi
python tools/prof_case.py 4736 0.0 40 > $O/${TAG}_prof_plain.log 2>&1; rc=$?; echo "prof_case rc=$rc"
if [ $rc -eq 0 ]; then
  cp convex-mpc-unitree-go2_b200/libcmpc.so $O/${TAG}_prof.so
  # the two kernels of one cmpc_solve call (Riccati pre-pass, condensed kernel on the work-list), third solve of the run
  ncu --set full --clock-control none --import-source on -k "regex:riccati2_lockstep_kernel|solve_fast_kernel" -s 4 -c 2 -f -o $O/${TAG}_prof \
      python tools/prof_case.py 4736 0.0 40 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  # DRAM traffic and durations of the same two kernels at the bench's own size (65 536 robots)
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k "regex:riccati2_lockstep_kernel|solve_fast_kernel" -s 4 -c 2 --csv --log-file $O/${TAG}_traffic_65536.csv \
      python tools/prof_case.py 65536 0.0 40 > $O/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
fi
python tools/closed_loop.py 1024 500 $O/${TAG}_closed_loop.json > /dev/null 2>&1; echo "closed loop rc=$?"
python bench.py --sweep > $O/${TAG}_sweep.log 2>&1; cp $O/sweep_active_set.json $O/${TAG}_sweep_active_set.json; echo "sweep rc=$?"
python tests/horizon_sweep.py 16384 > $O/${TAG}_horizon.log 2>&1; cp $O/horizon_sweep.json $O/${TAG}_horizon_sweep.json; echo "horizon rc=$?"
