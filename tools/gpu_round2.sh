#!/bin/bash
# One GPU-box visit of round 2: parity tests, both bench arms, ncu launch list, one --set full capture of the dominant
# kernel (each ncu pass only after the plain command exited 0), DRAM traffic at the bench's size, disturbed batches,
# horizon sweep, batch sweep, closed loop, replay.  Usage: gpurun -- 'bash tools/gpu_round2.sh <tag>'
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>$O/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python bench.py > $O/${TAG}_bench.log 2>$O/${TAG}_bench.err; rc=$?; echo "bench rc=$rc"; head -c 400 $O/${TAG}_bench.log; echo
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
fi
python tools/prof_case_prepass.py 65536 0.0 40 4 > $O/${TAG}_prof_plain.log 2>&1; rc=$?; echo "prof_case rc=$rc"
if [ $rc -eq 0 ]; then
  cp convex-mpc-unitree-go2_b200/libcmpc.so $O/${TAG}_prof.so
  ncu --set full --clock-control none --import-source on -k "regex:wrench_pdas_kernel" -s 2 -c 1 -f -o $O/${TAG}_prof \
      python tools/prof_case_prepass.py 65536 0.0 40 4 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k "regex:wrench_pdas_kernel|wrench_certificate_kernel|solve_fast_kernel" -s 6 -c 3 --csv --log-file $O/${TAG}_traffic_65536.csv \
      python tools/prof_case_prepass.py 65536 0.0 40 4 > $O/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
fi
for s in 0.05 0.3; do
  python bench.py --stress $s --steps 100 --no-cpu-baseline --no-extras > $O/${TAG}_bench_stress$s.log 2>$O/${TAG}_bench_stress$s.err; echo "stress $s rc=$?"
done
python tests/horizon_sweep.py 16384 > $O/${TAG}_horizon.log 2>&1; cp $O/horizon_sweep.json $O/${TAG}_horizon_sweep.json; echo "horizon rc=$?"
python bench.py --sweep > $O/${TAG}_sweep.log 2>&1; cp $O/sweep_active_set.json $O/${TAG}_sweep_active_set.json; echo "sweep rc=$?"
python tools/closed_loop.py 1024 500 $O/${TAG}_closed_loop_1024x500.json > /dev/null 2>&1; echo "closed loop rc=$?"
CMPC_PREPASS_MIN_BATCH=1 python tools/closed_loop.py 4096 200 $O/${TAG}_closed_loop_4096x200_shift.json shift > /dev/null 2>&1; echo "closed loop shift rc=$?"
CMPC_PREPASS_MIN_BATCH=1 python tools/closed_loop.py 4096 200 $O/${TAG}_closed_loop_4096x200.json > /dev/null 2>&1; echo "closed loop noshift rc=$?"
python tools/closed_loop.py 1 500 $O/${TAG}_closed_loop_single_robot.json > /dev/null 2>&1; echo "closed loop single rc=$?"
timeout 300 python tools/replay.py record 1024 200 $O/replay_1024x200.npz > $O/${TAG}_replay_record.log 2>&1; echo "replay record rc=$?"
timeout 120 python tools/replay.py replay $O/replay_1024x200.npz $O/${TAG}_replay_1024x200.json > $O/${TAG}_replay.log 2>&1; echo "replay rc=$?"
rm -f $O/replay_1024x200.npz
