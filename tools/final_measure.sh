#!/bin/bash
# One gpurun call on one B200: the measurements the round's documents quote (written under gpurun_out/).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r02_headline_parity.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > $O/r02_tests_final.log 2>&1; tail -3 $O/r02_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
{
echo "# bench.py, final tree of round 2, one B200 (N=1)"
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1
echo "# bench.py --impl reference --steps 5 --warmup 2"
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1
echo "# bench.py --config eval --batch 64 / 512"
timeout 600 python bench.py --config eval --batch 64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
timeout 600 python bench.py --config eval --batch 512 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
echo "# bench.py --config dense"
timeout 600 python bench.py --config dense --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
echo "# bench.py --no-graph (eager launches: the path ragged batches take)"
timeout 600 python bench.py --no-graph --steps 20 --warmup 5 --no-fp32 --no-cpu-baseline 2>/dev/null | tail -1
} > $O/r02_bench_final.log
cut -c1-200 $O/r02_bench_final.log
timeout 300 python tools/step_parts.py > $O/r02_step_parts.log 2>&1; tail -5 $O/r02_step_parts.log
timeout 300 python tools/bench_augment.py > $O/r02_bench_augment.json 2>$O/bench_aug.err; tail -c 400 $O/r02_bench_augment.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_train_step.csv python tools/profile_step.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_eval64.csv python tools/profile_step.py --eval --batch 64 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file $O/r02_metrics_train_step.csv python tools/profile_step.py > /dev/null 2>&1
ls -la $O | tail -15
