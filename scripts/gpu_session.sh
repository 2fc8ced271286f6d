#!/bin/bash
# One GPU-box session: parity tests, bench line, ncu launch lists and --set full captures.
# Everything lands in gpurun_out/ (scratch); scripts/ncu_to_profiles.py distils it into profiles/.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh r01 [stages]'
TAG=${1:-r01}
STAGES=${2:-"tests bench launches full"}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $O/gpu.txt
NCU="ncu --clock-control none --profile-from-start off"
for s in $STAGES; do case $s in
tests)
  timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 > $O/pytest_gpu.log ;;
bench)
  timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
  tail -3 $O/bench.err ;;
launches)
  # the launch list of the bench command itself (cold-cache, serialised per-launch times)
  timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
  # one steady-state step of each workload
  for k in rec det; do
    timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
        --log-file $O/step_$k.csv python scripts/one_step.py $k > /dev/null 2>&1
  done ;;
full)
  timeout 600 $NCU --set full --import-source on -o $O/ctc_full -f python scripts/one_step.py ctc 8192 > /dev/null 2>&1
  timeout 900 $NCU --set full --import-source on -o $O/rec_full -f python scripts/one_step.py rec > /dev/null 2>&1
  timeout 900 $NCU --set full --import-source on -k regex:'dwpw_fwd|dw_bwd|pwT_bwd|pw_wgrad|bnrelu_bwd|convt_fwd|convt_bwd|convt_wgrad|pool2|outconv|bce_' \
      -o $O/det_full -f python scripts/one_step.py det > /dev/null 2>&1
  for r in ctc_full rec_full det_full; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    ls -la $O/$r.ncu-rep
  done
  # keep the merge-back under 64 MiB: big reports stay on the box, their raw CSV comes home
  for r in rec_full det_full; do
    if [ -f $O/$r.ncu-rep ] && [ $(stat -c %s $O/$r.ncu-rep) -gt 20000000 ]; then
      ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null; rm -f $O/$r.ncu-rep; fi
  done ;;
esac; done
du -sh $O; ls $O
cat $O/pytest_gpu.log 2>/dev/null | tail -15
cat $O/bench.json 2>/dev/null
