#!/bin/bash
# One GPU-box session, in bounded stages (every stage has its own timeout; nothing here may run longer than a
# few minutes - a 25-minute overrun cost this round half its GPU budget). Everything lands in gpurun_out/$TAG
# (scratch); scripts/ncu_launches.py / ncu_summary.py / ncu_hot.py / traffic_json.py distil it into profiles/.
#   gpurun --timeout 600 -- 'bash scripts/gpu_session.sh r02 "tests bench"'
#   gpurun --timeout 600 -- 'bash scripts/gpu_session.sh r02 "steps full"'
TAG=${1:-r02}
STAGES=${2:-"tests bench"}
O=gpurun_out/$TAG
mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off --target-processes application-only"
for s in $STAGES; do case $s in
tests)
  timeout -s KILL 900 python -m pytest tests -m gpu -q -s --durations=8 2>&1 | grep -v "^Training\|it/s\|s/it" | tail -60 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log ;;
bench)
  timeout -s KILL 300 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err ;;
steps)
  # one steady-state step of each workload: device time + DRAM bytes of every launch
  for k in rec det; do
    timeout -s KILL 120 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
        --log-file $O/step_$k.csv python scripts/one_step.py $k > /dev/null 2>&1
  done ;;
launches)
  # the launch list of the bench command itself (first 1500 launches; --no-clocks: no nvidia-smi child under ncu)
  timeout -s KILL 150 ncu --clock-control none --target-processes application-only --metrics gpu__time_duration.sum -c 1500 \
      --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-clocks \
      > $O/bench_under_ncu.log 2>&1 ;;
full)
  # --set full replays every kernel ~40 times: a handful of launches per capture, never a whole step
  timeout -s KILL 120 $NCU --set full --import-source on -k regex:'gemm_tc' -c 6 -o $O/rec_gemm_full -f python scripts/one_step.py rec > /dev/null 2>&1
  timeout -s KILL 120 $NCU --set full --import-source on -k regex:'gru_' -c 4 -o $O/rec_gru_full -f python scripts/one_step.py rec > /dev/null 2>&1
  timeout -s KILL 90 $NCU --set full --import-source on -k regex:'ctc_alpha|ctc_beta' -o $O/ctc_full -f python scripts/one_step.py ctc 8192 > /dev/null 2>&1
  timeout -s KILL 150 $NCU --set full --import-source on -k regex:'sep_fwd|pw_wgrad_saved|sep_dw_bwd|convt_wgrad_staged' -s 6 -c 8 -o $O/det_full -f python scripts/one_step.py det 8 > /dev/null 2>&1 ;;
esac; done
du -sh $O; ls -la $O
