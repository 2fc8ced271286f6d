#!/bin/bash
# quick GPU check of the detection path: unit tests, bench-shape parity, det bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_det_gpu.py -x -q 2>&1 | tail -4
timeout 200 python -m pytest tests/test_bench_shapes_gpu.py -x -q -s -k "det" 2>&1 | grep -E "rel err|passed|failed|Error"
timeout 200 python bench.py --workload det --no-secondary --no-cpu-baseline --steps 10 > gpurun_out/det_b1.json 2>gpurun_out/det_b1.err
python -c "
import json; d=json.load(open('gpurun_out/det_b1.json')); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['first_step_loss']['rel_err'])"; tail -3 gpurun_out/det_b1.err
