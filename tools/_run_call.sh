python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s13_tests.log
B="python bench.py --frames 49 --height 576 --width 1024 --dtype bf16 --steps 3 --warmup 3 --no-e2e --no-parity --no-library-baseline --no-scene --no-cpu-baseline"
$B > gpurun_out/s13_cfg5_nofold.json 2> gpurun_out/s13_cfg5.err
UG_LN_FOLD=1 $B > gpurun_out/s13_cfg5_fold.json 2>> gpurun_out/s13_cfg5.err
