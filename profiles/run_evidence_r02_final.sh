#!/bin/bash
# Final evidence of round 2 on one B200 (under gpurun):  bash profiles/run_evidence_r02_final.sh
# GPU test suite, smoke, bench lines (default = configs[1] + configs[4] on one GPU, reference arm, configs[2], configs[3]),
# the angle-scan variants side by side, and the ncu launch list + full capture of the benchmark step (tag r02_v4).
set -u
OUT=gpurun_out/ev_r02_final
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; tail -2 $OUT/gpu_tests.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python bench.py > $OUT/bench_default.jsonl 2> $OUT/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.jsonl 2> $OUT/bench_reference.err
python bench.py --samples 512 --chirps 256 --antennas 192 --method esprit --frames 8 --chunk 8 --e2e-frames 8 --host-chunk 4 \
    --cpu-frames 0 --no-configs4 --sustain-s 0 --steps 5 > $OUT/bench_configs2.jsonl 2> $OUT/bench_configs2.err
python bench.py --threshold-db 31 --irls 3 --cpu-frames 2 --no-configs4 --sustain-s 0 > $OUT/bench_configs3.jsonl 2> $OUT/bench_configs3.err
python bench.py --antennas 16 --frames 512 --chunk 512 --workload configs1 --no-configs4 --cpu-frames 1 --sustain-s 0 \
    --e2e-frames 64 > $OUT/bench_a16_stages.jsonl 2> $OUT/bench_a16_stages.err
python profiles/angles_bench.py --antennas 8 --frames 1000 > $OUT/angles_a8.jsonl 2>&1
python profiles/angles_bench.py --antennas 16 --frames 512 > $OUT/angles_a16.jsonl 2>&1
RS_ANGLES_DEDUP=0 python profiles/angles_bench.py --antennas 16 --frames 512 > $OUT/angles_a16_nodedup.jsonl 2>&1
bash profiles/run_ncu.sh r02_v4 > $OUT/run_ncu.log 2>&1
ls -la $OUT gpurun_out/ncu_r02_v4
