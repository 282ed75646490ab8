#!/bin/bash
# Round-2 evidence on one B200 (under gpurun):  bash profiles/run_evidence_r02.sh
# bench lines of every BASELINE config that fits one GPU, the reference arm, the ncu launch list + full capture of the
# benchmark step, and an ncu capture of the general-covariance eigensolver (music_cov_kernel).
set -u
OUT=gpurun_out/ev_r02
mkdir -p $OUT
python bench.py > $OUT/bench_default.jsonl 2> $OUT/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.jsonl 2> $OUT/bench_reference.err
# configs[2]: MIMO 12Tx x 16Rx = 192 virtual channels, 512 x 256 cube, ESPRIT
python bench.py --samples 512 --chirps 256 --antennas 192 --method esprit --frames 8 --chunk 8 --e2e-frames 8 --host-chunk 4 \
    --cpu-frames 0 --no-configs4 --sustain-s 0 --steps 5 > $OUT/bench_configs2.jsonl 2> $OUT/bench_configs2.err
# configs[3] style: dense scene through the robust (Huber IRLS) solve; the batched eigensolver itself is timed below
python bench.py --threshold-db 31 --irls 3 --cpu-frames 2 --no-configs4 --sustain-s 0 > $OUT/bench_configs3.jsonl 2> $OUT/bench_configs3.err
python profiles/time_jacobi.py > $OUT/time_jacobi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:music_cov -c 2 -o $OUT/music_cov python profiles/time_jacobi.py > $OUT/ncu_music_cov.log 2>&1
bash profiles/run_ncu.sh r02_v2 > $OUT/run_ncu.log 2>&1
ls -la $OUT gpurun_out/ncu_r02_v2
