#!/bin/bash
# ncu evidence for one round: launch list (per-launch device time) + one full capture of every kernel.
# usage (under gpurun):  bash profiles/run_ncu.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out/ncu_$TAG
mkdir -p $OUT
CMD="python bench.py --frames 200 --chunk 100 --steps 1 --warmup 3 --e2e-frames 50 --cpu-frames 1 --no-configs4 --no-side-configs --sustain-s 0 --profile-passes 1"
KREGEX='regex:fft2d_|range_fft|doppler_fft|detect_|compact_masks|angles_|velocity_|recheck_|cube_stream|scatterer_plane'
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 50 -c 64 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KREGEX" -s 50 -c 16 \
    -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
ls -la $OUT
