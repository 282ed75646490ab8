#!/bin/bash
# Who gets the 16 SMs the 4-CTA clusters of the fused FFT kernel strand: the side FFT kernel (RS_K12_SIDE permille of the
# frames) or the fp64 recheck of the previous launch set (side stream)?  And how many clusters should the FFT kernel run?
for side in 0 60; do for ncl in 33 32 31 30 28; do
echo "== RS_K12_SIDE=$side RS_K12_CLUSTERS=$ncl"; RS_K12_SIDE=$side RS_K12_CLUSTERS=$ncl python profiles/overlap_probe.py 2>&1 | grep -E "^pipelined  |^serial|^no recheck  " ; done; done
