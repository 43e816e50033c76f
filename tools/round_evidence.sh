#!/bin/bash
# One GPU call's worth of end-of-round evidence: bench line, ncu captures of the dominant kernels (each after the
# same command has exited 0 without ncu), the launch list of the bench command and the tensor-core FIR A/B.
# usage (on the GPU box, from the repo root): bash tools/round_evidence.sh <tag>
tag=${1:-s}
o=gpurun_out
mkdir -p $o
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 300 $o/${tag}_bench.err
: > $o/${tag}_p.log
cap() { # what kernel-regex stem
    python tools/prof_one.py $1 3 >> $o/${tag}_p.log 2>&1 &&
        ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o $o/${tag}_$3 python tools/prof_one.py $1 3 >> $o/${tag}_p.log 2>&1
}
cap fftmag fft4096_tma fftmag
cap tc64 fir_tc_ts tc64
cap tc256 fir_tc_ts tc256
cap pfb pfb64_kernel pfb64
grep "ms/launch" $o/${tag}_p.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
tail -2 $o/${tag}_launches.csv | cut -c1-200
timeout 300 python tools/tc_check.py time > $o/${tag}_tc_time.log 2>&1
