#!/bin/bash
# Round-end measurement suite (one gpurun call):  bash tools/final_measure.sh <tag>
# -> gpurun_out/<tag>_{pytest.txt,bench.json,profile.txt,launches.csv,update.ncu-rep}
tag=${1:-r2final}
out=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest.txt 2>&1; tail -3 $out/${tag}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
timeout 200 python tools/profile_update.py 32768 3 u8s2d > $out/${tag}_profile.txt 2>&1; head -22 $out/${tag}_profile.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches.csv \
    python tools/profile_update.py 32768 1 u8s2d > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/${tag}_update \
    python tools/profile_update.py 32768 1 u8s2d > $out/${tag}_ncu.log 2>&1; tail -2 $out/${tag}_ncu.log
ls -la $out/${tag}_*
