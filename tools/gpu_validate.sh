#!/bin/bash
# One GPU-box pass: kernel check -> GPU test-suite -> (if green) ncu DRAM-traffic capture of one step + the bench line.
# Usage (from the repo root, under gpurun): bash tools/gpu_validate.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 150 python tools/encoder_check.py 1 7 130 1000 40000 > $out/${tag}_encoder_check.txt 2>&1
rc_check=$?
echo "encoder_check rc=$rc_check"; tail -30 $out/${tag}_encoder_check.txt
if [ $rc_check -eq 124 ]; then echo "encoder_check hung: stopping"; exit 1; fi
timeout 400 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
rc_test=$?
echo "pytest rc=$rc_test"; tail -25 $out/${tag}_pytest_gpu.log
if [ $rc_test -ne 0 ]; then
    timeout 200 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err
    echo "quick bench rc=$?"; tail -5 $out/${tag}_bench_quick.err; cut -c1-300 $out/${tag}_bench_quick.json
    exit 2
fi
timeout 200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file $out/${tag}_ncu_traffic_c2.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e --cuda-profiler \
    > $out/${tag}_ncu_traffic.json 2> $out/${tag}_ncu_traffic.err
echo "ncu rc=$?"; wc -l $out/${tag}_ncu_traffic_c2.csv
python tools/traffic_from_ncu.py $out/${tag}_ncu_traffic_c2.csv c2_lstm > $out/${tag}_traffic_summary.txt 2>&1 && cp profiles/kernel_traffic.json $out/${tag}_kernel_traffic.json
cat $out/${tag}_traffic_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_c2_n1.json 2> $out/${tag}_bench_c2_n1.err
echo "bench rc=$?"; tail -12 $out/${tag}_bench_c2_n1.err; cut -c1-400 $out/${tag}_bench_c2_n1.json
