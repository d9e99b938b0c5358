set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; cat gpurun_out/s3_bench.json | head -c 3000
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wave$ -c 1 -o gpurun_out/kwave3_c5 python tools/run_c5.py 20000 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cluster_merge -c 1 -o gpurun_out/kcluster python tools/run_tree.py 2000 400 0 > gpurun_out/kcluster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_rows -s 1 -c 1 -o gpurun_out/kbuildrows python tools/run_profile_batch.py 120 400 200 > gpurun_out/kbuildrows.log 2>&1
python tools/run_profile_batch.py 300 400 200
ls -la gpurun_out
