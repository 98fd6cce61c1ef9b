set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>&1; tail -c 400 gpurun_out/bench_ref_final.json
(for n in 4096 16384 32768 65536 131072 262144; do for m in warp tpb pair; do XQ_PLAYOUT_MODE=$m python scripts/playout_rate.py $n 4; done; done) > gpurun_out/mappings.txt 2>&1
python bench.py --fast --no-cpu --steps 3 > gpurun_out/fast_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pair.csv python bench.py --fast --no-cpu --steps 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
