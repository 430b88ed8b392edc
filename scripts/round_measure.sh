set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/r01e_bench.log 2>gpurun_out/r01e_bench.err; tail -1 gpurun_out/r01e_bench.log > gpurun_out/r01e_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 1 --no-baselines > gpurun_out/r01e_ncu_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_warp_gpu.py tests/test_nv12_gpu.py tests/test_overlap_gpu.py -q -x 2>&1 | tail -8 > gpurun_out/r01e_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_warp_gpu.py "tests/test_overlap_gpu.py::test_frame_loop_over_several_host_threads" -q -x 2>&1 | tail -8 > gpurun_out/r01e_racecheck.txt
