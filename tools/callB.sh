mkdir -p gpurun_out/r01b
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "heev or svd" > gpurun_out/r01b/pytest_heev.txt 2>&1; tail -5 gpurun_out/r01b/pytest_heev.txt
timeout 200 python tools/bench_heev.py > gpurun_out/r01b/heev_microbench.json 2> gpurun_out/r01b/heev_microbench.err; cat gpurun_out/r01b/heev_microbench.json; tail -3 gpurun_out/r01b/heev_microbench.err
timeout 500 python tools/debug_status.py > gpurun_out/r01b/debug_status.txt 2>&1; tail -12 gpurun_out/r01b/debug_status.txt | cut -c1-600
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r01b/pytest_gpu.txt 2>&1; tail -5 gpurun_out/r01b/pytest_gpu.txt
