#!/bin/bash
# first GPU trip: environment probe + kernel-1a parity tests
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
(nproc; free -g; python -c "import numba,os;print('numba',numba.__version__,'threads',numba.get_num_threads(),'cpus',os.cpu_count())"; lscpu | head -20) > gpurun_out/host.txt 2>&1
timeout 900 python -m pytest tests/test_window_stats_gpu.py -x -q -m gpu > gpurun_out/pytest_stats.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_stats.log
tail -30 gpurun_out/pytest_stats.log
