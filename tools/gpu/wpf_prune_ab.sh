# pruned stage B of the warp-per-frame n_fft 1920 kernel: parity tests, then the S3Gen bench workload with (B2A_WPF1920=1) and without (=2) pruning
python -m pytest tests/test_gpu_wpf1920.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_ragged.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -15
q() { B2A_WPF1920=$1 python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 2>gpurun_out/wpfp_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wpf=$1', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
q 1; q 2; q 1; q 2
