# ncu --set full capture of the warp-per-frame n_fft 1920 kernel (one launch of the S3Gen bench workload) + the parity tests first
python -m pytest tests/test_gpu_wpf1920.py -m gpu -x -q 2>&1 | tail -5
python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 5 --warmup 3 > gpurun_out/wpf_bench.json 2>gpurun_out/wpf_bench.err || exit 1
tail -c 600 gpurun_out/wpf_bench.json
ncu --set full --clock-control none --import-source on -k regex:wpf1920 -s 3 -c 1 -o gpurun_out/prof_wpf -f python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 2 --warmup 3 > gpurun_out/prof_wpf.log 2>&1
ls -la gpurun_out/prof_wpf.ncu-rep
