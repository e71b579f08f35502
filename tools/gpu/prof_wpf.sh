# warp-per-frame n_fft 1920 kernel: all GPU tests, the S3Gen bench line (both kernels), then one ncu --set full capture of the kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B2A_WPF1920=1 python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 > gpurun_out/wpf_bench.json 2>gpurun_out/wpf_bench.err || exit 1
B2A_WPF1920=0 python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 > gpurun_out/tiled_bench.json 2>gpurun_out/tiled_bench.err || exit 1
python -c "
import json
for f in ('wpf','tiled'):
    d=json.loads(open('gpurun_out/%s_bench.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
ncu --set full --clock-control none --import-source on -k regex:wpf1920 -s 3 -c 1 -o gpurun_out/prof_wpf3 -f python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 2 --warmup 3 > gpurun_out/prof_wpf3.log 2>&1
ls -la gpurun_out/prof_wpf3.ncu-rep
