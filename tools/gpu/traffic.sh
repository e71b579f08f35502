# ncu --set full at the bench's full batch sizes, for roofline.traffic (dram bytes per launch)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --workload whisper128 --no-cpu --no-e2e --steps 10 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('whisper128', round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"
ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -c 1 -o gpurun_out/prof_frontend_full_${TAG:-r3} -f python bench.py --workload whisper128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:istft_kernel -c 1 -o gpurun_out/prof_istft_hift_full_${TAG:-r3} -f python bench.py --workload istft_hift --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full.log gpurun_out/ncu_full2.log
