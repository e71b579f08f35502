# ncu --set full capture of the Whisper 128-mel frontend kernel at 256 clips (24000 tiles)
ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -c 1 -o gpurun_out/prof_frontend_${TAG:-x} -f python bench.py --workload whisper128 --batch 256 --no-cpu --no-e2e --no-secondary --steps 1 --warmup 3 > gpurun_out/ncu_${TAG:-x}.log 2>&1
tail -2 gpurun_out/ncu_${TAG:-x}.log
