# ncu --set full capture of the tensor-core Whisper front end (B2A_WHISPER_TC=1) at 256 clips
export B2A_WHISPER_TC=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_whisper_kernel -c 1 -o gpurun_out/prof_tc_${TAG:-x} -f python bench.py --workload whisper128 --batch 256 --no-cpu --no-e2e --no-secondary --steps 1 --warmup 3 > gpurun_out/ncu_tc_${TAG:-x}.log 2>&1
tail -2 gpurun_out/ncu_tc_${TAG:-x}.log
