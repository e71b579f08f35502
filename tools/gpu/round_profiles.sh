# Round evidence: default bench line, reference arm, launch list and ncu --set full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in istft_hift istft_kokoro funasr kaldi s3gen whisper80_1clip; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
# launch list of the default command (per-launch times are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_whisper128.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
TAG=${TAG:-r3} bash tools/gpu/prof_whisper.sh
ncu --set full --clock-control none --import-source on -k regex:istft_kernel -c 1 -o gpurun_out/prof_istft_hift_${TAG:-r3} -f python bench.py --workload istft_hift --batch 128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_istft_hift.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:istft_kernel -c 1 -o gpurun_out/prof_istft_kokoro_${TAG:-r3} -f python bench.py --workload istft_kokoro --batch 128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_istft_kokoro.log 2>&1
tail -2 gpurun_out/bench_default.json | cut -c1-1500
