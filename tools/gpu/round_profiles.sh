# Round evidence: smoke, default bench line, reference arm, every workload, launch list and ncu --set full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in istft_hift istft_kokoro hift_head whisper_segment funasr kaldi s3gen whisper80_1clip; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
# launch list of the default command (per-launch times are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"b2a|frontend_kernel|whisper_clamp" -c 40 --csv --log-file gpurun_out/launches_whisper128.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
TAG=${TAG:-r5} bash tools/gpu/prof_whisper.sh
ncu --set full --clock-control none -k regex:frontend_kernel -c 1 -o gpurun_out/prof_frontend_full_${TAG:-r5} -f python bench.py --workload whisper128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
for w in istft_hift istft_kokoro hift_head; do
  ncu --set full --clock-control none --import-source on -k regex:istft_kernel -c 1 -o gpurun_out/prof_${w}_${TAG:-r5} -f python bench.py --workload $w --batch 128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
done
ncu --set full --clock-control none -k regex:istft_kernel -c 1 -o gpurun_out/prof_istft_hift_full_${TAG:-r5} -f python bench.py --workload istft_hift --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
tail -n 1 gpurun_out/bench_default.json | cut -c1-300
