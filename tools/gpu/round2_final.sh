# Round-2 evidence: tests, smoke, the default bench line (with the workloads map), the reference arm, extra workloads, the launch list of the
# default command, an ncu --set full capture of the Whisper kernel.  Outputs under gpurun_out/r2f_*.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
for w in chatterbox128 voice_encoder stft_kokoro stft_hift hift_head whisper_segment whisper128_f16 whisper128_ragged whisper128_padded; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-e2e > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_whisper128.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/r2f_ncu_launches.log 2>&1
TAG=r2final bash tools/gpu/prof_whisper.sh
# the n_fft 1920 front end under both kernels (warp per frame / tiled) and the ncu capture of the warp-per-frame kernel
B2A_WPF1920=0 python bench.py --workload s3gen --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2f_bench_s3gen_tiled.json 2> gpurun_out/r2f_bench_s3gen_tiled.err
python bench.py --workload s3gen --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2f_bench_s3gen_wpf1920.json 2> gpurun_out/r2f_bench_s3gen_wpf1920.err
ncu --set full --clock-control none --import-source on -k regex:wpf1920 -s 3 -c 1 -o gpurun_out/prof_wpf_r2final -f python bench.py --workload s3gen --no-cpu --no-e2e --no-secondary --steps 2 --warmup 3 > gpurun_out/prof_wpf_r2final.log 2>&1
for f in gpurun_out/r2f_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('r2f_bench_')[1][:-5], 'ms', round(d.get('ms_per_step',0),4), 'value %.4g'%d['value'], 'frac', round((d.get('roofline') or {}).get('frac',0) or 0,4), 'e2e %.4g'%(e.get('value') or 0), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'launches', d.get('gpu_launches'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as ex: print(sys.argv[1], 'ERR', ex)
PY
done
