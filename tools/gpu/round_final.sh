# Round-end evidence after the statistics / LFR / Kokoro changes: tests, smoke, default bench line, reference arm, every workload,
# ncu headline metrics of the changed kernels.  Outputs under gpurun_out/final_*.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
for w in funasr kaldi istft_kokoro; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err; done
for w in istft_hift s3gen hift_head whisper_segment whisper128_ragged whisper80_1clip; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err; done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
for w in funasr kaldi; do
  echo "== $w"
  ncu --clock-control none --metrics $M -k regex:"frontend_kernel|colstat" -c 2 python bench.py --workload $w --no-cpu --no-e2e --steps 1 --warmup 3 2>&1 | grep -E "^  [a-z_A-Z].*\(|^    [a-z]" | sed -E "s/\(FrontendParams.*//; s/, Context.*//"
done > gpurun_out/final_secondary_metrics.txt 2>&1
echo "== istft_kokoro (128 clips)" >> gpurun_out/final_secondary_metrics.txt
ncu --clock-control none --metrics $M -k regex:"istft_kernel" -c 1 python bench.py --workload istft_kokoro --batch 128 --no-cpu --no-e2e --steps 1 --warmup 3 2>&1 | grep -E "^  [a-z_A-Z].*\(|^    [a-z]" >> gpurun_out/final_secondary_metrics.txt
for f in gpurun_out/final_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('final_bench_')[1][:-5], 'ms', round(d.get('ms_per_step',0),4), 'value %.4g'%d['value'], 'frac', round((d.get('roofline') or {}).get('frac',0) or 0,4), 'e2e %.4g'%(e.get('value') or 0), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'launches', d.get('gpu_launches'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as ex: print(sys.argv[1], 'ERR', ex)
PY
done
