# ncu headline metrics of the secondary front-end kernels (Fun-ASR, Kaldi, S3Gen) and the adjacent-row kernels
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
for w in funasr kaldi s3gen whisper_segment whisper128_ragged; do
  echo "== $w"
  ncu --clock-control none --metrics $M -k regex:"frontend_kernel|colstat|mel_segment|zero_tail" -c 3 python bench.py --workload $w --no-cpu --no-e2e --steps 1 --warmup 3 2>&1 | grep -E "^  [a-z_A-Z].*\(|^    [a-z]" | sed -E "s/\(FrontendParams.*//; s/, Context.*//"
done > gpurun_out/secondary_metrics.txt 2>&1
wc -l gpurun_out/secondary_metrics.txt
