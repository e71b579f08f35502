#!/bin/bash
# usage: tools/prof_report.sh <tag> <mangled kernel substring> <line ranges> [tiles]
# Reads gpurun_out/prof_frontend_<tag>.ncu-rep; prints headline metrics + per-region instruction / stall breakdown.
set -e
cd /root/repo/gpurun_out
ncu -i prof_frontend_$1.ncu-rep --page source --csv --print-source sass > src_$1.csv 2>/dev/null
ncu -i prof_frontend_$1.ncu-rep --page raw --csv > raw_$1.csv 2>/dev/null
mkdir -p /tmp/cub; cd /tmp/cub && rm -f *.cubin *.dis && cuobjdump -xelf all /root/repo/mlx_swift_audio_b200/libb200audio.so >/dev/null 2>&1; nvdisasm -g -c frontend.sm_100a.cubin > frontend.dis 2>/dev/null
cd /root/repo
python - <<PY
import csv
rows=list(csv.reader(open('/root/repo/gpurun_out/raw_$1.csv')))
d=dict(zip(rows[0],rows[2]))
for k in ['gpu__time_duration.sum','smsp__inst_issued.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__grid_size','dram__bytes_read.sum','dram__bytes_write.sum']+[h for h in rows[0] if 'issue_stalled' in h and 'per_issue_active' in h]:
    v=d.get(k)
    try:
        if float(v)<0.05: continue
    except: pass
    print(f"{k:95s} {v}")
print('instr per tile', float(d['smsp__inst_issued.sum'])/${4:-24000})
PY
python tools/prof_by_line.py gpurun_out/src_$1.csv /tmp/cub/frontend.dis "$2" 8 "$3" | tail -30
