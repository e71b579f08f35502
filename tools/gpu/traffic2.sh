# dram bytes per launch of the dominant kernel of every bench workload at its full batch size (for roofline.traffic): gpurun_out/traffic_<w>.csv
mkdir -p gpurun_out
for spec in "whisper128:frontend_kernel" "istft_hift:istft_kernel" "istft_kokoro:istft_kernel" "funasr:frontend_kernel" "kaldi:frontend_kernel" "s3gen:frontend_kernel" "whisper80_1clip:frontend_kernel" "chatterbox128:frontend_kernel" "voice_encoder:frontend_kernel" "stft_kokoro:small_stft_kernel" "stft_hift:small_stft_kernel" "hift_head:istft_kernel" "whisper128_f16:frontend_kernel"; do
  w=${spec%%:*}; k=${spec##*:}
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -c 1 --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --no-cpu --no-e2e --no-secondary --steps 1 --warmup 3 > /dev/null 2>&1
  python - "$w" <<'PY'
import csv,sys
w=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/traffic_{w}.csv")) if len(r)>10 and r[0].isdigit()]
v={}
for r in rows:
    val=float(r[-1].replace(",","")); u=r[-2]
    if "byte" in u.lower():
        mult={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
        v[r[-3]]=val*mult
    else: v[r[-3]]=(val,u)
print(w, int(v.get("dram__bytes_read.sum",0)+v.get("dram__bytes_write.sum",0)), v.get("gpu__time_duration.sum"))
PY
done
