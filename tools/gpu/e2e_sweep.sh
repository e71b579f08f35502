for mb in 16 32 64 128 192 384; do
  B2A_HOST_CHUNK_MB=$mb python bench.py --workload ${WL:-whisper128} --no-cpu --steps 3 --warmup 3 --e2e-steps 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_mb', $mb, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'audio-s/s %.3g' % d['e2e']['value'])"
done
