# 2-GPU call: multi-GPU parity tests, then weak-scaling bench with and without the gather (NCCL and fused peer stores)
set -x
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for g in fused nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --gather $g > gpurun_out/bench_n${N}_gather_$g.json 2> gpurun_out/bench_n${N}_gather_$g.err
  tail -c 700 gpurun_out/bench_n${N}_gather_$g.json; tail -3 gpurun_out/bench_n${N}_gather_$g.err
done
