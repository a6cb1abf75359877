set -x
mkdir -p gpurun_out
nvidia-smi -L
GG_BENCH_WATCHDOG_S=150 GGCUDA_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/n1_trace.json 2> gpurun_out/n1_trace.log; echo rc=$?
tail -12 gpurun_out/n1_trace.log
GG_BENCH_WATCHDOG_S=150 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2.json 2> gpurun_out/n2.log; echo rc=$?
cat gpurun_out/n2.json | cut -c1-600
tail -40 gpurun_out/n2.log
