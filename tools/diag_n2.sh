set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
GGCUDA_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.log; echo rc=$?
tail -3 gpurun_out/bench_r1e.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2.json 2> gpurun_out/n2.log; echo rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1e.json","gpurun_out/n2.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], d["config"]["stage_ms"], d["e2e"]["ms_per_frame"], d["config"]["counts"])
PY
