# bench.py at N GPUs on one box, the way the driver launches it; GG_BANDS modes: p2p (default: multicast), p2p_nomc, nccl
set -x
mkdir -p gpurun_out
N=${1:-8}
for mode in ${2:-p2p p2p_nomc}; do
  GG_BANDS=$mode GG_BENCH_WATCHDOG_S=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n${N}_$mode.json 2> gpurun_out/scale_n${N}_$mode.log; echo rc=$?
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_n*_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), d["config"]["band_assembly"][:40], {k:round(v,3) for k,v in d["config"]["stage_ms"].items()}, "own_max", round(d["config"]["rank_pipeline_ms_max"],3), "own0", round(d["config"]["rank0_pipeline_ms"],3))
    except Exception as e: print(f, "unreadable", e)
PY
