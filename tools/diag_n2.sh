set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -12; echo rc=$?
for mode in p2p p2p_nomc nccl; do
GG_BANDS=$mode GG_BENCH_WATCHDOG_S=150 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.log; echo rc=$?
tail -n 2 gpurun_out/n2_$mode.log | cut -c1-300
done
python - <<'PY'
import json
for m in ("p2p","p2p_nomc","nccl"):
    try:
        d=json.loads(open(f"gpurun_out/n2_{m}.json").read().strip().splitlines()[-1])
        print(m, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), d["config"]["band_assembly"], {k:round(v,3) for k,v in d["config"]["stage_ms"].items()})
    except Exception as e: print(m, "unreadable", e)
PY
