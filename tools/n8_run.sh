# 8-GPU measurement set: multi-GPU parity test, config3 weak scaling in both assembly modes, config5 strong scaling.
set -x
TAG=${1:-r2n}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_n8_gputest_multi.log 2>&1; tail -3 gpurun_out/${TAG}_n8_gputest_multi.log
for m in auto p2p; do
GG_BANDS=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_n8_$m.json 2> gpurun_out/${TAG}_n8_$m.err
done
GG_BANDS=auto python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --workload config5 > gpurun_out/${TAG}_n8_c5.json 2> gpurun_out/${TAG}_n8_c5.err
python - $TAG <<'PY'
import json, sys
tag = sys.argv[1]
for f in (tag + "_n8_auto", tag + "_n8_p2p", tag + "_n8_c5"):
    try:
        d = json.load(open("gpurun_out/" + f + ".json"))
        print(f, d["config"]["band_assembly"][:40], d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"], 1), {k: round(v, 3) for k, v in d["config"]["stage_ms"].items()}, d.get("frame_ok"), "e2e", round(d["e2e"]["ms_per_frame"], 3), d["e2e"].get("resident", {}).get("ms_per_frame"))
    except Exception as e:
        print(f, "ERR", e)
PY
