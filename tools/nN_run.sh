N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2t_n${N}_auto.json 2> gpurun_out/r2t_n${N}_auto.err
python -c "
import json
d=json.load(open('gpurun_out/r2t_n${N}_auto.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value'],1), d['config']['stage_ms'], d.get('frame_ok'), d['config']['band_assembly'][:50], round(d['e2e']['ms_per_frame'],3), d['e2e']['resident']['ms_per_frame'])" || tail -5 gpurun_out/r2t_n${N}_auto.err
