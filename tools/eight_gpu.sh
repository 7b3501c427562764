# usage (GPU box with 8 GPUs): bash tools/eight_gpu.sh   -> gpurun_out/g8_*.json
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 5 --no-cpu > gpurun_out/g${n}_bench.json 2> gpurun_out/g${n}_bench.err
done
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads([l for l in open(f'gpurun_out/g{n}_bench.json') if l.startswith('{')][-1])
        print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
        for r in d['e2e']['per_rank']: print('   ', r)
    except Exception as e:
        print(n, 'ERR', e)
PY
nvidia-smi topo -m > gpurun_out/g8_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/g8_topo.txt
