for p in 1 0; do
CMC_P2P=$p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$p bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['roofline']['per_direction']
print('p2p=$p', {k:round(v['ms_per_launch'],3) for k,v in p.items()}, 'ms/step', round(d['ms_per_step'],2), d['implementation']['exchange'])
"
done
