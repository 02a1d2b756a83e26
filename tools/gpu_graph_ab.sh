for g in 1 0; do
  RADE_B200_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('graph=$g N=8 value %.4g  %.4f ms/step' % (d['value'], d['ms_per_step']))"
done
RADE_B200_GRAPH=1 timeout 120 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('graph=1 N=1 value %.4g  %.4f ms/step' % (d['value'], d['ms_per_step']))"
