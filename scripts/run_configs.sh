#!/bin/bash
# Every BASELINE configuration through bench.py on this box (development aid):
#   scripts/run_configs.sh TAG NGPUS "CONFIGS" [extra bench.py flags]
# writes gpurun_out/TAG_config<k>_n<N>.json (+ .err).
TAG=$1; N=$2; CONFIGS=$3; shift 3
mkdir -p gpurun_out
for k in $CONFIGS; do
  out=gpurun_out/${TAG}_config${k}_n${N}
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --config $k --steps 3 --warmup 3 "$@" > $out.json 2> $out.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29520 + k)) bench.py --gpus $N --config $k --steps 3 --warmup 3 "$@" > $out.json 2> $out.err
  fi
  echo "config $k rc=$? $(python - <<PY
import json
try:
    d = json.loads(open('$out.json').read().strip().splitlines()[-1])
    e = d.get('e2e') or {}
    print('value %.0f  ms/step %.2f  e2e %s  roofline.frac %.3f fp32 %.3f kernel_ms %.3f launches %s' % (
        d['value'], d['ms_per_step'], ('%.0f' % e['value']) if e else None, d['roofline']['frac'],
        d['roofline']['fp32']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
except Exception as ex:
    print('no line:', ex)
PY
)"
  tail -3 $out.err
done
