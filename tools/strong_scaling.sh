N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
for w in mobilenetv2 yolov3tiny candy; do
  $TR bench.py --gpus $N --workload $w --scaling strong --steps 30 --warmup 5 --no-extra --no-cpu-baseline 2>gpurun_out/r02_strong_${w}_n$N.err | tail -1 > gpurun_out/r02_strong_${w}_n$N.json
  python -c "
import json;d=json.load(open('gpurun_out/r02_strong_${w}_n$N.json'));print('$w',d['n_gpus'],d['scaling'],d['config']['batch_per_gpu'],d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"
done
