for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-other-configs > gpurun_out/scale_$n.json 2>gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --no-other-configs > gpurun_out/scale_$n.json 2>gpurun_out/scale_$n.err; fi
  tail -1 gpurun_out/scale_$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value %.1f G' % (d['value']/1e9), 'ms/step %.3f' % d['ms_per_step'], 'e2e %.1f G' % (d['e2e']['value']/1e9), {k[:12]:round(v['ms'],3) for k,v in d['roofline']['kernels'].items()}, d.get('path_sharded_parity',{}).get('rel_vs_single'), d.get('path_sharded_parity',{}).get('identical_on_all_ranks'))"
done
