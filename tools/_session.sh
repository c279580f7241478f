D=gpurun_out/${1:-x1}; mkdir -p $D
for i in 1 2 3; do
  timeout 900 python bench.py --steps 20 --warmup 5 > $D/clk_$i.json 2> $D/clk_$i.err
  timeout 900 python bench.py --steps 20 --warmup 5 --no-clocks > $D/noclk_$i.json 2> $D/noclk_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/x1/*.json')):
    try:
        l=json.loads([x for x in open(f).read().splitlines() if x.startswith('{')][-1])
        v=l['also']['cfg4']
        print(f.split('/')[-1],'cfg3',round(l['value']),'cfg4',round(v['value']),'host',v['host_ms_per_step'][:3],'cpu',v['host_cpu_ms_per_step'][:3],'bracket',v['roofline']['avg_launch_ms'])
    except Exception as e: print(f,'ERR',e)
P
