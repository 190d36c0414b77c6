mkdir -p gpurun_out
for bps in 0 1; do
FSILBM_BPS=$bps timeout 400 python - <<P > gpurun_out/r02x_school2048r_bps$bps.json 2> gpurun_out/err_x$bps.txt
import os, sys, subprocess
sys.argv = ['bench.py', '--workload', 'school2048r', '--steps', '100', '--warmup', '10', '--no-cpu-baseline', '--no-parity-check']
import fsilbm3d_b200 as F
F.lib(); 
import bench
_orig = bench.run_gpu
def run(args):
    F._lib.ensure_init(0)
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_early_blocks_per_sm", int(os.environ['FSILBM_BPS'])))
    return _orig(args)
bench.run_gpu = run
bench.main()
P
echo "rc=$?"
done
for w in heave1024 plate512; do timeout 400 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02x_$w.json 2>/dev/null; done
python - <<'P'
import json
for t in ('school2048r_bps0','school2048r_bps1','heave1024','plate512'):
    try:
        d=json.load(open(f'gpurun_out/r02x_{t}.json')); r=d['roofline']
        print(t, round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), d['clocks']['sm_mhz'], (d.get('parity_check') or {}).get('ok'))
    except Exception as e: print(t, 'ERR', e)
P
