# round 2 session AM: heave1024 on one GPU -- what the IBM grid beside the update costs (lean build, fewer blocks, no overlap)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --workload heave1024 --steps 100 --warmup 10 --no-cpu-baseline --no-parity-check "$@" > gpurun_out/r02am_heave1024_$tag.json 2> gpurun_out/err_am_$tag.txt; echo "$tag rc=$?"; }
run base
run lean --opt ibm_early_lean=1
run b74 --opt ibm_early_blocks=74
run b74lean --opt ibm_early_blocks=74 --opt ibm_early_lean=1
run b296 --opt ibm_early_blocks_per_sm=2
run noearly --opt ibm_early=0
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02am_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), round(r.get('frac'),4), round(r['collide_alone']['kernel_ms'],4), d.get('clocks',{}).get('sm_mhz'), d['details']['structural_solver']['host_ms_per_step_all_bodies'])
    except Exception as e: print(f, 'ERR', e)
P
