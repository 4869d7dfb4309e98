#!/bin/bash
# A/B of the tile tick kernel (engine_tile.inl) for environment / library variants: bench value + ncu kernel durations
out=${1:-gpurun_out/tile_ab.jsonl}
: > $out
run() {  # label, workload, env...
  label=$1; shift; wl=$1; shift
  env "$@" python bench.py --workload $wl --no-ess --no-cpu --no-secondary --steps 6 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'label':'$label','wl':'$wl','value':d['value'],'ms_per_step':d['ms_per_step'],'grad_ms':d['roofline']['avg_launch_ms'],'elem_frac':d['roofline_elementwise']['frac'],'accept':d['config'].get('mean_accept'),'mhz':d['clocks']['sm_mhz']}))" | tee -a $out
}
# kernel durations (ncu, serialised, cold cache): avg us of the tile kernel per variant
nk() {  # label, workload, env...
  label=$1; shift; wl=$1; shift
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tile_tick -s 8 -c 24 --csv python bench.py --workload $wl --no-ess --no-cpu --no-secondary --steps 2 --warmup 1 2>/dev/null | \
    python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>5 and r[0].isdigit()]
v=[float(r[-1]) for r in rows]
import json; print(json.dumps({'label':'$label','wl':'$wl','ncu_avg_us':(sum(v)/len(v)/1000 if v else None),'n':len(v),'kernel':rows[0][4][:60] if rows else None}))" | tee -a $out
}
if [ "$2" = "tests" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $out; fi
nk w16 c2 B2H_TILE_TICK=1
nk w12 c2 B2H_LIB=build/lib_w12/libb200hmc.so
nk w8 c2 B2H_LIB=build/lib_w8/libb200hmc.so
nk w16_wpc8 c2 B2H_TILE_WPC=8
