#!/bin/bash
# usage: tools/sweep.sh "ENV1=.. ENV2=.." ...   (each arg = one configuration; 2M-row corpus, quick ncu DRAM read-out)
CMD="python bench.py --nv 2000000 --steps 3 --warmup 3 --skip-cpu-baseline"
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg $CMD > gpurun_out/sw_$i.log 2>&1
  env $cfg ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpc__cycles_elapsed.max.per_second --clock-control none -k regex:score_ -s 7 -c 1 --csv --log-file gpurun_out/sw_ncu_$i.csv $CMD > /dev/null 2>&1
  python - "$cfg" gpurun_out/sw_$i.log gpurun_out/sw_ncu_$i.csv <<'PY'
import sys, json, csv
cfg, log, ncu = sys.argv[1:4]
d = json.loads(open(log).read().strip().splitlines()[-1])
m = {}
try:
    rows = [l for l in open(ncu) if not l.startswith("==")]
    for r in csv.DictReader(rows):
        m[r["Metric Name"]] = r["Metric Value"] + r["Metric Unit"]
except Exception as e:
    m = {"ncu": str(e)}
print("[%s] q/s %d ms %.1f filt_ms %.1f TF %d frac %.3f clk %s | ncu: dram_rd %s time %s hit %s tensor %s clk %s" % (
    cfg, d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["achieved"], d["roofline"]["frac"],
    d["clocks"].get("sm_mhz_min_max"), m.get("dram__bytes_read.sum"), m.get("gpu__time_duration.sum"),
    m.get("lts__t_sector_hit_rate.pct"), m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    m.get("gpc__cycles_elapsed.max.per_second")))
PY
done
