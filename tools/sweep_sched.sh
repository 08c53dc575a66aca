#!/bin/bash
# usage: tools/sweep_sched.sh "ENV1=.. ENV2=.." ...   one line per configuration: sustained TFLOP/s of the filter
# kernel (tools/score_bench.py, 2M-row corpus) and its DRAM traffic / L2 hit rate from a one-launch ncu read-out.
CMD="python tools/score_bench.py --nv ${SWEEP_NV:-2000000} --nq ${SWEEP_NQ:-8192} --thr ${SWEEP_THR:-0.0683} --cap 16384 --reps 3"
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg $CMD --sustain 1.0 > gpurun_out/ss_$i.log 2>&1
  env $cfg ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpc__cycles_elapsed.max.per_second --clock-control none -k regex:score_ -s 3 -c 1 --csv --log-file gpurun_out/ss_ncu_$i.csv $CMD > /dev/null 2>&1
  python - "$cfg" gpurun_out/ss_$i.log gpurun_out/ss_ncu_$i.csv <<'PY'
import sys, csv
cfg, log, ncu = sys.argv[1:4]
sus = [l.strip() for l in open(log) if "sustained" in l]
m = {}
try:
    rows = [l for l in open(ncu) if not l.startswith("==")]
    for r in csv.DictReader(rows):
        m[r["Metric Name"]] = r["Metric Value"] + r["Metric Unit"]
except Exception as e:
    m = {"ncu": str(e)}
print("[%s] %s | ncu: dram_rd %s lts_bytes %s time %s hit %s tensor %s clk %s" % (
    cfg, " ".join(sus)[:170], m.get("dram__bytes_read.sum"), m.get("lts__t_bytes.sum"), m.get("gpu__time_duration.sum"),
    m.get("lts__t_sector_hit_rate.pct"), m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    m.get("gpc__cycles_elapsed.max.per_second")), flush=True)
PY
done
