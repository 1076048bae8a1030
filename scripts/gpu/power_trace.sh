#!/bin/bash
# board power and SM clock sampled every 50 ms while bench.py runs 20 steps: is the step at the power cap throughout?
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=timestamp,power.draw,power.limit,clocks.sm,clocks.mem,temperature.gpu,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 50 > $O/pw_trace.csv 2>/dev/null &
SMI=$!
sleep 1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-agent > $O/pw_bench.log 2>&1; echo "bench rc=$?"
sleep 1
kill $SMI
python - <<'PY'
import csv, json, statistics as st
rows=[r for r in csv.reader(open("gpurun_out/r2/pw_trace.csv")) if len(r) >= 7]
pw=[float(r[1]) for r in rows]; lim=float(rows[0][2]); clk=[float(r[3]) for r in rows]; cap=[r[6].strip() for r in rows]
busy=[i for i,p in enumerate(pw) if p > 0.6*lim]
print("samples %d (50 ms apart), power limit %.0f W" % (len(rows), lim))
print("samples above 60 %% of the limit: %d; among them power median %.0f W, p10 %.0f W, max %.0f W; SM clock median %.0f MHz (min %.0f, max %.0f); sw_power_cap active in %d"
      % (len(busy), st.median([pw[i] for i in busy]), sorted([pw[i] for i in busy])[len(busy)//10], max(pw), st.median([clk[i] for i in busy]),
         min(clk[i] for i in busy), max(clk[i] for i in busy), sum(1 for i in busy if cap[i].lower().startswith("active"))))
d=json.loads(open("gpurun_out/r2/pw_bench.log").read().strip().splitlines()[-1])
print("bench: %.2f ms/step, %.4e evals/s" % (d["ms_per_step"], d["value"]))
PY
