#!/bin/bash
# power_trace.sh OUT.csv CMD... — run CMD while sampling SM clock, instantaneous board power and throttle reasons (20 ms)
OUT=$1; shift
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw.instant,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 20 > $OUT &
SMI=$!
sleep 0.5
"$@"
sleep 0.3
kill $SMI
