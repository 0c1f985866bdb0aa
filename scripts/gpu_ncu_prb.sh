#!/bin/bash
T=${1:-r02_ac}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prb_pass -s 7 -c 1 -o gpurun_out/${T}_prbpass -f python scripts/time_prb.py 512 8 1 > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_prbpass.ncu-rep --page source --csv > gpurun_out/${T}_prbpass_source.csv 2>/dev/null
ncu -i gpurun_out/${T}_prbpass.ncu-rep --page details --csv > gpurun_out/${T}_prbpass_details.csv 2>/dev/null
ls -la gpurun_out/${T}_prbpass*; rm -f gpurun_out/${T}_prbpass.ncu-rep
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_prbpass_source.csv')))
print(rows[0][:12])
PY
