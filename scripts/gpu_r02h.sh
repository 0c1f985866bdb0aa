#!/bin/bash
T=${1:-r02_h}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_prb.csv python scripts/time_prb.py 512 8 1 > gpurun_out/${T}_ncu_list.log 2>&1
python - <<'PY'
import csv,re,collections,sys
rows=list(csv.reader(open('gpurun_out/'+sys.argv[1]+'_launches_prb.csv'))) if len(sys.argv)>1 else []
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prb_pass -s 3 -c 3 -o gpurun_out/${T}_prbpass_full -f python scripts/time_prb.py 512 8 1 > gpurun_out/${T}_ncu_prbpass.log 2>&1
ncu -i gpurun_out/${T}_prbpass_full.ncu-rep --page details --csv > gpurun_out/${T}_prbpass_details.csv 2>&1
tail -3 gpurun_out/${T}_ncu_list.log
