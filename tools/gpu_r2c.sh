#!/bin/bash
# parity of the default build, timing with the in-kernel timeline, then the launch list of two composes under ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2c_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r2c_parity.log
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=rustfst_b200/librustfst_b200_$v.so
  echo "== variant '${v:-default}'"
  B200_LIB=$lib B200_COOP_TRACE=1 timeout 300 python tools/profile_run.py --reps 3 > gpurun_out/r2c_${v:-default}.log 2>&1; echo "rc=$?"
  grep "^\[ws\]" gpurun_out/r2c_${v:-default}.log | tail -1
  tail -1 gpurun_out/r2c_${v:-default}.log | cut -c1-400
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python tools/profile_run.py --reps 2 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[]
for l in csv.reader(open('gpurun_out/r2c_launches.csv', errors='ignore')):
    if len(l) > 10 and l[0].isdigit():
        rows.append(l)
# header positions: find Kernel Name and Metric Value columns
hdr=None
for l in csv.reader(open('gpurun_out/r2c_launches.csv', errors='ignore')):
    if 'Kernel Name' in l: hdr=l; break
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
half=rows[len(rows)//2:]
for r in half:
    print(r[kn][:70].ljust(70), r[mv], r[mu])
PY
