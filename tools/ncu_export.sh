#!/bin/bash
# usage: tools/ncu_export.sh <report.ncu-rep> <out-prefix> <n-launches>
# Exports the raw page and per-launch SASS pages of an ncu report as CSV (run on the GPU box so that
# only the small CSVs have to travel back).
rep=$1; out=$2; n=$3
ncu -i "$rep" --page raw --csv > "${out}_raw.csv" 2>/dev/null
for ((i = 0; i < n; i++)); do
  ncu -i "$rep" --page source --csv --print-source sass --launch-skip $i --launch-count 1 > "${out}_sass_$i.csv" 2>/dev/null
done
