#!/bin/bash
# ncu_export.sh REPORT OUTPREFIX: the three source-page csv exports the budget tools read (runs on CPU)
ncu -i $1 --page source --csv --print-source cuda,sass > $2_both.csv 2>/dev/null
ncu -i $1 --page source --csv --print-source cuda > $2_cuda.csv 2>/dev/null
ncu -i $1 --page source --csv --print-source sass > $2_sass.csv 2>/dev/null
