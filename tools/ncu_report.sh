#!/bin/bash
# usage: tools/ncu_report.sh <rep> <out.txt> [min_pct]  -- headline metrics, opcode mix, stall reasons, hot source lines
rep=$1; out=$2; pct=${3:-1.5}
{
python tools/ncu_summary.py "$rep" 2>/dev/null
echo; echo "stall reasons (warps per issued instruction):"
ncu -i "$rep" --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for v in rows[2:]:
    print(' kernel:', v[h.index('Kernel Name')][:90])
    for i,n in enumerate(h):
        if ('issue_stalled' in n and 'per_issue_active' in n) or 'smsp__warps_eligible.avg.per' in n or 'smsp__warps_active.avg.per' in n or 'average_warp_latency' in n:
            try:
                if float(v[i] or 0)>0.05: print('  %-90s %s'%(n, v[i]))
            except ValueError: pass
"
echo; echo "source lines with >= $pct % of the executed instructions or of the stall samples:"
python tools/ncu_lines.py "$rep" "$pct" 2>/dev/null
} > "$out"
