#!/bin/bash
# Round-2 GPU call 5 (1 GPU): bench JSON; ncu of every other kernel, summarised on the box (the report itself is too big to travel).
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c5_bench_n1.json 2> gpurun_out/r02_c5_bench_n1.err; echo "bench n1 exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c5_bench_ref.json 2> gpurun_out/r02_c5_bench_ref.err; echo "ref exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"fill|heights|transition|gather|meshlet|publish|visibility|brick|edit|records|pack|commit" \
    -o /tmp/r02_aux_kernels -f python tools/profile_kernels.py > gpurun_out/r02_c5_ncu_aux.log 2>&1; echo "ncu aux exit $?"
python tools/summarize_ncu.py --multi /tmp/r02_aux_kernels.ncu-rep gpurun_out/r02_aux_kernels_ncu.txt > /dev/null; echo "summary exit $?"
ncu -i /tmp/r02_aux_kernels.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=[h.index(k) for k in ('ID','Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes.sum.per_second') if k in h]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] if i<len(r) else '' for i in keep])
" > gpurun_out/r02_aux_kernels_launches.csv
ls -la gpurun_out | tail -8; du -sh gpurun_out
