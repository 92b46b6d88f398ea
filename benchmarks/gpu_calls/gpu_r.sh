#!/bin/bash
# round 2, GPU call R: ncu evidence for every kernel of the path (one --set full capture each, after a plain run exits 0);
# the reports are summarised ON the box (table row + hottest source lines) and only the small ones travel back
mkdir -p gpurun_out
timeout 600 python benchmarks/ncu_kernels.py C3 200000 10 > gpurun_out/r_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/r_plain.log; exit 1; }
for k in k_score_rows k_score_rows_tc k_accum_feature k_prop_score k_hyper_features k_cat_sweep_spec k_cat_sweep_flat; do
  extra="-c 2"
  case $k in k_cat_sweep_spec) extra="-c 3";; k_accum_feature) extra="-c 3";; esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k(<|\$)" $extra -f -o /tmp/r_$k python benchmarks/ncu_kernels.py C3 200000 10 > gpurun_out/r_$k.log 2>&1
  python benchmarks/ncu_table.py /tmp/r_$k.ncu-rep > gpurun_out/r_$k.table.md 2>/dev/null
  ncu -i /tmp/r_$k.ncu-rep --page source --csv --print-source cuda,sass > /tmp/r_$k.src.csv 2>/dev/null
  python benchmarks/ncu_lines.py /tmp/r_$k.src.csv 14 > gpurun_out/r_$k.lines.txt 2>/dev/null
  ncu -i /tmp/r_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_shared_atom.sum']
idx=[h.index(k) for k in keep if k in h]
w=csv.writer(sys.stdout)
for row in r: w.writerow([row[i] for i in idx])
" > gpurun_out/r_$k.raw.csv
  sz=$(stat -c %s /tmp/r_$k.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -lt 6000000 ]; then cp /tmp/r_$k.ncu-rep gpurun_out/; fi
  echo "$k size=$sz"; tail -n 1 gpurun_out/r_$k.table.md | cut -c1-300
done
# the launch list of the default bench command's step on a reduced block (per-launch times: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r_launches.csv python bench.py --rows 100000 --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r_launches.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/r_launches.csv) lines"
du -sh gpurun_out
