#!/bin/bash
# Refresh at the final HEAD of round 2 (after the fixed-point slicing kernel and the Newton-Schulz driver changes):
# product + slicing kernel under `ncu --set full`, and the C4 launch list.  Summarised on the box by summarize_r2.py.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
TC_VARIANTS="8,10" timeout 300 $NCU -k regex:'ozaki_gemm_kernel|slice_rows_kernel' -c 4 -o $O/ncu_r2_tc python tests/run_tc_gemm.py 2000 > $O/ncu_r2_tc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/launches_r2_c4.csv python tests/run_configs.py c4 > $O/launches_r2_c4.log 2>&1
python profiles/summarize_r2.py $O/summaries_r2_final
rm -f $O/ncu_r2_tc.ncu-rep
cat $O/summaries_r2_final/ncu_tc_gemm_r2.md | tail -4
head -16 $O/summaries_r2_final/launches_r2_c4_summary.md
