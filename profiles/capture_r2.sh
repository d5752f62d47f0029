#!/bin/bash
# Round-2 ncu captures, one gpurun call (1 GPU).  Outputs land in gpurun_out/ and are summarised here by
# `python profiles/summarize_r2.py` into profiles/*.md.  Numbers printed by runs under ncu are never bench values.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
# 1. tensor-core product + slicing kernel (PSD projection, N = 2000)
TC_VARIANTS="8,10" timeout 300 $NCU -k regex:'ozaki_gemm_kernel|slice_rows_kernel' -c 3 -o $O/ncu_r2_tc python tests/run_tc_gemm.py 2000 > $O/ncu_r2_tc.log 2>&1
# 2. config C4: projection / rhs kernel at m = 2e6, load / scale / store kernels of the projection, SpMV epilogues
timeout 600 $NCU -k regex:'proj_rhs_kernel|psd_large_load_kernel|ns_scale_kernel|ns_store_kernel|recover_mu_kernel|wx_update_kernel' -s 6 -c 12 -o $O/ncu_r2_c4 python tests/run_configs.py c4 > $O/ncu_r2_c4.log 2>&1
# 3. config C2: CG vector kernels, residual epilogues (first termination check at iteration 1), windowed SpMV
timeout 600 $NCU -k regex:'cg_init|cg_update|spmv_win_kernel|proj_rhs_kernel' -s 4 -c 40 -o $O/ncu_r2_c2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_r2_c2.log 2>&1
# 4. config C5 (|V| = 3000 to keep the host-side chordal analysis short): batched small-cone Jacobi
timeout 600 $NCU -k regex:'psd_small_kernel' -s 2 -c 3 -o $O/ncu_r2_c5 python tests/run_c5.py 3000 5 > $O/ncu_r2_c5.log 2>&1
# 5. block-Jacobi fallback (N = 1000)
COSMO_B200_PSD_TC=0 timeout 600 $NCU -k regex:'bj_pivot_kernel|bj_cols_kernel|bj_rows_kernel' -s 30 -c 6 -o $O/ncu_r2_bj python tests/run_psd_sign_timing.py 1000 > $O/ncu_r2_bj.log 2>&1
# launch lists (shares of a step): C2 bench and C4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1500 --csv --log-file $O/launches_r2_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/launches_r2_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/launches_r2_c4.csv python tests/run_configs.py c4 > $O/launches_r2_c4.log 2>&1
# summarise on the box (only 64 MiB of gpurun_out/ travel back) and keep just the product kernel's report
python profiles/summarize_r2.py $O/summaries_r2
ls -la $O/*.ncu-rep $O/launches_r2_*.csv
rm -f $O/ncu_r2_c4.ncu-rep $O/ncu_r2_c2.ncu-rep $O/ncu_r2_c5.ncu-rep $O/ncu_r2_bj.ncu-rep
ls $O/summaries_r2
