bash tools/ab_variants.sh 'pf*' 2>&1 | tee gpurun_out/r01b_ab_prefetch.txt
bash tools/ab_thin.sh 2>&1 | tee gpurun_out/r01b_ab_thin.txt
bash tools/prof_variant.sh pf1 r01b_ncu_pf1 headline 4 kolb_pool2
bash tools/prof_variant.sh ../lib r01b_ncu_thin config3 4 thin_persistent
