for m in 4 5 6; do echo MINB=$m; OM_A3_FEAT_MINB=$m python tools/bench_a3.py --steps 30 | grep -o '"task_kernel_ms": [0-9.]*'; done
