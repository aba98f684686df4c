for d in 0 1 2; do echo dbg $d; OM_DBG=$d bash tools/gpu/run10.sh; done
