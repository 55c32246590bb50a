for d in 0 8 24 1; do echo "DBG=$d"; PFASR_GEMM_DBG=$d python scripts/gemm_probe.py 2>&1 | grep -E "tile (256|0) " | grep -E " (2048|1536) +512|  512 +(512|2048)" ; done
