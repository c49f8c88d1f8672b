#!/bin/bash
# Round 2, GPU session U (1 GPU): L2 fetch granularity (cudaLimitMaxL2FetchGranularity) 32 / 64 (default) / 128 bytes.
# The per-path state is gathered by 32-byte sectors; if L2 fetches 64 bytes per missed sector, half of what those gathers
# pull from HBM is never used (session L: the MIS top-level pass reads ~300 B of DRAM per ray for ~135 B of records).
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
import ctypes
cudart = ctypes.CDLL("libcudart.so")
v = ctypes.c_size_t(0)
print("cudaLimitMaxL2FetchGranularity default:", cudart.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)
PY
WORKLOAD=c4-1080p timeout 900 tools/ab_env.sh 2 "RAYITO_B200_L2_FETCH=64" "RAYITO_B200_L2_FETCH=32" "RAYITO_B200_L2_FETCH=128" "X=1" > $O/r2u_ab_c4.log 2>&1; cat $O/r2u_ab_c4.log
WORKLOAD=c5-64spp timeout 600 tools/ab_env.sh 1 "RAYITO_B200_L2_FETCH=64" "RAYITO_B200_L2_FETCH=32" "RAYITO_B200_L2_FETCH=128" > $O/r2u_ab_c5.log 2>&1; cat $O/r2u_ab_c5.log
