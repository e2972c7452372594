set -u
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_temporal_bwd_mma2 -s 1 -c 1 -f -o gpurun_out/r01m_conv_bwd2 python tools/ncu_step.py 2 > gpurun_out/r01m_ncu.log 2>&1
ls -la gpurun_out/r01m_conv_bwd2.ncu-rep
