#!/bin/bash
# compute-sanitizer memcheck over the smoke path (flow stage n=2 with the fused up-sampling chain kernel and the TMA-staged FB kernel,
# 7x7 PPM / loss kernels, sparse correspondence, TMA-fed tcgen05 GEMM)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_m_plain.log 2>&1 || { tail -5 gpurun_out/r02_m_plain.log; exit 1; }
timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_m_memcheck.log 2>&1
echo "exit $?"; tail -12 gpurun_out/r02_m_memcheck.log
