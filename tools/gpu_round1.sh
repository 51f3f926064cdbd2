#!/bin/bash
# First GPU bring-up: every group in its own process so a trap in one cannot poison the others.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
PT="python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 900"
echo "=== G0 diag"; timeout 300 python tools/umma_diag.py > gpurun_out/g0_diag.log 2>&1; echo "rc=$?"
echo "=== G1 simt coarse"; timeout 900 $PT -k "mutual_matching or ksize1 or simtcorr" > gpurun_out/g1_coarse_simt.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g1_coarse_simt.log
echo "=== G2 simt refine"; timeout 900 $PT -k "simt33" > gpurun_out/g2_refine_simt.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g2_refine_simt.log
echo "=== G3 umma gemm"; timeout 600 $PT -k "umma_gemm" > gpurun_out/g3_umma.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g3_umma.log
echo "=== G4 tc refine + corr"; timeout 900 $PT -k "tc33 or tc31 or tc11 or tccorr or ragged or empty" > gpurun_out/g4_tc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g4_tc.log
echo "=== G5 e2e"; timeout 1500 $PT -k "end_to_end or golden or full_size" > gpurun_out/g5_e2e.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/g5_e2e.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
