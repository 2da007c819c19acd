"""FP64 pipe microbenchmarks on the GPU box: DFMA chains and the DMMA shapes PTX offers (all lower to DMMA.8x8x4 SASS on sm_100a)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cafe5_b200.model import measure_fp64_peak
for name, k in (("dfma", 0), ("dmma m8n8k4", 1), ("dmma m16n8k8", 2), ("dmma m16n8k16", 3)):
    print("%-14s %.2f TFLOP/s" % (name, measure_fp64_peak(0, k)))
for w in (4, 8):
    print("dmma m8n8k4, %2d warps/SM (8 independent tiles per warp): %.2f TFLOP/s" % (w, measure_fp64_peak(0, 1 | (w << 8))))
