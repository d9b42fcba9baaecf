"""Timing experiment: one evaluation with a library built with -DFWI_UMMA_TRACE=1 (tools/build_variant.sh); CTA (0,0) prints the SM
clock at the MMA-issuer / epilogue handshake events of a few accumulator uses.   FWI_VARIANT_LIB=... python tools/umma_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import _lib
if os.environ.get("FWI_VARIANT_LIB"):
    _lib.LIB_PATH = os.environ["FWI_VARIANT_LIB"]
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import mc_oracle as orc
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
prob = fw.SourceInversion(d, G)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
metric = sys.argv[2] if len(sys.argv) > 2 else "VR"
fl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
M = torch.randn((9, N), device="cuda")
prob.eval_dev(M, fw.METRICS.index(metric), fl | 16)
torch.cuda.synchronize()
