#!/usr/bin/env python
"""The slots workload of bench.py alone (sp_nnue_batch_device: update src level -> dst level + evaluate, 65,536 chains),
for A/B runs of library variants.  usage: [SP_NNUE_LIB=...] python tools/prof_slots.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.cuda.set_device(0)
g = bench.Gpu(0)
boards, _moves, starts = bench.make_workload(0, 0, "playouts")
rec, bad = bench.extra_slots(g, 0, 1, boards, starts, steps)
print(f"slots: {rec['value']:.1f} M update+eval/s  ({rec['ms_per_round']*1e3:.1f} us per round of {bench.SLOT_STATES})  e2e {rec['e2e']['value']:.1f}  mismatches {bad}  kernels ms/step {rec['kernel_ms_per_step']}")
