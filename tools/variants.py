#!/usr/bin/env python
"""Build the tuning variants of libsp_nnue.so (compile-time knobs at the top of kernels.cu / selfplay_gpu.cu) into
stormphrax_b200/_lib/variants/<name>.so.  They travel to the GPU box with `gpurun` and are selected with
SP_NNUE_LIB=<path> (tests, bench and tools all load the library through stormphrax_b200.api).
usage: python tools/variants.py [name ...]        (no names: all)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stormphrax_b200 import build as B

VARIANTS = {
    # dense head: the pre-ldmatrix L1 loop (A quads packed from two LDS.128), and three warps per scheduler
    "head_lds": {"SP_HEAD_LDMATRIX": 0},
    "head_c12": {"SP_HEAD_CONSUMERS": 12},
    # board enumeration: next round's ray fetched under the current round
    "enq_prefetch": {"SP_ENQ_PREFETCH": 1},
    "enq_unroll2": {"SP_ENQ_UNROLL": 2},
    # slot update kernel (self-play path): CTAs per SM
    "slots_static": {"SP_SLOTS_TICKET": 0},
    "slots_ticket1": {"SP_SLOTS_TICKET": 1},
    "slots3": {"SP_SLOTS_MIN_BLOCKS": 3},
    "slots4": {"SP_SLOTS_MIN_BLOCKS": 4},
    # tensor-core full refresh: CTA 0 prints its clocks per phase
    "group_timing": {"SP_GROUP_TIMING": 1},
    "group_timing_unroll2": {"SP_GROUP_TIMING": 1, "SP_ENQ_UNROLL": 2},
    "group_unroll2": {"SP_ENQ_UNROLL": 2},
    "group_unroll8": {"SP_ENQ_UNROLL": 8},
    # tcgen05 head: which side bounds it?  (results are wrong by construction: timing only)
    "umma_notail": {"SP_HEAD_UMMA_NOTAIL": 1},
    "umma_nogather": {"SP_HEAD_UMMA_NOGATHER": 1},
    "umma_k1": {"SP_HEAD_UMMA_KATOMS": 1},
    "umma_loaders3": {"SP_HEAD_UMMA_LOADERS": 3},
    "umma_loaders5": {"SP_HEAD_UMMA_LOADERS": 5},
    "umma_notail_loaders3": {"SP_HEAD_UMMA_NOTAIL": 1, "SP_HEAD_UMMA_LOADERS": 3},
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    out_dir = os.path.join(B.LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    for name in names:
        path = os.path.join(out_dir, name + ".so")
        B.build(defines=VARIANTS[name], out=path)
        print(path)


main()
