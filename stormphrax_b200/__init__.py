"""stormphrax_b200 -- a B200-native batched NNUE evaluator behind Stormphrax's eval API.

Only the hot path lives here: network upload, feature extraction, feature-transformer
accumulator refresh / incremental update, pairwise activation, int8 L1 and int32 L2/L3.
See DESIGN.md for the path and its boundary, include/sp_nnue.h for the C-ABI.
"""
__version__ = "0.1.0"
