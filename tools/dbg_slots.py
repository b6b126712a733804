"""Debug aid: slot refresh + update of a few dozen items through the GENERAL kernels (SP_NNUE_SMALL=0), compared with the full refresh."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SP_NNUE_SMALL"] = "0"
from stormphrax_b200 import api, net as N
net = N.synthetic(1234)
boards, _m, starts = api.playouts(3, 64, 40, threads=2)
first = starts[:-1].astype(np.int64)
n = len(first)
with api.Nnue(net.image, 0) as ctx:
    want = ctx.eval_full(boards)
    ctx.slots_reserve(2 * n)
    ids = np.arange(n, dtype=np.uint32)
    ctx.refresh(ids, boards[first])
    a = ctx.eval_slots(ids)
    b = ctx.update_eval(ids, ids + n, boards[first + 1])
    print("refresh mismatches", int((a != want[first]).sum()), "update mismatches", int((b != want[first + 1]).sum()))
