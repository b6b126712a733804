import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stormphrax_b200 import api, net as N
net = N.synthetic(1234)
kw = dict(concurrency=96, total_games=150, depth=3, nodes_per_move=400, max_plies=50, seed=21)
dev, sd = api.selfplay(net.image, 0, resident=True, **kw)
print("resident ok", sd)
