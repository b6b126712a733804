#!/usr/bin/env python
"""Golden values from the reference engine's OWN search (oracle/_ref/sp_engine_cpu = its unmodified search.cpp / bench.cpp /
thread.cpp / position.cpp + stock CPU evaluation, built by oracle/engine/Makefile): `bench` node counts at small depths
(src/bench.cpp:95-170, the engine's own determinism signature) and the eval checksum of the datagen-style playout check.
The same engine sources linked against libsp_nnue.so (sp_engine_b200) must reproduce them on the GPU
(tests/test_engine_dropin.py).

    python tests/golden/make_engine_golden.py      # where /root/reference exists (make -C oracle/engine first)
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stormphrax_b200 import net as N  # noqa: E402

ENGINE = os.path.join(ROOT, "oracle", "_ref", "sp_engine_cpu")
NET_SEED = 7


def run(net_path, *args):
    return subprocess.run([ENGINE, net_path, *map(str, args)], check=True, capture_output=True, text=True).stdout


def main():
    with tempfile.TemporaryDirectory() as tmp:
        net_path = os.path.join(tmp, "tame.nnue")
        N.synthetic(NET_SEED, tame=True).image.tofile(net_path)
        out = {"net": {"seed": NET_SEED, "tame": True}, "bench_nodes": {}, "evalcheck": {}}
        for depth in (1, 2, 3):
            m = re.search(r"^(\d+) nodes (\d+) nps", run(net_path, "bench", depth), re.M)
            out["bench_nodes"][str(depth)] = int(m.group(1))
        m = re.search(r"evalcheck: (\d+) positions, (\d+) mismatches, checksum ([0-9a-f]+)", run(net_path, "evalcheck", 20, 42))
        out["evalcheck"] = {"games": 20, "seed": 42, "positions": int(m.group(1)), "mismatches": int(m.group(2)), "checksum": m.group(3)}
        # independent fixed-depth searches and miniature self-play games: what the fiber scheduler must reproduce in batches
        m = re.search(r"search nodes:((?: \d+)+)", run(net_path, "searches", 24, 3, 0))
        out["searches"] = {"n": 24, "depth": 3, "nodes": [int(x) for x in m.group(1).split()]}
        m = re.search(r"games: .* nodes: (\d+) nodes .* checksum ([0-9a-f]+)", run(net_path, "games", 16, 400, 4, 42, 0))
        out["games"] = {"n": 16, "soft_nodes": 400, "plies": 4, "seed": 42, "nodes": int(m.group(1)), "checksum": m.group(2)}
    with open(os.path.join(ROOT, "tests", "golden", "engine_seed7.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
