#!/usr/bin/env python
"""Convert the reference's data/M3500.txt (VERTEX2/EDGE2 text) into tests/golden/m3500.npz.

Parsing follows the demo loader (/root/reference/examples/aprilsam_demo.c:52-99): EDGE2
columns are IDout IDin dx dy dth I11 I12 I22 I33 I13 I23 and the loader writes them to
W[0],W[1],W[4],W[8],W[2],W[5] (upper triangle only, lower left stays 0).
Run in the build container only (the reference tree does not exist on the GPU box).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aprilsam_b200.harness import PoseGraphData  # noqa: E402


def parse(path: str) -> PoseGraphData:
    init, ea, eb, ez, eW = [], [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "VERTEX2":
                assert int(t[1]) == len(init)
                init.append([float(t[2]), float(t[3]), float(t[4])])
            elif t[0] == "EDGE2":
                ea.append(int(t[1]))
                eb.append(int(t[2]))
                ez.append([float(t[3]), float(t[4]), float(t[5])])
                i11, i12, i22, i33, i13, i23 = (float(x) for x in t[6:12])
                eW.append([i11, i12, i13, 0.0, i22, i23, 0.0, 0.0, i33])
            else:
                raise ValueError(t[0])
    return PoseGraphData(np.array(init, dtype=np.float64), np.array(ea, dtype=np.int32), np.array(eb, dtype=np.int32),
                         np.array(ez, dtype=np.float64), np.array(eW, dtype=np.float64))


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/M3500.txt"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "m3500.npz")
    d = parse(src)
    d.save(dst)
    print(f"{d.n_nodes} nodes, {d.n_edges} edges -> {dst}")
