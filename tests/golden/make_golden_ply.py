#!/usr/bin/env python
"""Golden map files written by the REFERENCE's own PLY writer (oracle/_ref/ply_ref = /root/reference/src/tinyply.h driven
like GaussianModel::saveMap).  Run in the authoring container (needs /root/reference); commits tests/golden/map_*.ply and the
inputs they were made from."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def model(P, M, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    f = lambda *s: rng.normal(size=s).astype(np.float32)
    return dict(xyz=f(P, 3), f_dc=f(P, 1, 3), f_rest=f(P, M, 3), opacity=f(P), scale=f(P, 3), rotation=f(P, 4))


def blob(m, P, M):
    return (np.array([P, M], np.uint32).tobytes() +
            b"".join(np.ascontiguousarray(m[k], np.float32).tobytes() for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rotation")))


if __name__ == "__main__":
    from oracle.ref_build.build_ply_ref import build
    exe = build()
    assert exe and os.path.isfile(exe), "needs /root/reference (tinyply.h)"
    for name, P, M, seed in (("map_deg3", 11, 15, 1), ("map_deg0", 5, 0, 2), ("map_empty", 0, 15, 3)):
        m = model(P, M, seed)
        tmp = os.path.join(HERE, name + ".blob")
        open(tmp, "wb").write(blob(m, P, M))
        subprocess.check_call([exe, tmp, os.path.join(HERE, name + ".ply")])
        os.remove(tmp)
        np.savez(os.path.join(HERE, name + "_inputs.npz"), P=P, M=M, **m)
        print(name, os.path.getsize(os.path.join(HERE, name + ".ply")), "bytes")
