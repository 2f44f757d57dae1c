#!/usr/bin/env python3
"""Build the shipped assets from the reference's data files (run in the build container only).

/root/reference does not exist on the GPU box, so the compiled model tables and the raw
motion frames travel with the repo:
  * assets/dp_env_v3.model.npz   -- ModelTables compiled by deepmimic_mujoco_b200.mjcf from
        /root/reference/src/mujoco/humanoid_deepmimic/envs/asset/dp_env_v3.xml
  * assets/motions/<clip>.npz    -- the numeric "Frames" array + "Loop" of
        /root/reference/src/mujoco/motions/humanoid3d_<clip>.txt (data, not code)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.mjcf import compile_mjcf, save_tables  # noqa: E402

REF = os.environ.get("DMB_REFERENCE", "/root/reference")
XML = os.path.join(REF, "src/mujoco/humanoid_deepmimic/envs/asset/dp_env_v3.xml")
MOT = os.path.join(REF, "src/mujoco/motions")
OUT = os.path.join(ROOT, "deepmimic_mujoco_b200", "assets")


def main():
    os.makedirs(os.path.join(OUT, "motions"), exist_ok=True)
    mt = compile_mjcf(XML)
    save_tables(mt, os.path.join(OUT, "dp_env_v3.model.npz"))
    print("model: nq=%d nv=%d nu=%d nbody=%d ngeom=%d npair=%d nM=%d mass=%.3f" % (
        mt.nq, mt.nv, mt.nu, mt.nbody, mt.ngeom, mt.npair, mt.nM, mt.total_mass()))
    for fn in sorted(os.listdir(MOT)):
        if not fn.endswith(".txt"):
            continue
        with open(os.path.join(MOT, fn)) as f:
            d = json.load(f)
        name = fn[len("humanoid3d_"):-len(".txt")]
        fr = np.array(d["Frames"], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, "motions", name + ".npz"), frames=fr, loop=np.array(str(d.get("Loop", "wrap"))))
        print("clip %-16s %4d frames dt=%.6f" % (name, fr.shape[0], fr[0, 0]))


if __name__ == "__main__":
    main()
