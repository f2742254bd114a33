import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lio_slam_b200 import synth
from lio_slam_b200.liogpu import LioGpu
from oracle.oracle import Oracle
o = Oracle("port")
world = synth.make_world(1234)
clouds, poses = [], []
for k in range(50):
    p = synth.path_pose(0.5 * k)
    ds, _ = o.voxel_grid(synth.to_packed(synth.make_scan(world, p, 32, seed=700 + k, cols=900)), 0.4)
    clouds.append(ds); poses.append(p.astype(np.float32))
poses = np.array(poses)
g = LioGpu()
for k, c in enumerate(clouds):
    g.keyframe_put(k, c)
for _ in range(2):
    out, info, st = g.publish_local_map(list(range(50)), poses, poses[-1], local_mapping_surf_leaf_size=0.2)
tgt, _ = g.merge_keyframes(list(range(50)), poses, 0.4)
wrong = poses[25].copy(); wrong[3] += 0.7; wrong[4] -= 0.6; wrong[2] += 0.03
src = o.transform_cloud(clouds[25], wrong)
for _ in range(2):
    T, ii = g.icp_align(src, tgt)
print(info["n_cropped"], ii["iterations"], src.shape, tgt.shape)
