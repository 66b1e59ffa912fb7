import sys, cProfile, pstats, io
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from flope_b200 import _lib, synth
from flope_b200.posenet import PoseResNet
from flope_b200.predictor import FastPosePredictor
sd = synth.random_state_dict(0)
frames_np, masks_np, det, b5_np = bench.frame_batch(np, _lib, synth, 1, 8, synth.FRAME_SEED + 5)
net = PoseResNet(device="cuda:0", max_batch=8, crop_hw=224); net.load_state_dict(sd)
pred = FastPosePredictor("cuda:0", detector=lambda rgb: (det[0].astype(np.int16), masks_np[0]), posenet=net, crop_hw=224, interp=_lib.INTERP_LINEAR)
rgb = frames_np[0]
for _ in range(50): pred.get_flower_poses(rgb, None)
pr = cProfile.Profile(); pr.enable()
for _ in range(500): pred.get_flower_poses(rgb, None)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
