"""flope_b200 - B200-native implementation of FloPE's batched flower-pose inference path.

Public surface (mirrors the reference, wvu-irl/flope):
  flope_b200.posenet.PoseResNet            sunflower/models/posenet.py
  flope_b200.predictor.FastPosePredictor   sunflower/predictor/fast_pose_predictor.py
  flope_b200.predictor.PosePredictor       sunflower/predictor/pose_predictor.py
  flope_b200.mvg / flope_b200.conversion   the hot-path members of sunflower/utils/{mvg,conversion}.py
  flope_b200.image_manipulation            get_depth_value / shrink_mask of sunflower/utils/image_manipulation.py
  flope_b200.flower_model.FlowerModel      sunflower/predictor/flower_model.py
  flope_b200.aggregate.Env3D               scripts/flower_pose_aggregrator.py
  flope_b200.shard / flope_b200.pipeline   multi-GPU sharding, several steps in flight on one GPU
Everything computes through the C ABI of libflope_b200.so (include/flope_b200.h).
"""
__version__ = "0.1.0"
