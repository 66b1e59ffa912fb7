"""CPU oracle for the FloPE batched flower-pose inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``flope_b200/`` may import this
package.  The only legal importers are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only
as the checker / the CPU arm, never as the thing shipped.

The reference (wvu-irl/flope) is pure Python, so the oracle is Python too
(numpy / torch-CPU fp32 / cv2 / scipy - the reference's own dependencies):

  oracle.boxes     squarify_bb / bb_in_frame / filter_very_large_bb
                   (sunflower/utils/mvg.py:324-362)
  oracle.resize    the crop-batch loop (sunflower/predictor/pose_predictor.py:138-153,
                   fast_pose_predictor.py:108-123) on the real cv2, plus an integer
                   restatement of cv2's uint8 INTER_LANCZOS4 / INTER_LINEAR resize
  oracle.posenet   PoseResNet (sunflower/models/posenet.py:5-34), eval mode, fp32
  oracle.rotation  procrustes_to_rotmat (sunflower/utils/conversion.py:54-58),
                   nullify_yaw_batch (sunflower/utils/mvg.py:240-251), Rt assembly
  oracle.depth     get_depth_value / shrink_mask (sunflower/utils/image_manipulation.py:21-96),
                   get_points3d (sunflower/utils/mvg.py:387-408) - the depth / translation branch
  oracle.detector_post  the post-processing half of FastPosePredictor.get_bbox_mask
                   (sunflower/predictor/fast_pose_predictor.py:48-57)
  oracle.pipeline  the composed path frame+mask+boxes(+depth, K) -> (N,4,4)

Pinning status (also in DESIGN.md):
  * boxes, PoseResNet, nullify_yaw_batch, procrustes_to_rotmat's reshape: PINNED
    against the reference's own functions imported from /root/reference
    (tests/golden/make_golden.py; fixtures committed under tests/golden/).
  * get_depth_value / shrink_mask / get_points3d: PINNED against the reference's own
    functions (tests/golden/depth.npz; matplotlib / plotly stubbed for the import).
  * get_bbox_mask post-processing: PINNED against the reference method itself run with a stub
    detector (tests/golden/yolo_post.npz).
  * cv2.resize arithmetic: pinned against the cv2 build in this image (4.13.0;
    the reference pins 4.10.0.84).
  * roma.special_procrustes (roma==1.5.1, environment.yml:214) is a third-party
    dependency absent from /root/reference and from this image, and the
    reference has no test or golden vector for it: PARITY UNPINNED for that one
    function.  It is restated from its documented behaviour
    (R = U diag(1,1,det(U V^T)) V^T from the SVD of the 3x3).
  * The reference ships no tests, no golden vectors and no fixtures for this
    path at all; every fixture here was produced by running
    the reference code in the authoring container.
  * Deliberate deviation: the oracle runs PoseNet in eval() mode under
    no_grad().  As written the reference predictors never call .eval(), which
    leaves dropout and batch-statistics BatchNorm live and the output random
    (SURVEY.md section 0, D4); parity is only definable in eval mode.
"""
