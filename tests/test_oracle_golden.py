"""Pin the oracle (oracle/) against fixtures produced by the real reference code.

The fixtures under tests/golden/ come from tests/golden/make_golden.py, which imports
wvu-irl/flope from /root/reference; these tests run anywhere (no reference needed).
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import boxes as obox
from oracle import posenet as onet
from oracle import rotation as orot
from flope_b200 import synth


def test_squarify_and_in_frame_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "boxes.npz"))
    H, W = g["frame_hw"]
    for dtype in (np.int16, np.int64):
        got = np.array([obox.squarify_bb(b) for b in g["boxes"].astype(dtype)])
        assert np.array_equal(got, g["squarified"])
    keep = np.array([obox.bb_in_frame(s, (H, W, 3)) for s in g["squarified"]])
    assert np.array_equal(keep, g["keep"])
    sq, keep2 = obox.squarify_filter(g["boxes"], (H, W, 3))
    assert np.array_equal(keep2, g["keep"]) and np.array_equal(sq, g["squarified"][g["keep"]])
    # squares are square and contain the original box
    s = g["squarified"]
    assert np.all((s[:, 2] - s[:, 0]) == (s[:, 3] - s[:, 1]))


def test_squarify_reference_samples():
    # samples recorded in SURVEY.md appendix F from the reference
    assert obox.squarify_bb(np.array([10, 20, 110, 51], np.int16)) == [10, -15, 110, 85]
    assert obox.squarify_bb([10, 20, 41, 120]) == [-25, 20, 75, 120]
    assert obox.squarify_bb([10.5, 20.2, 110.7, 51.9]) == [10, -14, 110, 85]


def test_filter_very_large_bb_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "boxes.npz"))
    assert np.array_equal(obox.filter_very_large_bb(g["vlb_in"]), g["vlb_out"])


def test_nullify_yaw_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "yaw.npz"))
    got = orot.nullify_yaw_batch(g["R"].astype(np.float64))
    assert got.dtype == np.float64
    assert np.abs(got - g["R_yaw_nullified"]).max() < 1e-14
    cf = orot.nullify_yaw_closed_form(g["R"])
    assert np.abs(cf - g["R_yaw_nullified"]).max() < 1e-6   # inputs are fp32 rotations (not exactly orthonormal)


def test_posenet_state_dict_and_outputs_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "posenet_seed0.npz"))
    net = onet.build(seed=synth.WEIGHT_SEED)
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["state_keys"])
    assert len(sd) == 124          # 122 trunk tensors (incl. num_batches_tracked) + fc_rot weight/bias
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["state_shapes"])
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    same_bits = np.array_equal(np.frombuffer(h.digest(), np.uint8), g["state_sha"])
    probe = np.concatenate([sd["base.conv1.weight"].flatten()[:16].numpy(),
                            sd["base.fc.0.weight"].flatten()[:16].numpy(), sd["fc_rot.bias"].numpy()])
    assert same_bits or np.allclose(probe, g["probe_weights"], atol=0), "seeded init differs from the reference's"
    for size, nb in ((224, 8),):
        x = synth.mixed_crops(nb, size)
        assert np.array_equal(np.frombuffer(hashlib.sha256(x.numpy().tobytes()).digest(), np.uint8), g[f"in_sha_{size}"])
        r9 = onet.forward_fp32(net, x).numpy()
        assert np.abs(r9 - g[f"r9_{size}"]).max() < 2e-5
        rot = orot.procrustes_to_rotmat(torch.from_numpy(r9)).numpy()
        assert orot.geodesic_deg(rot, g[f"rot_{size}"]).max() < 1e-2


def test_posenet_512_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "posenet_seed0.npz"))
    net = onet.build(seed=synth.WEIGHT_SEED)
    x = synth.mixed_crops(2, 512)
    r9 = onet.forward_fp32(net, x).numpy()
    assert np.abs(r9 - g["r9_512"]).max() < 2e-5


def test_procrustes_properties():
    g = torch.Generator().manual_seed(3)
    m = torch.randn(512, 9, generator=g)
    m[0] = 0.0                                    # degenerate
    m[1] = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, -1])   # reflection
    m[2, 3:] = m[2, :3].repeat(2)                 # rank 1
    r = orot.procrustes_to_rotmat(m)
    assert r.shape == (512, 3, 3)
    assert torch.allclose(torch.det(r), torch.ones(512), atol=1e-4)
    eye = torch.eye(3).expand(512, 3, 3)
    assert torch.allclose(r @ r.transpose(1, 2), eye, atol=1e-4)
    # nearest rotation: trace(R^T M) is not improved by small perturbations
    M = m.reshape(-1, 3, 3)[8:]
    base = torch.einsum('nij,nij->n', r[8:], M)
    from scipy.spatial.transform import Rotation as sciR
    for k in range(4):
        dR = torch.from_numpy(sciR.from_rotvec(0.05 * np.random.default_rng(k).normal(size=(M.shape[0], 3))).as_matrix()).float()
        assert torch.all(torch.einsum('nij,nij->n', dR @ r[8:], M) <= base + 1e-4)


def test_product_random_init_equals_oracle_init():
    sd = synth.random_state_dict(synth.WEIGHT_SEED)
    ref = onet.build(synth.WEIGHT_SEED).state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k


def _davenport_q_method(m):
    """Nearest rotation to M (maximises tr(R^T M)) WITHOUT an SVD: Davenport's q-method - the optimal unit quaternion is
    the eigenvector of the symmetric 4x4 matrix K(M) with the largest eigenvalue (Wahba's problem with attitude profile M).
    An independent second derivation of what oracle.rotation.special_procrustes (R = U diag(1,1,det) V^T) computes."""
    m = np.asarray(m, np.float64)
    out = np.empty_like(m)
    for i, B in enumerate(m):
        S, sig = B + B.T, np.trace(B)
        z = np.array([B[1, 2] - B[2, 1], B[2, 0] - B[0, 2], B[0, 1] - B[1, 0]])
        K = np.zeros((4, 4))
        K[:3, :3] = S - sig * np.eye(3)
        K[:3, 3] = K[3, :3] = z
        K[3, 3] = sig
        w, v = np.linalg.eigh(K)
        q = v[:, -1]
        qv, q4 = q[:3], q[3]
        qx = np.array([[0, -qv[2], qv[1]], [qv[2], 0, -qv[0]], [-qv[1], qv[0], 0]])
        out[i] = (q4 * q4 - qv @ qv) * np.eye(3) + 2 * np.outer(qv, qv) - 2 * q4 * qx
    return out


def test_special_procrustes_agrees_with_an_svd_free_formulation():
    """roma.special_procrustes is absent (parity unpinned, oracle/rotation.py): at least pin the restatement against a
    second, independent formulation of 'the rotation nearest to M' - on random matrices, on small 9-vectors like a
    random-init head produces, on reflections (det < 0) and on rank-2 matrices, where the answer is still unique."""
    rng = np.random.default_rng(12)
    m = rng.standard_normal((600, 3, 3))
    m[:200] *= 0.05
    m[200:300] = m[200:300] @ np.diag([1.0, 1.0, -1.0])                       # plenty of det < 0
    u, s, vt = np.linalg.svd(m[300:400])
    s[:, 2] = 0.0                                                             # rank 2
    m[300:400] = (u * s[:, None, :]) @ vt
    m[400] = np.diag([1.0, 1.0, -1.0])                                        # pure reflection: the nearest rotation is not unique
    want = orot.special_procrustes(torch.from_numpy(m)).numpy()
    got = _davenport_q_method(m)
    sv = np.linalg.svd(m, compute_uv=False)
    gap = sv[:, 1] + np.sign(np.linalg.det(m)) * sv[:, 2]                     # uniqueness margin of the maximiser
    ok = gap / sv[:, 0] > 1e-6
    assert ok.sum() >= 590
    assert orot.geodesic_deg(got[ok], want[ok]).max() < 1e-5
    obj = lambda R: np.einsum('nij,nij->n', R, m)                             # both maximise tr(R^T M), unique or not
    assert np.abs(obj(got) - obj(want)).max() < 1e-9
    assert np.abs(np.linalg.det(want) - 1).max() < 1e-9
