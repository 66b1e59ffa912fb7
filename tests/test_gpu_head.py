"""GPU parity: fused pose head (Procrustes + yaw nullification) vs the oracle, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import rotation as orot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(cuda_lib):
    e = cuda_lib.Engine(0, max_batch=1, crop_hw=32)
    yield e
    e.close()


def test_procrustes_matches_svd_oracle(eng):
    g = torch.Generator().manual_seed(9)
    m = torch.randn(4096, 9, generator=g)
    m[:1024] *= 0.05                                       # random-init heads give small 9-vectors
    m[1] = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, -1])     # reflection: det < 0
    m[2] = torch.tensor([2., 0, 0, 0, 1, 0, 0, 0, 0.5])
    R, _ = eng.pose_head(m.cuda(), want_yaw=False)
    torch.cuda.synchronize()
    R = R.cpu()
    want = orot.procrustes_to_rotmat(m.double())           # fp64 SVD is the tighter reference
    sv = torch.linalg.svdvals(m.reshape(-1, 3, 3).double())
    well = ((sv[:, 1] + torch.sign(torch.det(m.reshape(-1, 3, 3).double())) * sv[:, 2]) / sv[:, 0]) > 1e-3
    ang = orot.geodesic_deg(R.numpy(), want.numpy())
    assert ang[well.numpy()].max() < 1e-2
    assert torch.allclose(torch.det(R), torch.ones(R.shape[0]), atol=1e-5)
    eye = torch.eye(3).expand(R.shape[0], 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5)
    want32 = orot.procrustes_to_rotmat(m)                  # the fp32 torch path of the reference
    assert np.median(orot.geodesic_deg(R.numpy(), want32.numpy())) < 1e-3


def test_degenerate_inputs_still_return_rotations(eng):
    m = torch.zeros(4, 9)
    m[1, 0] = 1.0                                          # rank 1
    m[2] = torch.tensor([1., 2, 3, 2, 4, 6, 3, 6, 9])      # rank 1 symmetric
    m[3] = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 0])      # rank 2: unique answer = identity
    R, Ry = eng.pose_head(m.cuda())
    torch.cuda.synchronize()
    R = R.cpu()
    assert torch.isfinite(R).all() and torch.isfinite(Ry.cpu()).all()
    assert torch.allclose(torch.det(R), torch.ones(4), atol=1e-5)
    assert torch.allclose(R[3], torch.eye(3), atol=1e-6)


def test_yaw_nullification_matches_reference_golden(eng, golden_dir):
    g = np.load(os.path.join(golden_dir, "yaw.npz"))
    Ry = eng.nullify_yaw(torch.from_numpy(g["R"]).reshape(-1, 9).cuda())
    torch.cuda.synchronize()
    got = Ry.cpu().numpy()
    assert got.dtype == np.float64 and got.shape == g["R_yaw_nullified"].shape
    assert np.abs(got - g["R_yaw_nullified"]).max() < 1e-6
    assert np.abs(got[:, 0, 1]).max() == 0.0               # yaw-free rotations have R01 == 0


def test_python_mirrors(golden_dir):
    from flope_b200 import conversion, mvg
    g = np.load(os.path.join(golden_dir, "yaw.npz"))
    out = mvg.nullify_yaw_batch(g["R"])
    assert out.dtype == np.float64 and np.abs(out - g["R_yaw_nullified"]).max() < 1e-6
    r9 = torch.randn(5, 9, generator=torch.Generator().manual_seed(1))
    R = conversion.procrustes_to_rotmat(r9.cuda())
    assert R.shape == (5, 3, 3)
    assert orot.geodesic_deg(R.cpu().numpy(), orot.procrustes_to_rotmat(r9).numpy()).max() < 1e-2
