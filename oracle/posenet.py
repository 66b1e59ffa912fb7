"""Oracle: PoseResNet, eval mode, fp32 CPU (test infrastructure only).

Restates sunflower/models/posenet.py:5-34: torchvision ResNet-18 trunk with
``avgpool -> AdaptiveAvgPool2d(1)``, ``fc -> Linear(512, 2048) + ReLU``, a second
(idempotent) ReLU, dropout (identity in eval mode) and ``fc_rot = Linear(2048, 9)``.

Differences from the reference as written, both forced by the environment and
both recorded in DESIGN.md:
  * the reference constructor asks torchvision for the IMAGENET1K_V1 weights
    (posenet.py:10), which needs a download; there is no network, so the trunk
    is built with ``weights=None`` (random init under the caller's seed) exactly
    as the north star's "random-init weights" wording asks;
  * the oracle is always in eval() mode under no_grad().

Construction order (trunk, then base.fc, then fc_rot) matches the reference so a
given torch seed produces the same state_dict as the reference class built with
the ``weights=None`` patch; tests/golden/make_golden.py checks that.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.models as tvm


class PoseResNet(nn.Module):
    def __init__(self, backbone_out_dim=2048, dropout=0.5):
        super().__init__()
        self.base = tvm.resnet18(weights=None)                      # posenet.py:10 (weights patched)
        fc_in = self.base.fc.in_features
        self.base.avgpool = nn.AdaptiveAvgPool2d(1)                 # posenet.py:12
        self.base.fc = nn.Sequential(nn.Linear(fc_in, backbone_out_dim), nn.ReLU())   # posenet.py:13-16
        self.fc_rot = nn.Linear(backbone_out_dim, 9)                # posenet.py:19
        self.dropout = dropout

    def extract_features(self, x):                                  # posenet.py:24-29
        f = F.relu(self.base(x))
        if self.dropout > 0:
            f = F.dropout(f, p=self.dropout, training=self.training)
        return f

    def forward(self, x):                                           # posenet.py:31-34
        return self.fc_rot(self.extract_features(x))


def build(seed=0):
    """Seeded random-init PoseResNet in eval mode (weights seed 0)."""
    torch.manual_seed(seed)
    return PoseResNet().eval()


@torch.no_grad()
def forward_fp32(model, x, chunk=64):
    """(B,3,H,W) float32 in [0,1] -> (B,9) float32, CPU."""
    model = model.eval()
    x = torch.as_tensor(x, dtype=torch.float32)
    outs = [model(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)]
    return torch.cat(outs) if outs else torch.zeros(0, 9)


@torch.no_grad()
def trunk_activations(model, x):
    """Named intermediate activations (NCHW fp32) for layer-by-layer parity tests."""
    b = model.base
    acts = {}
    y = b.relu(b.bn1(b.conv1(x)))
    acts["stem"] = y
    y = b.maxpool(y)
    acts["maxpool"] = y
    for li, layer in enumerate([b.layer1, b.layer2, b.layer3, b.layer4], 1):
        for bi, blk in enumerate(layer):
            y = blk(y)
            acts[f"layer{li}.{bi}"] = y
    y = torch.flatten(b.avgpool(y), 1)
    acts["pool"] = y
    y = F.relu(b.fc(y))
    acts["feat"] = y
    acts["r9"] = model.fc_rot(y)
    return acts
