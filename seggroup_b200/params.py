"""Parameter set of SegModel (seggroup/model.py:65-166, 676-681) as a flat dict keyed like the reference
state_dict, for the functional pipeline, the bench and the tests."""
from __future__ import annotations

import torch

TRAINABLE = ["mlp_1.bn1.weight", "mlp_1.bn1.bias", "mlp_1.conv1.0.weight",
             "mlp_2.bn1.weight", "mlp_2.bn1.bias", "mlp_2.conv1.0.weight",
             "gcn_2.fc.weight",
             "mlp_3.bn1.weight", "mlp_3.bn1.bias", "mlp_3.conv1.0.weight",
             "mlp_3.bn2.weight", "mlp_3.bn2.bias", "mlp_3.conv2.0.weight",
             "gcn_3.fc.weight",
             "classifier.linear1.weight", "classifier.bn1.weight", "classifier.bn1.bias",
             "classifier.linear2.weight", "classifier.linear2.bias"]


def init_params(seed: int = 1, bn_gamma_scale: float | None = None) -> dict:
    """Default initialisation in the construction order of SegModel.__init__ (model.py:676-681), so
    `torch.manual_seed(seed)` gives the reference's initial weights; `bn_gamma_scale` multiplies
    mlp_1.bn1.weight (the calibrated weight sets of SURVEY.md 8d)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.ones(c)
        sd[prefix + ".bias"] = torch.zeros(c)

    bn("mlp_1.bn1", 64); sd["mlp_1.conv1.0.weight"] = nn.Conv2d(6, 64, 1, bias=False).weight.detach().clone()
    bn("mlp_2.bn1", 64); sd["mlp_2.conv1.0.weight"] = nn.Conv2d(18, 64, 1, bias=False).weight.detach().clone()
    sd["gcn_2.fc.weight"] = nn.Linear(192, 192, bias=False).weight.detach().clone()
    bn("mlp_3.bn1", 64); sd["mlp_3.conv1.0.weight"] = nn.Conv2d(18, 64, 1, bias=False).weight.detach().clone()
    bn("mlp_3.bn2", 64); sd["mlp_3.conv2.0.weight"] = nn.Conv2d(64, 64, 1, bias=False).weight.detach().clone()
    sd["gcn_3.fc.weight"] = nn.Linear(256, 256, bias=False).weight.detach().clone()
    sd["classifier.linear1.weight"] = nn.Linear(256, 128, bias=False).weight.detach().clone()
    bn("classifier.bn1", 128)
    l2 = nn.Linear(128, 40)
    sd["classifier.linear2.weight"] = l2.weight.detach().clone(); sd["classifier.linear2.bias"] = l2.bias.detach().clone()
    if bn_gamma_scale is not None:
        sd["mlp_1.bn1.weight"] = sd["mlp_1.bn1.weight"] * bn_gamma_scale
    return sd
