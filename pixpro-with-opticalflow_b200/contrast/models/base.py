import torch.nn as nn


class BaseModel(nn.Module):
    """Base model holding the online encoder (contrast/models/base.py:4-20)."""

    def __init__(self, base_encoder, args):
        super().__init__()
        self.encoder = base_encoder(low_dim=args.feature_dim, head_type=args.head_type)

    def forward(self, x1, x2):
        raise NotImplementedError
