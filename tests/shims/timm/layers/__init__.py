"""AttentionPoolLatent with the constructor utils/mim_vit.py:247-250 uses (only built when attn_pool=True)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class AttentionPoolLatent(nn.Module):
    def __init__(self, in_features, num_heads=8, mlp_ratio=4.0, norm_layer=nn.LayerNorm):
        super().__init__()
        self.num_heads = num_heads
        self.latent = nn.Parameter(torch.zeros(1, 1, in_features))
        self.q = nn.Linear(in_features, in_features)
        self.kv = nn.Linear(in_features, in_features * 2)
        self.proj = nn.Linear(in_features, in_features)
        self.norm = norm_layer(in_features)
        self.fc1 = nn.Linear(in_features, int(in_features * mlp_ratio))
        self.fc2 = nn.Linear(int(in_features * mlp_ratio), in_features)

    def forward(self, x):
        B, N, C = x.shape
        h = self.num_heads
        q = self.q(self.latent.expand(B, -1, -1)).reshape(B, 1, h, C // h).transpose(1, 2)
        k, v = self.kv(x).reshape(B, N, 2, h, C // h).permute(2, 0, 3, 1, 4).unbind(0)
        y = self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, 1, C))
        y = y + self.fc2(F.gelu(self.fc1(self.norm(y))))
        return y[:, 0]
