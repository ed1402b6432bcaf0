"""Seeded synthetic inputs shared by fixtures, parity tests and the bench (TEST/BENCH helper).

SURVEY.md section 8(d): raw = uniform [0,1) packed-Bayer (B,4,T,T); cond = uniform (B,4,256,256)
for single tiles; coord = 2-channel normalised pixel-centre coordinates, channel 0 = x in [-1,1]
along W, channel 1 = y in [-1,1] along H (the reference smoke tests feed randn triples of the same
shapes, LiteISP.py:2670-2672).
"""
import torch


def coord_map(T, B=1, y0=-1.0, y1=1.0, x0=-1.0, x1=1.0):
    ys = torch.linspace(y0, y1, T)
    xs = torch.linspace(x0, x1, T)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([xx, yy])[None].repeat(B, 1, 1, 1).contiguous()


def make_inputs(T, seed=1234, B=1, cond_size=256):
    g = torch.Generator().manual_seed(seed)
    raw = torch.rand(B, 4, T, T, generator=g)
    cond = torch.rand(B, 4, cond_size, cond_size, generator=g)
    return [raw, cond, coord_map(T, B)]
