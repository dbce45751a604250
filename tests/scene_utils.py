"""Seeded small scenes shared by CPU and GPU tests."""
import math

import numpy as np
import torch


def small_scene(seed=0, n=300, image_size=(40, 48), K=8, focal=60.0, dist=4.0, aniso=True, views=1):
    """Random blob of Gaussians in front of `views` cameras; returns a dict of CPU tensors."""
    import voge_oracle as vo
    g = torch.Generator().manual_seed(seed)
    verts = (torch.rand(n, 3, generator=g) * 2 - 1) * 0.8
    s = torch.exp(torch.rand(n, generator=g) * 1.5 + 3.0)          # inverse covariance 20..90
    sig = torch.zeros(n, 3, 3)
    sig[:, 0, 0] = s; sig[:, 1, 1] = s; sig[:, 2, 2] = s
    if aniso:
        A = torch.randn(n, 3, 3, generator=g) * 0.25
        M = torch.eye(3)[None] + A
        sig = (M @ M.transpose(1, 2)) * s.view(-1, 1, 1)            # SPD, non-trivial off-diagonals
        sig[: n // 4] += torch.randn(n // 4, 3, 3, generator=g) * 0.3   # a few NON-symmetric ones
    H, W = image_size
    azim = torch.linspace(20, 200, views)
    elev = torch.linspace(10, -25, views)
    R, T = vo.look_at_view(dist, elev, azim)
    return dict(verts=verts, sigmas=sig, R=R, T=T, focal=focal, principal=(W / 2.0 - 1.5, H / 2.0 + 0.75),
                image_size=image_size, K=K, colors=torch.rand(n, 3, generator=g))
