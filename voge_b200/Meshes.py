"""Gaussian parameter containers (reference VoGE/Meshes.py:5-57).  Calling a container returns
`(verts, sigmas, radians)`; `radians` is carried for API compatibility but unused by the renderer
(reference Renderer.py:111)."""
import torch
import torch.nn as nn


class GaussianMeshesNaive:
    """Plain (non-Module) holder of verts (N,3) and sigmas ((N,), (N,3) or (N,3,3))."""

    def __init__(self, verts, sigmas, radians=None):
        self.verts, self.sigmas, self.radians = verts, sigmas, radians

    def to(self, device):
        self.verts = self.verts.to(device)
        self.sigmas = self.sigmas.to(device)
        if self.radians is not None:
            self.radians = self.radians.to(device)
        return self

    def __call__(self):
        return self.verts, self.sigmas, self.radians

    def __getitem__(self, item):
        rad = None if self.radians is None else self.radians[item]
        return GaussianMeshesNaive(self.verts[item], self.sigmas[item], rad)


class GaussianMeshes(nn.Module):
    """nn.Module holder: verts / sigmas / radians are nn.Parameters; `gradianted_args` (name kept
    from the reference) selects which of the three receive gradients."""

    def __init__(self, verts, sigmas, radians=None, gradianted_args=None):
        super().__init__()
        flags = list(gradianted_args) if gradianted_args is not None else [True, True, True]
        self.verts = nn.Parameter(verts, requires_grad=bool(flags[0]))
        self.sigmas = nn.Parameter(sigmas, requires_grad=bool(flags[1]))
        if radians is None:
            self.radians = None
            flags[2] = False
        else:
            self.radians = nn.Parameter(radians, requires_grad=bool(flags[2]))
        self.gradianted_args = flags

    def grad_parameters(self):
        params = (self.verts, self.sigmas, self.radians)
        return tuple(p for p, on in zip(params, self.gradianted_args) if on)

    def forward(self):
        return self.verts, self.sigmas, self.radians


DeformedGaussianMeshes = GaussianMeshes
