"""Vertex-normal oracle — TEST INFRASTRUCTURE ONLY.

Restates the JIT branch of `Mesh::recompute_vertex_normals` (src/render/mesh.cpp:311-330: angle-weighted face normals,
Thuermer & Wuethrich, scattered to the vertices and normalised) in torch, so that autograd gives the adjoint the
reference obtains from Dr.Jit's AD when `params.update()` re-derives the normals from AD-attached positions
(mesh.cpp:82-87).  Pinned: the forward values are held to the reference's own `recompute_vertex_normals` (its built
`scalar_rgb` module, tests/golden/reference_vectors.npz `normals/*`) by tests/test_vertex_normals.py."""
import numpy as np
import torch


def recompute_vertex_normals(pos: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """pos (V,3) float, faces (F,3) int64 → unit vertex normals (V,3); differentiable w.r.t. pos."""
    v = [pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]]
    n = torch.cross(v[1] - v[0], v[2] - v[0], dim=-1)
    n = n / torch.linalg.norm(n, dim=-1, keepdim=True)                               # mesh.cpp:318
    normals = torch.zeros_like(pos)
    for i in range(3):
        d0 = v[(i + 1) % 3] - v[i]; d0 = d0 / torch.linalg.norm(d0, dim=-1, keepdim=True)
        d1 = v[(i + 2) % 3] - v[i]; d1 = d1 / torch.linalg.norm(d1, dim=-1, keepdim=True)
        x = (d0 * d1).sum(-1)
        ang = torch.acos(torch.clamp(x, -1, 1))                                        # dr::safe_acos, mesh.cpp:324
        nn = n * ang[:, None]
        nn = torch.where(torch.isfinite(nn).all(-1, keepdim=True), nn, torch.zeros_like(nn))
        normals = normals.index_add(0, faces[:, i], nn)                                # scatter_reduce(Add), mesh.cpp:327-328
    ln = torch.linalg.norm(normals, dim=-1, keepdim=True)
    bogus = torch.tensor([1.0, 0.0, 0.0], dtype=pos.dtype)
    return torch.where(ln > 0, normals / torch.where(ln > 0, ln, torch.ones_like(ln)), bogus)


def normals_backward(pos: np.ndarray, faces: np.ndarray, grad_nrm: np.ndarray, dtype=torch.float64) -> np.ndarray:
    """Jᵀ·grad_nrm by autograd."""
    p = torch.tensor(np.asarray(pos), dtype=dtype, requires_grad=True)
    n = recompute_vertex_normals(p, torch.tensor(np.asarray(faces).astype(np.int64)))
    (n * torch.tensor(np.asarray(grad_nrm), dtype=dtype)).sum().backward()
    return p.grad.numpy()
