"""Oracle (test infrastructure only) for the point-set / mesh-regulariser ops of SURVEY 8f rank 4: plain numpy /
torch-CPU restatements of pytorch3d.loss.chamfer_distance, mesh_edge_loss, mesh_laplacian_smoothing(uniform) and
mesh_normal_consistency as published upstream (recalled; PyTorch3D is not vendored -> parity unpinned, like the
rest of oracle/).  Loops are written for clarity, sizes in the tests are small."""
import numpy as np
import torch


def chamfer(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x (N,P1,3), y (N,P2,3), float64 -> scalar: mean over batch of mean_i min_j |x_i-y_j|^2 + mean_j min_i."""
    d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)
    return (d.min(2)[0].mean(1) + d.min(1)[0].mean(1)).mean()


def unique_edges(faces: np.ndarray):
    e = set()
    for a, b, c in faces.tolist():
        for u, v in ((a, b), (b, c), (c, a)):
            e.add((min(u, v), max(u, v)))
    return sorted(e)


def edge_loss(verts: np.ndarray, faces: np.ndarray, target: float = 0.0) -> float:
    e = unique_edges(faces)
    return float(np.mean([(np.linalg.norm(verts[a] - verts[b]) - target) ** 2 for a, b in e]))


def laplacian_uniform(verts: np.ndarray, faces: np.ndarray) -> float:
    nb = {i: set() for i in range(len(verts))}
    for a, b in unique_edges(faces):
        nb[a].add(b); nb[b].add(a)
    tot = 0.0
    for i in range(len(verts)):
        if nb[i]:
            tot += np.linalg.norm(np.mean([verts[j] for j in nb[i]], axis=0) - verts[i])
        else:
            tot += np.linalg.norm(verts[i])     # upstream keeps L_ii = -1 for a vertex without neighbours
    return float(tot / len(verts))


def normal_consistency(verts: np.ndarray, faces: np.ndarray) -> float:
    by_edge = {}
    for fi, (a, b, c) in enumerate(faces.tolist()):
        for u, v, o in ((a, b, c), (b, c, a), (c, a, b)):
            by_edge.setdefault((min(u, v), max(u, v)), []).append(o)
    vals = []
    for (u, v), opp in by_edge.items():
        for i in range(len(opp)):
            for j in range(i + 1, len(opp)):
                e = verts[v] - verts[u]
                n0 = np.cross(e, verts[opp[i]] - verts[u])
                n1 = -np.cross(e, verts[opp[j]] - verts[u])
                vals.append(1.0 - float(n0 @ n1) / max(np.linalg.norm(n0) * np.linalg.norm(n1), 1e-8))
    return float(np.mean(vals)) if vals else 0.0
