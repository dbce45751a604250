"""OFF / COFF meshes and GOFF Gaussian files (reference VoGE/Converter/IO.py).

GOFF layout: line 0 "GOFF", line 1 "<n> <sigma width: 1|3|6|9> <has radians: 0|1>", then n vertex lines,
n sigma lines and (optionally) n radian lines."""
import numpy as np
import torch


def _floats(lines):
    return np.array(" ".join(lines).split(), dtype=np.float32)


def load_off(file_name, to_torch=False, ignore_color=False):
    """-> (verts (n,3) f32, faces (m,k) i32[, vert_color][, face_color]); reference IO.py:11-57."""
    with open(file_name) as f:
        lines = f.readlines()
    head = lines[0].strip()
    if ignore_color or head.startswith("OFF"):
        colored = False
    elif head.startswith("COFF"):
        colored = True
    else:
        raise Exception("Unsupported OFF format: %s" % head)
    counts = lines[1].split()
    n_points, n_faces = int(counts[0]), int(counts[1])
    verts = _floats(lines[2:2 + n_points]).reshape(n_points, -1)
    out = [verts[:, 0:3], None]
    if colored and verts.shape[1] > 3:
        out.append(verts[:, 3:])
    faces = np.array(" ".join(lines[2 + n_points:2 + n_points + n_faces]).split(), dtype=np.float64)
    faces = faces.reshape(n_faces, -1)
    k = int(faces[0][0])
    out[1] = faces[:, 1:k + 1].astype(np.int32)
    if colored and faces.shape[1] > k + 1:
        out.append(faces[:, k + 1:].astype(np.float32))
    return tuple(torch.from_numpy(np.ascontiguousarray(t)) for t in out) if to_torch else tuple(out)


def load_goff(file_name, to_torch=False):
    """-> (points (n,3), sigma (n,w) | (n,3,3) | split pair for w = 6, radians | None); reference IO.py:60-87."""
    with open(file_name) as f:
        lines = f.readlines()
    n_points, l_sigma, if_radian = (int(v) for v in lines[1].split()[:3])
    points = _floats(lines[2:2 + n_points]).reshape(-1, 3)
    sigma = _floats(lines[2 + n_points:2 + 2 * n_points]).reshape(-1, l_sigma)
    if l_sigma == 6:
        sigma = (sigma[:, :3], sigma[:, 3:])
    elif l_sigma == 9:
        sigma = sigma.reshape(-1, 3, 3)
    radian = _floats(lines[2 + 2 * n_points:]) if if_radian else None
    if not to_torch:
        return points, sigma, radian
    tt = lambda a: tuple(torch.from_numpy(x) for x in a) if isinstance(a, tuple) else torch.from_numpy(a)
    return torch.from_numpy(points), tt(sigma), torch.from_numpy(radian) if radian is not None else None


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t


def save_off(file_name, vertices, faces, vert_color=None, face_color=None):
    vertices, faces, vert_color, face_color = _np(vertices), _np(faces), _np(vert_color), _np(face_color)
    rows = ["OFF" if vert_color is None and face_color is None else "COFF", "%d %d 0" % (vertices.shape[0], faces.shape[0])]
    for i, v in enumerate(vertices):
        vals = list(v[:3]) + (list(vert_color[i]) if vert_color is not None else [])
        rows.append(" ".join("%.16f" % x for x in vals))
    for i, f in enumerate(faces):
        row = "%d " % len(f) + " ".join("%d" % x for x in f)
        if face_color is not None:
            row += " " + " ".join("%.16f" % x for x in face_color[i])
        rows.append(row)
    with open(file_name, "w") as fl:
        fl.write("\n".join(rows) + "\n")


def save_goff(file_name, points, sigmas, radians=None):
    if isinstance(sigmas, tuple):
        sigmas = np.concatenate([_np(s) for s in sigmas], axis=1)
    points, sigmas, radians = _np(points), _np(sigmas), _np(radians)
    sigmas = sigmas.reshape(sigmas.shape[0], -1)
    rows = ["GOFF", "%d %d %d" % (points.shape[0], sigmas.shape[1], 0 if radians is None else 1)]
    rows += [" ".join("%.16f" % x for x in v) for v in points]
    rows += [" ".join("%.16f" % x for x in v) for v in sigmas]
    if radians is not None:
        rows += ["%.16f" % v for v in radians]
    with open(file_name, "w") as fl:
        fl.write("\n".join(rows) + "\n")


def to_torch(*args):
    return [torch.from_numpy(t).type(torch.float32) if t is not None else None for t in args]


def pre_process_pascal(verts, *args):
    """PASCAL3D+ axis convention (x, y, z) -> (x, z, -y); reference IO.py:170-175."""
    if torch.is_tensor(verts):
        verts = torch.cat((verts[:, 0:1], verts[:, 2:3], -verts[:, 1:2]), dim=1)
    else:
        verts = np.concatenate((verts[:, 0:1], verts[:, 2:3], -verts[:, 1:2]), axis=1)
    return (verts,) + args
