"""Mints tests/golden/selfocc_*.npz: self-crossing node chains seen by a pinhole camera, the reference's self-occlusion
loop (trackdlo_node.cpp:280-343) run literally on cv2's own raster (cv2.line, thickness dlo_pixel_width), and the visibility
lists that follow from it (:316-360).  Needs opencv-python (the build container has 4.13); the GPU box only reads the files."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import raster
from trackdlo_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def chain(kind, n, rng):
    t = np.linspace(0, 1, n)
    if kind == "loop":            # one loop: the far strand passes under the near one
        x = 0.30 * np.sin(2 * np.pi * t * 1.25) * (0.4 + t); y = 0.22 * np.sin(4 * np.pi * t * 0.9 + 0.4) * (1.1 - 0.5 * t)
        z = 0.62 + 0.05 * np.cos(2 * np.pi * t * 1.25) + 0.02 * t
    elif kind == "eight":         # figure of eight, two crossings
        x = 0.28 * np.sin(2 * np.pi * t); y = 0.16 * np.sin(4 * np.pi * t); z = 0.55 + 0.06 * np.cos(2 * np.pi * t) + 0.03 * np.sin(6 * np.pi * t)
    elif kind == "coil":          # tight coil: many nodes on top of each other
        x = 0.10 * np.cos(6 * np.pi * t) + 0.15 * (t - 0.5); y = 0.10 * np.sin(6 * np.pi * t); z = 0.50 + 0.10 * t
    else:                         # "edge": a loop leaving the image on two sides
        x = 0.55 * np.sin(2 * np.pi * t * 1.1) * (0.5 + t); y = 0.40 * np.sin(4 * np.pi * t + 0.2); z = 0.60 + 0.05 * np.cos(2 * np.pi * t)
    Y = np.stack([x, y, z], 1) + rng.normal(0, 0.002, (n, 3))
    return Y


def mint(name, kind, n, rows, cols, width, seed, hidden=()):
    rng = np.random.default_rng(seed)
    Y = chain(kind, n, rng)
    P = synth.camera_matrix(rows, cols, f=915.0 * cols / 1280.0)
    # the observed cloud: points near the nodes, except `hidden` index ranges (nodes the depth camera lost)
    keep = np.ones(n, bool)
    for a, b in hidden: keep[a:b] = False
    X = np.concatenate([Y[i] + rng.normal(0, 0.002, (12, 3)) for i in range(n) if keep[i]])
    node_coord = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(Y, axis=0), axis=1))])
    free_cv2 = raster.self_occlusion(Y, P, rows, cols, width, use_raster=True, line_fn=cv2.line)
    free_or = raster.self_occlusion(Y, P, rows, cols, width)
    assert (free_cv2 == free_or).all(), name
    dmin = np.sqrt(((Y[:, None, :] - X[None, :, :]) ** 2).sum(-1)).min(1)
    thr, d_vis = 0.008, 0.06
    vis = [i for i in range(n) if free_cv2[i] and dmin[i] <= thr]
    ext = []
    for i in range(len(vis) - 1):
        ext.append(vis[i])
        if abs(node_coord[vis[i + 1]] - node_coord[vis[i]]) <= d_vis:
            ext.extend(range(vis[i] + 1, vis[i + 1]))
    ext.append(vis[-1])
    pix = np.array(raster.project_pixels(Y, P))
    np.savez_compressed(os.path.join(OUT, f"selfocc_{name}.npz"), Y=Y, X=X, proj=P, rows=rows, cols=cols, pixel_width=width, node_coord=node_coord,
                        cv2_not_self_occluded=free_cv2.astype(np.int32), visible=np.array(vis, np.int32), visible_ext=np.array(ext, np.int32),
                        pixels=pix, visibility_threshold=thr, d_vis=d_vis, cv2_version=cv2.__version__)
    inimg = ((pix[:, 0] >= 0) & (pix[:, 0] < cols) & (pix[:, 1] >= 0) & (pix[:, 1] < rows)).sum()
    print(f"{name}: n={n} {cols}x{rows} width={width}: self-occluded {int((~free_cv2).sum())}, visible {len(vis)}, extended {len(ext)}, nodes inside the image {inimg}")


if __name__ == "__main__":
    mint("loop", "loop", 45, 720, 1280, 40, 1)
    mint("eight", "eight", 50, 720, 1280, 40, 2, hidden=((20, 21), (33, 35)))
    mint("coil", "coil", 60, 480, 640, 25, 3)
    mint("edge", "edge", 40, 360, 480, 41, 4, hidden=((5, 8),))
    mint("thin", "eight", 30, 240, 320, 3, 5)
