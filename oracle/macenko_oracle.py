"""ORACLE (test infrastructure, not product code) -- Macenko stain normalisation in fp64 NumPy.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.

PARITY UNPINNED: the reference snapshot (KatherLab/STAMP v2.5.0) contains NO Macenko code -- the
word occurs once, in README.md:35.  BASELINE.json's north_star nevertheless names it as the first
stage of the hot path, so this file restates the published algorithm (Macenko et al., "A method for
normalizing histology slides for quantitative analysis", ISBI 2009) in the widely used NumPy form
(the one STAMP v1 and torchstain ship), as fixed in SURVEY.md 8c:

    OD = -log((I + 1) / Io), Io = 240;   drop pixels with any OD < beta (0.15)
    eigh(cov(ODhat)) -> plane of the two largest eigenvectors; project; phi = atan2
    stain vectors at the alpha-th and (100 - alpha)-th percentile of phi (alpha = 1)
    H = the vector with the larger first component
    C = lstsq(HE, OD^T) over ALL pixels; maxC = percentile(C, 99, axis=1)
    C *= maxCRef / maxC;  I' = Io * exp(-HERef . C);  clip to [0, 255];  truncate to uint8

with HERef = [[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]], maxCRef = [1.9705, 1.0308].
The fit is pooled over ``tiles_per_fit`` consecutive tiles ("over the tile batch": one stain matrix
per group), then applied per pixel.  Two choices this restatement pins that the NumPy original
leaves to LAPACK / chance: eigenvector signs are canonicalised (largest: positive component sum;
second: positive first component -- the resulting stain vectors do not depend on the signs), and a
group with fewer than 16 tissue pixels is passed through unchanged.
"""

from __future__ import annotations

import numpy as np

HE_REF = np.array([[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]])
MAX_C_REF = np.array([1.9705, 1.0308])
MIN_TISSUE_PIXELS = 16


def fit(tiles_u8: np.ndarray, Io: float = 240.0, alpha: float = 1.0, beta: float = 0.15):
    """tiles_u8 [..., 3] uint8 -> (HE [3,2], maxC [2]) or None if too little tissue."""
    I = tiles_u8.reshape(-1, 3).astype(np.float64)
    OD = -np.log((I + 1.0) / Io)
    ODhat = OD[~np.any(OD < beta, axis=1)]
    if ODhat.shape[0] < MIN_TISSUE_PIXELS:
        return None
    _, eigvecs = np.linalg.eigh(np.cov(ODhat.T))
    E = eigvecs[:, 1:3].copy()          # columns: second largest, largest
    if E[:, 1].sum() < 0:
        E[:, 1] *= -1
    if E[0, 0] < 0:
        E[:, 0] *= -1
    That = ODhat @ E
    phi = np.arctan2(That[:, 1], That[:, 0])
    min_phi, max_phi = np.percentile(phi, alpha), np.percentile(phi, 100.0 - alpha)
    v_min = E @ np.array([np.cos(min_phi), np.sin(min_phi)])
    v_max = E @ np.array([np.cos(max_phi), np.sin(max_phi)])
    HE = np.stack([v_min, v_max], 1) if v_min[0] > v_max[0] else np.stack([v_max, v_min], 1)
    C = np.linalg.lstsq(HE, OD.T, rcond=None)[0]
    maxC = np.array([np.percentile(C[0], 99), np.percentile(C[1], 99)])
    return HE, maxC


def apply(tiles_u8: np.ndarray, HE: np.ndarray, maxC: np.ndarray, Io: float = 240.0) -> np.ndarray:
    shape = tiles_u8.shape
    I = tiles_u8.reshape(-1, 3).astype(np.float64)
    OD = -np.log((I + 1.0) / Io)
    C = np.linalg.lstsq(HE, OD.T, rcond=None)[0]
    C *= (MAX_C_REF / maxC)[:, None]
    out = Io * np.exp(-HE_REF @ C)
    return np.clip(out, 0, 255).T.reshape(shape).astype(np.uint8)


def normalize(tiles_u8: np.ndarray, tiles_per_fit: int | None = None, Io: float = 240.0,
              alpha: float = 1.0, beta: float = 0.15):
    """tiles_u8 [B,H,W,3] -> (normalised uint8 [B,H,W,3], HE [G,3,2], maxC [G,2], valid [G])."""
    B = tiles_u8.shape[0]
    tpf = tiles_per_fit or B
    G = (B + tpf - 1) // tpf
    out = np.empty_like(tiles_u8)
    HEs, maxCs, valid = np.zeros((G, 3, 2)), np.zeros((G, 2)), np.zeros(G, dtype=bool)
    for g in range(G):
        sl = slice(g * tpf, min(B, (g + 1) * tpf))
        r = fit(tiles_u8[sl], Io, alpha, beta)
        if r is None:
            out[sl] = tiles_u8[sl]
            continue
        HEs[g], maxCs[g], valid[g] = r[0], r[1], True
        out[sl] = apply(tiles_u8[sl], r[0], r[1], Io)
    return out, HEs, maxCs, valid
