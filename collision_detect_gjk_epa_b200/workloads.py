"""Synthetic convex-hull pair workloads.

Everything is unit length scale because every tolerance in the reference is absolute
(GCLIB_GJKEPA.f90:106,123,140,157,199,203,1172,...).  Hull store layout (the one the C ABI takes):
hull h is the Fortran ``p(n,3)`` column-major block ``[x(n) | y(n) | z(n)]`` starting at double
index ``3*hull_off[h]``; ``hull_off`` has ``n_hulls+1`` entries in vertex units.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PairBatch:
    verts: np.ndarray      # float64, packed hull blocks
    hull_off: np.ndarray   # int64, n_hulls+1 (vertex units)
    pair_a: np.ndarray     # int32, n_pairs
    pair_b: np.ndarray     # int32, n_pairs
    name: str = ""

    @property
    def n_pairs(self) -> int:
        return int(self.pair_a.shape[0])

    @property
    def n_hulls(self) -> int:
        return int(self.hull_off.shape[0] - 1)

    def hull(self, h: int) -> np.ndarray:
        """(n,3) array of hull h's vertices."""
        o0, o1 = int(self.hull_off[h]), int(self.hull_off[h + 1])
        n = o1 - o0
        return self.verts[3 * o0:3 * o1].reshape(3, n).T.copy()

    def algorithmic_bytes_per_pair(self, out_bytes: int) -> float:
        """24*(nA+nB) vertex bytes + 8 index bytes + outputs."""
        n = np.diff(self.hull_off)
        return float(np.mean(24.0 * (n[self.pair_a] + n[self.pair_b])) + 8 + out_bytes)


def _quat_rot(rng: np.random.Generator, m: int) -> np.ndarray:
    q = rng.standard_normal((m, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    r = np.empty((m, 3, 3))
    r[:, 0, 0] = 1 - 2 * (y * y + z * z); r[:, 0, 1] = 2 * (x * y - z * w); r[:, 0, 2] = 2 * (x * z + y * w)
    r[:, 1, 0] = 2 * (x * y + z * w); r[:, 1, 1] = 1 - 2 * (x * x + z * z); r[:, 1, 2] = 2 * (y * z - x * w)
    r[:, 2, 0] = 2 * (x * z - y * w); r[:, 2, 1] = 2 * (y * z + x * w); r[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return r


def _box8_local(rng, m):
    he = rng.uniform(0.25, 1.0, size=(m, 3))
    # vertex order: x sign fastest, then y, then z
    sx = np.array([-1, 1, -1, 1, -1, 1, -1, 1.0])
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    s = np.stack([sx, sy, sz], axis=1)             # (8,3)
    return s[None, :, :] * he[:, None, :]          # (m,8,3)


def _hull_local(rng, m, n):
    d = rng.standard_normal((m, n, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    ax = rng.uniform(0.5, 1.0, size=(m, 1, 3))
    return d * ax                                   # ellipsoid surface: every point is extreme


def _pack_uniform(pts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """pts (H,n,3) -> packed blocks, hull_off."""
    h, n, _ = pts.shape
    verts = np.ascontiguousarray(pts.transpose(0, 2, 1)).reshape(-1)     # per hull [x|y|z]
    hull_off = np.arange(h + 1, dtype=np.int64) * n
    return verts, hull_off


def make_pairs(kind: str, n_pairs: int, sigma: float, seed: int = 12345, n_vert: int = 8) -> PairBatch:
    """Per-pair hulls: pair i = (hull 2i, hull 2i+1); body B translated by N(0, sigma^2)^3."""
    rng = np.random.default_rng(seed)
    m = 2 * n_pairs
    if kind == "box8":
        loc = _box8_local(rng, m)
    elif kind == "hull":
        loc = _hull_local(rng, m, n_vert)
    else:
        raise ValueError(kind)
    rot = _quat_rot(rng, m)
    pts = np.einsum("mij,mvj->mvi", rot, loc)
    off = rng.standard_normal((n_pairs, 3)) * sigma
    pts[1::2] += off[:, None, :]
    verts, hull_off = _pack_uniform(pts)
    pa = np.arange(0, m, 2, dtype=np.int32)
    pb = pa + 1
    return PairBatch(verts, hull_off, pa, pb, name=f"{kind}{pts.shape[1]}_s{sigma}")


def make_mixed(n_pairs: int, sigma: float, seed: int = 12345, n_lo: int = 4, n_hi: int = 64) -> PairBatch:
    """cfg4: hull(n) with n ~ U{n_lo..n_hi} per body."""
    rng = np.random.default_rng(seed)
    m = 2 * n_pairs
    n = rng.integers(n_lo, n_hi + 1, size=m)
    hull_off = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(n, out=hull_off[1:])
    total = int(hull_off[-1])
    owner = np.repeat(np.arange(m), n)
    d = rng.standard_normal((total, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ax = rng.uniform(0.5, 1.0, size=(m, 3))
    loc = d * ax[owner]
    rot = _quat_rot(rng, m)
    pts = np.einsum("vij,vj->vi", rot[owner], loc)
    off = np.zeros((m, 3))
    off[1::2] = rng.standard_normal((n_pairs, 3)) * sigma
    pts += off[owner]
    # pack per hull as [x(n)|y(n)|z(n)]
    verts = np.empty(3 * total)
    local = np.arange(total) - hull_off[owner]
    base = 3 * hull_off[owner]
    verts[base + local] = pts[:, 0]
    verts[base + n[owner] + local] = pts[:, 1]
    verts[base + 2 * n[owner] + local] = pts[:, 2]
    pa = np.arange(0, m, 2, dtype=np.int32)
    return PairBatch(verts, hull_off, pa, pa + 1, name=f"mixed{n_lo}-{n_hi}_s{sigma}")


def make_shared_store(n_bodies: int, n_pairs: int, n_vert: int = 16, seed: int = 12345,
                      spacing: float = 1.0, chunk: int = 1 << 20) -> PairBatch:
    """cfg5: a shared store of n_bodies hull(n_vert) bodies on a jittered cubic lattice (pitch
    `spacing`, body i at lattice site (i % g, i//g % g, i//g^2)) and a broad-phase-like pair list:
    each pair joins a random body with one of its 13 forward lattice neighbours, so x-neighbours are
    adjacent in the store and y/z-neighbours are far apart in memory.  Built in chunks to bound host
    memory at 10M bodies."""
    rng = np.random.default_rng(seed)
    g = int(np.ceil(n_bodies ** (1.0 / 3.0)))
    verts = np.empty(3 * n_vert * n_bodies)
    vview = verts.reshape(n_bodies, 3, n_vert)
    for lo in range(0, n_bodies, chunk):
        hi = min(n_bodies, lo + chunk)
        m = hi - lo
        loc = _hull_local(rng, m, n_vert)
        rot = _quat_rot(rng, m)
        pts = np.einsum("mij,mvj->mvi", rot, loc)
        idx = np.arange(lo, hi)
        site = np.stack([idx % g, (idx // g) % g, idx // (g * g)], axis=1).astype(np.float64)
        pos = (site + rng.uniform(-0.3, 0.3, size=(m, 3))) * spacing
        pts += pos[:, None, :]
        vview[lo:hi] = pts.transpose(0, 2, 1)
    hull_off = np.arange(n_bodies + 1, dtype=np.int64) * n_vert
    fwd = np.array([(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
                    if (dz, dy, dx) > (0, 0, 0)], dtype=np.int64)           # 13 forward neighbours
    stride = fwd[:, 0] + g * fwd[:, 1] + g * g * fwd[:, 2]
    a = rng.integers(0, n_bodies, size=n_pairs, dtype=np.int64)
    b = (a + stride[rng.integers(0, 13, size=n_pairs)]) % n_bodies
    return PairBatch(verts, hull_off, a.astype(np.int32), b.astype(np.int32), name=f"store{n_bodies}x{n_vert}")


def make_shared_store_gpu(device, n_bodies: int, n_pairs: int, n_vert: int = 16, seed: int = 12345,
                          spacing: float = 1.0, chunk: int = 1 << 20):
    """make_shared_store's recipe generated with torch on `device` (numpy takes minutes for cfg5's 10M bodies / 100M
    pairs) and delivered in PINNED host tensors (verts float64, hull_off int64, pair_a / pair_b int32), as a host that
    keeps its geometry in page-locked memory would hold them.  Same distribution, different random stream."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    g = int(np.ceil(n_bodies ** (1.0 / 3.0)))
    verts = torch.empty(3 * n_vert * n_bodies, dtype=torch.float64, pin_memory=True)
    vview = verts.view(n_bodies, 3, n_vert)
    f64 = dict(dtype=torch.float64, device=device, generator=gen)
    for lo in range(0, n_bodies, chunk):
        hi = min(n_bodies, lo + chunk)
        m = hi - lo
        d = torch.randn((m, n_vert, 3), **f64)
        d /= d.norm(dim=2, keepdim=True)
        loc = d * (0.5 + 0.5 * torch.rand((m, 1, 3), **f64))
        q = torch.randn((m, 4), **f64)
        q /= q.norm(dim=1, keepdim=True)
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        rot = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                           2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).view(m, 3, 3)
        pts = torch.einsum("mij,mvj->mvi", rot, loc)
        idx = torch.arange(lo, hi, device=device)
        site = torch.stack([idx % g, (idx // g) % g, idx // (g * g)], dim=1).to(torch.float64)
        pts += ((site + (torch.rand((m, 3), **f64) * 0.6 - 0.3)) * spacing)[:, None, :]
        vview[lo:hi].copy_(pts.transpose(1, 2))
    hull_off = (torch.arange(n_bodies + 1, dtype=torch.int64) * n_vert).pin_memory()
    fwd = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dz, dy, dx) > (0, 0, 0)]
    stride = torch.tensor([dx + g * dy + g * g * dz for dx, dy, dz in fwd], dtype=torch.int64, device=device)
    pair_a = torch.empty(n_pairs, dtype=torch.int32, pin_memory=True)
    pair_b = torch.empty(n_pairs, dtype=torch.int32, pin_memory=True)
    for lo in range(0, n_pairs, 16 * chunk):
        hi = min(n_pairs, lo + 16 * chunk)
        a = torch.randint(0, n_bodies, (hi - lo,), device=device, generator=gen, dtype=torch.int64)
        b = (a + stride[torch.randint(0, 13, (hi - lo,), device=device, generator=gen)]) % n_bodies
        pair_a[lo:hi].copy_(a.to(torch.int32))
        pair_b[lo:hi].copy_(b.to(torch.int32))
    torch.cuda.synchronize(device)
    return verts, hull_off, pair_a, pair_b


def unit_cubes(offset=(0.5, 0.0, 0.0)) -> PairBatch:
    """cfg1: A = [-.5,.5]^3, B = A + offset."""
    sx = np.array([-1, 1, -1, 1, -1, 1, -1, 1.0])
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    a = 0.5 * np.stack([sx, sy, sz], axis=1)
    b = a + np.asarray(offset, dtype=np.float64)[None, :]
    verts, hull_off = _pack_uniform(np.stack([a, b]))
    return PairBatch(verts, hull_off, np.array([0], np.int32), np.array([1], np.int32), name="cubes")


@dataclass
class PosedBatch:
    """Body-space shapes + per-body poses + body pairs (input of gjkepa_batch_posed_f64)."""
    shape_verts: np.ndarray    # float64, packed shape blocks [x(n) | y(n) | z(n)]
    shape_off: np.ndarray      # int64, n_shapes+1
    body_shape: np.ndarray     # int32, n_bodies
    pose: np.ndarray           # float64 (n_bodies, 12): R row-major, t
    pair_a: np.ndarray         # int32 body numbers
    pair_b: np.ndarray

    @property
    def n_pairs(self) -> int:
        return int(self.pair_a.shape[0])


def make_posed(n_shapes: int, n_bodies: int, n_pairs: int, sigma: float = 0.85, seed: int = 4242,
               n_lo: int = 8, n_hi: int = 8, kind: str = "box") -> PosedBatch:
    """Bodies instance a library of shapes (boxes, or hull(n) with n ~ U{n_lo..n_hi}) under random rigid poses;
    pairs join bodies whose centres are N(0, sigma^2) apart (body b+1 is placed relative to body b for b even)."""
    rng = np.random.default_rng(seed)
    sizes = np.full(n_shapes, 8) if kind == "box" else rng.integers(n_lo, n_hi + 1, size=n_shapes)
    shape_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    shape_verts = np.empty(3 * int(shape_off[-1]))
    for n in np.unique(sizes):
        idx = np.nonzero(sizes == n)[0]
        loc = _box8_local(rng, len(idx)) if kind == "box" else _hull_local(rng, len(idx), int(n))
        blk = np.ascontiguousarray(loc.transpose(0, 2, 1)).reshape(len(idx), 3 * n)
        shape_verts[(3 * shape_off[idx][:, None] + np.arange(3 * n)[None, :])] = blk
    body_shape = rng.integers(0, n_shapes, size=n_bodies).astype(np.int32)
    rot = _quat_rot(rng, n_bodies).reshape(n_bodies, 9)
    t = rng.uniform(-50.0, 50.0, size=(n_bodies, 3))
    t[1::2] = t[0::2][: len(t[1::2])] + rng.standard_normal((len(t[1::2]), 3)) * sigma
    pose = np.concatenate([rot, t], axis=1)
    pa = (2 * rng.integers(0, n_bodies // 2, size=n_pairs)).astype(np.int32)
    pb = pa + 1
    return PosedBatch(shape_verts, shape_off, body_shape, pose, pa, pb.astype(np.int32))


def make_touching_mix(n_pairs: int, seed: int = 77, every: int = 5, n_vert: int = 20) -> PairBatch:
    """Random hull(n_vert) pairs with, every `every`-th pair, two axis-aligned unit cubes carrying their 12
    edge midpoints as extra vertices (20 vertices, same hull) that touch, nearly touch or overlap along x by
    an exact amount.  These pairs reach the reference's 1e-8 bands (on-face containment F:1248, coplanar exits
    F:157/F:203) inside warps whose other pairs do not -- the case a lockstep kernel must survive."""
    assert n_vert == 20
    b = make_pairs("hull", n_pairs, 0.85, seed=seed, n_vert=n_vert)
    rng = np.random.default_rng(seed + 1)
    c = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], float)
    mids = np.array([(c[i] + c[j]) / 2 for i in range(8) for j in range(i + 1, 8) if np.sum(c[i] != c[j]) == 1])
    cube = 0.5 * np.concatenate([c, mids])                     # (20,3)
    gaps = np.array([1.0, 1.0 - 5e-9, 1.0 + 5e-9, 1.0 - 1e-6, 0.75, 1.0 + 1e-6, 0.5, 1.0, 0.999999995])
    v = b.verts.reshape(-1, 3, n_vert)
    for i in range(0, n_pairs, every):
        shift = rng.integers(-2, 3, size=3).astype(float)       # exact small integers: the cubes stay axis-aligned
        perm = rng.permutation(n_vert)
        off = np.array([gaps[(i // every) % len(gaps)], 0.25 * ((i // every) % 3), 0.125 * ((i // every) % 2)])
        v[2 * i] = (cube[perm] + shift).T
        v[2 * i + 1] = (cube[rng.permutation(n_vert)] + shift + off).T
    return PairBatch(v.reshape(-1), b.hull_off, b.pair_a, b.pair_b, name="touching_mix")


CONFIGS = {
    # name: (factory, kwargs, mode)   -- BASELINE.json configs[1..4]
    "cfg2": (make_pairs, dict(kind="box8", sigma=1.0), "gjk"),
    "cfg3": (make_pairs, dict(kind="hull", n_vert=32, sigma=0.85), "epa"),
    "cfg4": (make_mixed, dict(sigma=0.25), "epa"),
    "cfg4deep": (make_pairs, dict(kind="hull", n_vert=64, sigma=0.05), "epa"),
}


def make_nonfinite(kind: str, n_pairs: int = 20000, n_vert: int = 8, seed: int = 11) -> PairBatch:
    """Random hull pairs with a tenth of the bodies corrupted: one coordinate NaN ("nan") or +-Inf ("inf"), every
    coordinate Inf ("infall"), scaled by 1e300 / 1e150 / 1e-160 ("huge2" / "huge" / "tiny"), or all zero ("zero");
    "tiny2" scales everything by 1e-300 and "far" shifts everything by 1e12.  Garbage in -- the reference would print
    and STOP on some of it -- but the flag path is still fully defined arithmetic, so the kernels must follow the oracle."""
    b = make_pairs("hull", n_pairs, 0.7, seed=seed, n_vert=n_vert)
    rng = np.random.default_rng(seed + 1000)
    v = b.verts.reshape(-1, 3, n_vert)
    m = len(v)
    idx = rng.integers(0, m, size=m // 10)
    ax, vx = rng.integers(0, 3, len(idx)), rng.integers(0, n_vert, len(idx))
    if kind == "nan":
        v[idx, ax, vx] = np.nan
    elif kind == "inf":
        v[idx, ax, vx] = np.inf * rng.choice([-1.0, 1.0], len(idx))
    elif kind == "infall":
        v[idx] = np.inf
    elif kind == "huge2":
        v[idx] *= 1e300
    elif kind == "huge":
        v[idx] *= 1e150
    elif kind == "tiny":
        v[idx] *= 1e-160
    elif kind == "tiny2":
        v *= 1e-300
    elif kind == "far":
        v += 1e12
    elif kind == "zero":
        v[idx] = 0.0
    else:
        raise ValueError(kind)
    b.name = f"nonfinite_{kind}"
    return b
