"""A SECOND, independent restatement of pcl::VoxelGrid<PointXYZRGB>::applyFilter, written in numpy from SURVEY.md Appendix A.8
alone (not from oracle/sdr_oracle.c and not from csrc/voxel.cu), to triangulate the one stage no importable reference pins.
PCL 1.14 lines each step restates (filters/include/pcl/filters/impl/voxel_grid.hpp, as recalled in A.8):
  getMinMax3D over finite points                      -> bounding box in fp32
  inverse_leaf_size_ = 1 / leaf_size_                 -> fp32 reciprocal per axis
  dx = int64((max - min) * inv) + 1, guard dx*dy*dz > INT32_MAX -> "Leaf size is too small", output = input
  min_b_ = floor(min * inv), max_b_, div_b_ = max_b_ - min_b_ + 1, divb_mul_ = (1, div_b.x, div_b.x * div_b.y)
  ijk = int(floor(p * inv) - float(min_b_)),  idx = ijk . divb_mul_
  sort by idx, one centroid per run of equal idx (common/impl/accumulators.hpp: xyz sums in fp32 / n, colour channels
  summed as float and divided by n, converted back to uint8 by truncation)
TEST INFRASTRUCTURE ONLY."""
import numpy as np

INT32_MAX = 2 ** 31 - 1


def voxelgrid_np(points, bgr, leaf, drop_mask=None):
    """points [..][3] f32, bgr [..][3] u8 or None, leaf scalar or 3-tuple, drop_mask [..] bool or None.
    Returns dict(idx, cnt, xyz, rgb, minb, divb), or None when the index guard trips."""
    f32 = np.float32
    pts = np.ascontiguousarray(points, f32).reshape(-1, 3)
    keep = np.isfinite(pts).all(axis=1)
    if drop_mask is not None:
        keep &= ~np.asarray(drop_mask, bool).ravel()
    src = np.nonzero(keep)[0]                                   # original point indices, ascending
    p = pts[src]
    leaf = np.array([leaf] * 3 if np.isscalar(leaf) else leaf, f32)
    inv = (f32(1.0) / leaf).astype(f32)
    if len(p) == 0:
        return dict(idx=np.empty(0, np.uint32), cnt=np.empty(0, np.uint32), xyz=np.empty((0, 3), f32), rgb=np.empty((0, 3), np.uint8),
                    minb=np.zeros(3, np.int32), divb=np.ones(3, np.int32))
    mn, mx = p.min(axis=0), p.max(axis=0)
    d = ((mx - mn).astype(f32) * inv).astype(f32).astype(np.int64) + 1        # C cast: truncation toward zero
    prod = int(d[0]) * int(d[1]) * int(d[2])
    prod = (prod + 2 ** 63) % 2 ** 64 - 2 ** 63                               # int64 wrap, as the C multiplication
    if prod > INT32_MAX:
        return None
    minb = np.floor((mn * inv).astype(f32)).astype(np.int32)
    maxb = np.floor((mx * inv).astype(f32)).astype(np.int32)
    divb = maxb - minb + 1
    # PCL does all of this in 32-bit ints; for absurd boxes (the reference's literal leaf 0.005 on a millimetre cloud) the
    # products wrap. Work in int64 and keep the low 32 bits: identical to two's-complement int32 arithmetic. The sort key is
    # PCL's `unsigned int idx`.
    mul = np.array([1, int(divb[0]), int(divb[0]) * int(divb[1])], np.int64)
    ijk = (np.floor((p * inv).astype(f32)) - minb.astype(f32)).astype(f32).astype(np.int32)
    idx = (ijk.astype(np.int64) * mul).sum(axis=1) & 0xFFFFFFFF
    order = np.lexsort((src, idx))                                            # by idx, ties in ascending point index
    sidx = idx[order]
    starts = np.nonzero(np.r_[True, sidx[1:] != sidx[:-1]])[0]
    cnt = np.diff(np.r_[starts, len(sidx)])
    # fp32 sums in point order: add the k-th member of every run at once (k = 0, 1, ...), each add rounded to fp32
    ps = p[order]
    if bgr is not None:
        cs = np.ascontiguousarray(bgr, np.uint8).reshape(-1, 3)[src][order][:, ::-1].astype(f32)    # r, g, b
    else:
        cs = np.zeros((len(ps), 3), f32)
    sx = np.zeros((len(starts), 3), f32)
    sc = np.zeros((len(starts), 3), f32)
    for k in range(int(cnt.max())):
        live = cnt > k
        pos = starts[live] + k
        sx[live] = (sx[live] + ps[pos]).astype(f32)
        sc[live] = (sc[live] + cs[pos]).astype(f32)
    n = cnt.astype(f32)[:, None]
    xyz = (sx / n).astype(f32)
    rgb = (sc / n).astype(f32).astype(np.uint32).astype(np.uint8)             # truncation
    return dict(idx=sidx[starts].astype(np.uint32), cnt=cnt.astype(np.uint32), xyz=xyz, rgb=rgb, minb=minb, divb=divb.astype(np.int32))


def adversarial_clouds():
    """[(name, points[N][3], bgr[N][3], leaf)]: the cases a second reader would get wrong first."""
    rng = np.random.default_rng(11)
    f32 = np.float32
    out = []
    # points exactly on voxel faces (multiples of the leaf, +-1 ulp), negative coordinates
    base = (rng.integers(-40, 40, (4000, 3)) * 5).astype(f32)
    jitter = rng.integers(-1, 2, base.shape)
    on_faces = np.nextafter(base, base + jitter.astype(f32) * f32(10)).astype(f32)
    out.append(("faces", on_faces, rng.integers(0, 256, (4000, 3), dtype=np.uint8), 5.0))
    # every point alone in its voxel
    g = np.stack(np.meshgrid(np.arange(12), np.arange(9), np.arange(7), indexing="ij"), -1).reshape(-1, 3).astype(f32) * f32(7.5) - f32(30)
    out.append(("singletons", g, rng.integers(0, 256, (len(g), 3), dtype=np.uint8), 5.0))
    # many points per voxel (long fp32 sums), anisotropic leaf, non-finite entries
    dense = (rng.random((30000, 3)) * np.array([60, 40, 25]) - np.array([30, 20, 0])).astype(f32)
    dense[::97, 1] = np.nan
    dense[5::131, 2] = np.inf
    out.append(("dense_nonfinite", dense, rng.integers(0, 256, (30000, 3), dtype=np.uint8), (4.0, 2.5, 6.0)))
    # colours whose mean truncates (e.g. 255, 254 -> 254.5 -> 254)
    two = np.repeat((rng.random((500, 3)) * 100).astype(f32), 2, axis=0)
    col = np.tile(np.array([[255, 0, 1], [254, 1, 2]], np.uint8), (500, 1))
    out.append(("colour_truncation", two, col, 5.0))
    # the int32 guard: a bounding box of 3000^3 voxels
    far = np.array([[0, 0, 0], [15000, 15000, 15000], [3, 3, 3]], f32)
    out.append(("guard", far, np.zeros((3, 3), np.uint8), 5.0))
    # a single point; all points identical
    out.append(("one_point", np.array([[1.25, -2.5, 900.0]], f32), np.array([[9, 8, 7]], np.uint8), 5.0))
    out.append(("identical", np.tile(np.array([[-7.5, 2.5, 33.0]], f32), (50, 1)), rng.integers(0, 256, (50, 3), dtype=np.uint8), 5.0))
    return out
