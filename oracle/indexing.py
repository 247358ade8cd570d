"""Oracle (test infrastructure only): numpy restatement of the reference's
``misc/indexing.py`` -- PathIndex tables, edge->affinity, dense transition matrix and the
random-walk propagation.

Pinned by ``tests/golden/rw_*.npz`` / ``path_index.json`` (outputs of the unmodified
reference, see tests/golden/make_golden.py).
"""
import numpy as np


# --------------------------------------------------------------------------- R1
def half_plane_offsets(radius):
    """Destination offsets (dy, dx) in reference enumeration order.

    Follows misc/indexing.py:22-31: first (0, x) for x = 1..r-1, then for y = 1..r-1 every
    x in (-r, r) with x^2 + y^2 < r^2.
    """
    out = [(0, x) for x in range(1, radius)]
    for y in range(1, radius):
        for x in range(-radius + 1, radius):
            if x * x + y * y < radius * radius:
                out.append((y, x))
    return out


def path_points(dy, dx):
    """Grid points on the straight path from (0,0) to (dy,dx), destination first.

    misc/indexing.py:33-48: every integer point of the bounding box whose squared distance
    to the line, cross^2/len^2, is < 1 (an integer test: cross^2 < len^2), stably sorted by
    descending L1 norm (the box is scanned y-major, x ascending).
    """
    len_sq = dy * dy + dx * dx
    ylo, yhi = min(0, dy), max(0, dy)
    xlo, xhi = min(0, dx), max(0, dx)
    pts = []
    for y in range(ylo, yhi + 1):
        for x in range(xlo, xhi + 1):
            cross = dy * x - dx * y
            if cross * cross < len_sq:
                pts.append((y, x))
    pts.sort(key=lambda p: -(abs(p[0]) + abs(p[1])))   # python sort is stable
    return pts


def search_paths_dst(radius):
    """Paths grouped by length (ascending) + destination list (misc/indexing.py:18-56)."""
    by_len = {}
    for (dy, dx) in half_plane_offsets(radius):
        pts = path_points(dy, dx)
        by_len.setdefault(len(pts), []).append(pts)
    groups = [np.asarray(by_len[k], dtype=np.int64) for k in sorted(by_len)]
    dst = np.concatenate([g[:, 0] for g in groups], axis=0)
    return groups, dst


class PathIndex:
    """Same public attributes as the reference class (misc/indexing.py:6-16)."""

    def __init__(self, radius, default_size):
        self.radius = radius
        self.radius_floor = int(np.ceil(radius) - 1)
        self.search_paths, self.search_dst = search_paths_dst(radius)
        self.path_indices, self.src_indices, self.dst_indices = self._tables(default_size)

    def _tables(self, size):
        """misc/indexing.py:58-88: flat indices of every path point for every source pixel
        of the window rows [0, H - rf), cols [rf, W - rf)."""
        H, W = int(size[0]), int(size[1])
        rf = self.radius_floor
        ch, cw = H - rf, W - 2 * rf
        ys = np.arange(ch, dtype=np.int64)[:, None]
        xs = np.arange(cw, dtype=np.int64)[None, :] + rf
        src = (ys * W + xs).reshape(-1)
        tables = []
        for g in self.search_paths:                      # g: [n_paths, L, 2]
            off = g[:, :, 0] * W + g[:, :, 1]            # [n_paths, L]
            tables.append(src[None, None, :] + off[:, :, None])
        dst = np.concatenate([t[:, 0] for t in tables], axis=0)
        return tables, src, dst


# --------------------------------------------------------------------------- R3
def edge_to_affinity(edge_padded, path_indices):
    """aff[d, p] = 1 - max over the path of the padded edge map (misc/indexing.py:91-109).

    edge_padded: float32 [Hp, Wp]; returns float32 [n_dst, n_src]."""
    flat = np.asarray(edge_padded, dtype=np.float32).reshape(-1)
    out = []
    for t in path_indices:
        out.append(np.float32(1) - flat[t].max(axis=1))
    return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------- N4 (training side)
def to_affinity(edge, path_indices):
    """AffinityDisplacementLoss.to_affinity (net/resnet50_irn.py:162-175): edge float32 [B, H, W] (un-padded; PathIndex built
    for default_size (H, W)); per length group gather `edge.view(B,-1)` at the path indices and take 1 - max over the path.
    Returns (aff float32 [B, n_dst, n_src], arg int64 [B, n_dst, n_src]) where arg is the flat H*W index of the FIRST maximum
    along the path (np.argmax = max_pool2d's choice): what autograd differentiates through."""
    B = edge.shape[0]
    flat = np.asarray(edge, dtype=np.float32).reshape(B, -1)
    affs, args = [], []
    for t in path_indices:                     # t: int64 [n_paths, L, n_src]
        g = flat[:, t]                         # [B, n_paths, L, n_src]
        at = g.argmax(axis=2)                  # first maximum
        affs.append(np.float32(1) - np.take_along_axis(g, at[:, :, None, :], axis=2)[:, :, 0, :])
        args.append(np.take_along_axis(np.broadcast_to(t[None], g.shape), at[:, :, None, :], axis=2)[:, :, 0, :])
    return np.concatenate(affs, axis=1), np.concatenate(args, axis=1)


def to_affinity_backward(grad_aff, arg, hw):
    """Gradient of `to_affinity` w.r.t. the edge map: aff = 1 - edge[arg]  =>  grad_edge[arg] -= grad_aff (the scatter-add that
    index_select's backward performs, after max_pool2d's backward has routed each gradient to its maximum).  float64
    accumulation: the reference's index_add_ order is not defined, the exact sum is the fair target.  Returns float64 [B, hw]."""
    B = grad_aff.shape[0]
    out = np.zeros((B, hw), np.float64)
    for b in range(B):
        np.add.at(out[b], arg[b].reshape(-1), -grad_aff[b].reshape(-1).astype(np.float64))
    return out


# --------------------------------------------------------------------------- R4
def affinity_sparse2dense(aff, src, dst, n_vertices):
    """Symmetric dense matrix with unit diagonal (misc/indexing.py:112-129)."""
    A = np.zeros((n_vertices, n_vertices), np.float32)
    s = np.broadcast_to(src[None, :], dst.shape)
    np.add.at(A, (s.reshape(-1), dst.reshape(-1)), aff.reshape(-1))
    A[np.arange(n_vertices), np.arange(n_vertices)] += 1
    np.add.at(A, (dst.reshape(-1), s.reshape(-1)), aff.reshape(-1))
    return A


# --------------------------------------------------------------------------- R5
def to_transition_matrix(A, beta, times, flush_denormal=True):
    """T = A^beta / colsum, squared `times` times in fp32 (misc/indexing.py:132-139)."""
    import torch
    prev = torch.is_flush_denormal() if hasattr(torch, "is_flush_denormal") else False
    if flush_denormal:
        torch.set_flush_denormal(True)
    try:
        S = torch.pow(torch.from_numpy(A), beta)
        T = S / torch.sum(S, dim=0, keepdim=True)
        for _ in range(times):
            T = torch.matmul(T, T)
    finally:
        if flush_denormal:
            torch.set_flush_denormal(prev)
    return T.numpy()


# --------------------------------------------------------------------------- R6
def propagate_to_edge(x, edge, radius=5, beta=10, exp_times=8):
    """Faithful dense walk (misc/indexing.py:141-167).  x [..., h, w], edge [1, h, w] (or
    [h, w]) float32 -> [C, 1, h, w] float32.  O((hw)^2) memory: small grids only."""
    x = np.asarray(x, np.float32)
    edge = np.asarray(edge, np.float32).reshape(x.shape[-2:])
    h, w = x.shape[-2:]
    Hp, Wp = h + radius, w + 2 * radius
    pi = PathIndex(radius, (Hp, Wp))
    ep = np.ones((Hp, Wp), np.float32)
    ep[:h, radius:radius + w] = edge
    aff = edge_to_affinity(ep, pi.path_indices)
    A = affinity_sparse2dense(aff, pi.src_indices, pi.dst_indices, Hp * Wp)
    A = A.reshape(Hp, Wp, Hp, Wp)[:h, radius:radius + w, :h, radius:radius + w]
    A = np.ascontiguousarray(A).reshape(h * w, h * w)
    T = to_transition_matrix(A, beta, exp_times)
    xs = x.reshape(-1, h, w) * (np.float32(1) - edge)
    rw = xs.reshape(-1, h * w) @ T
    return rw.reshape(-1, 1, h, w).astype(np.float32)


def stencil_weights(edge, radius=5, beta=10):
    """Per-offset weights W[d, y, x] = a(p, p+d)^beta in float32 with a = 0 when p+d leaves
    the image (the reference pads the edge map with 1.0, misc/indexing.py:150), plus the
    offset list.  Equivalent to R1-R4 without the dense matrix."""
    edge = np.asarray(edge, np.float32)
    h, w = edge.shape[-2:]
    edge = edge.reshape(h, w)
    groups, dst = search_paths_dst(radius)
    r = radius
    ep = np.ones((h + 2 * r, w + 2 * r), np.float32)
    ep[r:r + h, r:r + w] = edge
    W = []
    for g in groups:
        for pts in g:
            m = None
            for (py, px) in pts:
                v = ep[r + py:r + py + h, r + px:r + px + w]
                m = v if m is None else np.maximum(m, v)
            a = np.float32(1) - m
            W.append(np.power(a, np.float32(beta)).astype(np.float32))
    return np.stack(W, 0), [tuple(int(v) for v in d) for d in dst]


def propagate_stencil(x, edge, radius=5, beta=10, n_iter=256, dtype=np.float64):
    """The same walk as `propagate_to_edge` written as n_iter applications of
    y_j <- (sum_i a_ij^beta y_i) / s_j; `dtype` is the state /
    accumulator precision.  With float64 this is the "truth" both the reference's fp32
    squaring and the CUDA kernel approximate."""
    x = np.asarray(x, np.float32)
    h, w = x.shape[-2:]
    e = np.asarray(edge, np.float32).reshape(h, w)
    W, offs = stencil_weights(e, radius, beta)
    Wd = W.astype(dtype)
    y = (x.reshape(-1, h, w) * (np.float32(1) - e)).astype(dtype)
    s = np.ones((h, w), dtype)
    r = radius
    for k, (dy, dx) in enumerate(offs):
        s += Wd[k]
        sh = np.zeros((h + 2 * r, w + 2 * r), dtype)
        sh[r + dy:r + dy + h, r + dx:r + dx + w] = Wd[k]
        s += sh[r:r + h, r:r + w]
    C = y.shape[0]
    buf = np.zeros((C, h + 2 * r, w + 2 * r), dtype)
    for _ in range(n_iter):
        buf[:, r:r + h, r:r + w] = y
        acc = y.copy()
        for k, (dy, dx) in enumerate(offs):
            # forward tap: neighbour p+d seen from p with weight W_d(p)
            acc += Wd[k] * buf[:, r + dy:r + dy + h, r + dx:r + dx + w]
        wy = np.zeros_like(buf)
        for k, (dy, dx) in enumerate(offs):
            # mirrored tap: neighbour p-d seen from p with weight W_d(p-d)
            wy[:, r:r + h, r:r + w] = Wd[k] * y
            acc += wy[:, r - dy:r - dy + h, r - dx:r - dx + w]
        y = acc / s
    return y.reshape(-1, 1, h, w)
