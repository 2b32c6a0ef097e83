"""CPU simulation of pyramid-walk strategies for the FIRST NN pass of C3 (no warm start): how many pyramid nodes, leaf cells and
points per query does each visit?  Evidence for the round-2 plan, not part of the product or the tests.

  dfs      nearest-first branch-and-bound from the root (what k_nn_grid_walk does for a chain head)
  dfs+nb   the same with the bound taken from the previous query's answer (chain followers; spatially sorted source)
  bfs+nb   level-synchronous frontier walk (what k_nn_grid_walk_warp does) with the chain bound, leaves in rounds of 32
  bfs+pr   frontier walk with a bound from a greedy nearest-child probe first (cost of the probe included)
  bfs+mm   frontier walk, NO initial bound: every level tightens the bound with min over the frontier of the largest distance
           to a child's box (an occupied box holds a point, so the nearest neighbour is at most that far)
  bfs+mmnb the same starting from the chain bound

python tools/walk_sim.py [n_chains]
"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
from bench import WORKLOADS, make_inputs
from pcreg_b200 import synth

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 400
w = WORKLOADS['c3']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
model = np.asarray(model, dtype=np.float64)
lo, hi = model.min(0), model.max(0)
cell = float(np.cbrt(np.prod(hi - lo) / min(2 ** 27, 32 * model.shape[0])))
dims0 = np.floor((hi - lo) / cell).astype(np.int64) + 1
cells = np.minimum(np.floor((model - lo) / cell).astype(np.int64), dims0 - 1)
nlev = 1
while (1 << (nlev - 1)) < dims0.max():
    nlev += 1


def key(c):
    return (c[:, 2] << 40) | (c[:, 1] << 20) | c[:, 0]


def k3(x, y, z):
    return (z << 40) | (y << 20) | x


occ = [set(np.unique(key(cells >> l)).tolist()) for l in range(nlev)]
order = np.argsort(key(cells), kind='stable')
ks = key(cells)[order]
uk, start = np.unique(ks, return_index=True)
leaf_pts = {int(k): (model[order[s:e]] - lo) / cell for k, s, e in zip(uk, start, np.append(start[1:], len(ks)))}   # cell units
print('cell %.4f dims %s levels %d occupied leaves %d (%.2f points each)' % (cell, dims0, nlev, len(uk), model.shape[0] / len(uk)))


def box_lb2(q, c, edge):
    c = np.asarray(c, dtype=np.float64)
    d = np.maximum(np.maximum(c * edge - q, q - (c + 1.0) * edge), 0.0)
    return float(d @ d)


def children(l, c):
    x, y, z = c
    for k in range(8):
        cc = (2 * x + (k & 1), 2 * y + ((k >> 1) & 1), 2 * z + (k >> 2))
        if k3(*cc) in occ[l - 1]:
            yield cc


def scan_leaf(q, c, best, arg):
    p = leaf_pts[k3(*c)]
    d = ((p - q) ** 2).sum(1)
    i = int(d.argmin())
    if d[i] < best:
        best, arg = float(d[i]), p[i]
    return best, arg, len(p)


def dfs(q, best):
    nodes = leaves = pts = 0
    arg = None
    stack = [(0.0, nlev - 1, (0, 0, 0))]
    while stack:
        lb, l, c = stack.pop()
        if lb > best:
            continue
        nodes += 1
        if l == 0:
            best, arg, n = scan_leaf(q, c, best, arg)
            leaves += 1
            pts += n
            continue
        ch = [(box_lb2(q, cc, float(1 << (l - 1))), cc) for cc in children(l, c)]
        for lbc, cc in sorted(ch, key=lambda t: -t[0]):                      # far first -> the near child is popped first
            if lbc <= best:
                stack.append((lbc, l - 1, cc))
    return best, arg, nodes, leaves, pts


def probe(q):
    l, c, steps = nlev - 1, (0, 0, 0), 0
    while l > 0:
        c = min(children(l, c), key=lambda cc: box_lb2(q, cc, float(1 << (l - 1))))
        l -= 1
        steps += 1
    best, arg, n = scan_leaf(q, c, np.inf, None)
    return best, steps, n


def box_ub2(q, c, edge):
    c = np.asarray(c, dtype=np.float64)
    d = np.maximum(np.abs(c * edge - q), np.abs((c + 1.0) * edge - q))
    return float(d @ d)


def bfs(q, best, minmax=False):
    nodes = leaves = pts = 0
    arg = None
    fr = [(0.0, (0, 0, 0))]
    for l in range(nlev - 1, 0, -1):
        nodes += len(fr)
        nxt = []
        for lbp, c in fr:
            if lbp > best:
                continue
            for cc in children(l, c):
                edge = float(1 << (l - 1))
                lbc = box_lb2(q, cc, edge)
                if lbc <= best:
                    nxt.append((lbc, cc))
                    if minmax:
                        best = min(best, box_ub2(q, cc, edge) * (1 + 1e-9))
        fr = [t for t in nxt if t[0] <= best]
    nodes += len(fr)
    for r in range(0, len(fr), 32):                                          # rounds of 32 leaves, bound re-tightened between rounds
        b0 = best
        for lbc, c in fr[r:r + 32]:
            if lbc <= b0:
                best, arg, n = scan_leaf(q, c, best, arg)
                leaves += 1
                pts += n
    return best, arg, nodes, leaves, pts, len(fr)


g = np.random.default_rng(0)
sidx = np.lexsort((src[:, 0] // 1.0, src[:, 1] // 1.0, src[:, 2] // 1.0))      # coarse spatial order (the library Morton-sorts)
stats = {k: [] for k in ('dfs', 'dfs+nb', 'bfs+nb', 'bfs+pr', 'bfs+mm', 'bfs+mmnb')}
fmax = 0
for _ in range(n_chains):
    h = int(g.integers(0, T0.shape[0]))
    i0 = int(g.integers(0, src.shape[0] - 3))
    prev = None
    for j in range(3):                                                       # a chain of 3 consecutive queries (WALK_CHAIN)
        q = (synth.apply_T(src[sidx[i0 + j]][None], T0[h])[0] - lo) / cell
        if prev is None:
            best, arg, n, lf, p = dfs(q, np.inf)
            stats['dfs'].append((n, lf, p))
        else:
            bnd = float(((q - prev) ** 2).sum()) * (1 + 1e-9)
            best, arg, n, lf, p = dfs(q, bnd)
            stats['dfs+nb'].append((n, lf, p))
            b2, a2, n, lf, p, f = bfs(q, bnd)
            stats['bfs+nb'].append((n, lf, p))
            assert abs(b2 - best) <= 1e-9 * max(1.0, best)
            if arg is None:
                arg = prev
            b4, a4, n, lf, p, f = bfs(q, bnd, True)
            stats['bfs+mmnb'].append((n, lf, p))
            assert abs(b4 - best) <= 1e-9 * max(1.0, best)
        b5, a5, n, lf, p, f = bfs(q, np.inf, True)
        stats['bfs+mm'].append((n, lf, p))
        assert abs(b5 - best) <= 1e-9 * max(1.0, best)
        fmm = max(globals().get('fmm', 0), f)
        pb, st, pn = probe(q)
        b3, a3, n, lf, p, f = bfs(q, pb * (1 + 1e-9))
        fmax = max(fmax, f)
        stats['bfs+pr'].append((n + st, lf + 1, p + pn))
        assert min(b3, pb) <= best * (1 + 1e-9) + 1e-12
        prev = arg
print('r = sqrt(best) of the last query: %.1f cells; largest leaf frontier of bfs+pr: %d, of bfs+mm: %d' % (np.sqrt(best), fmax, fmm))
for k, v in stats.items():
    a = np.array(v, dtype=float)
    print('%-7s n=%5d  nodes/q %7.1f  leaves/q %6.1f  pts/q %7.1f' % (k, len(v), a[:, 0].mean(), a[:, 1].mean(), a[:, 2].mean()))
d = np.array(stats['dfs'], dtype=float)
f = np.array(stats['dfs+nb'], dtype=float)
print('chained per-lane walk (1 head + 2 followers): nodes/q %.1f leaves/q %.1f pts/q %.1f   [GPU counters, pass 0: 98.7 / 33.7 / 91.7]' % (
    (d[:, 0].mean() + 2 * f[:, 0].mean()) / 3, (d[:, 1].mean() + 2 * f[:, 1].mean()) / 3, (d[:, 2].mean() + 2 * f[:, 2].mean()) / 3))
