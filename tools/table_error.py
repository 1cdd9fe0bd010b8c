"""Off-node accuracy of a Kerr transfer table set (VERDICT r1 item 1c / ADVICE high).

For a requested (a, theta) a table set gives a blended slice g(r_s, phi_t), Jn(r_s, phi_t); the yardstick is a node
ray-traced AT (a, theta) (same rows and columns, so slices compare row by row).  Per table row the two numbers that
decide a spectrum are
    flux      F_s = sum_t (w_t + w_{t+1})/2 dphi,           w = Jn g^3          (weight of the ring)
    centroid  c_s = sum_t W_t (s_t + s_{t+1})/2 / F_s,      s = ln g / dlnE     (mean shift in energy bins)
(the hat deposit of the transfer kernel preserves both exactly).  Reported: flux ratio - 1 and centroid shift in bins
of the 1000-bin metric grid, worst row with r >= 1.5 r_isco and worst row overall.

    python tools/table_error.py [dense_tables.npz] [truth_nodes.npz]

The blend schemes live here as numpy restatements; the production scheme is the one rg_transfer.cuh and
oracle/kerr_oracle.c implement.  Test infrastructure / design tool only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DLNE = np.log(1e7) / 1000.0


def risco(a):
    a = np.asarray(a, dtype=np.float64)
    z1 = 1 + (1 - a * a) ** (1 / 3) * ((1 + a) ** (1 / 3) + (1 - a) ** (1 / 3))
    z2 = np.sqrt(3 * a * a + z1 * z1)
    return 3 + z2 - np.sign(a) * np.sqrt((3 - z1) * (3 + z1 + 2 * z2))


def row_radii(a, NR, dl):
    rh = 1 + np.sqrt(1 - a * a)
    return rh + (risco(a) - rh) * 10.0 ** (np.arange(NR) * dl)


def row_metrics(g, jn, dlnE=DLNE):
    """per-row flux and centroid (bins) of a slice [NR][NPHI]"""
    g = g.astype(np.float64)
    jn = jn.astype(np.float64)
    w = jn * g ** 3
    s = np.log(g) / dlnE
    w1, s1 = np.roll(w, -1, axis=1), np.roll(s, -1, axis=1)
    W = 0.5 * (w + w1)
    F = W.sum(axis=1) * (2 * np.pi / g.shape[1])
    c = (W * 0.5 * (s + s1)).sum(axis=1) / W.sum(axis=1)
    return F, c


def lagrange_weights(x, xs):
    """Lagrange weights of the nodes xs (any count) at x"""
    w = np.ones(len(xs))
    for i in range(len(xs)):
        for k in range(len(xs)):
            if k != i:
                w[i] *= (x - xs[k]) / (xs[i] - xs[k])
    return w


def stencil(x, nodes, order):
    """indices and Lagrange weights of the `order` nodes around x (order 2 = linear, 4 = cubic), clamped at the ends"""
    n = len(nodes)
    i = int(np.searchsorted(nodes, x, side="right")) - 1
    i = min(max(i, 0), n - 2)
    if order == 2 or n < 4:
        idx = [i, i + 1]
    else:
        i0 = min(max(i - 1, 0), n - 4)
        idx = [i0, i0 + 1, i0 + 2, i0 + 3]
    xs = np.array([nodes[k] for k in idx])
    return idx, lagrange_weights(x, xs)


def blend(tab, a, theta, order_a=2, order_i=2, xa="risco", xi="theta", fg=None, fj=None):
    """blended slice of table set `tab` at (a, theta).
    xa: spin coordinate ('risco' | 'a'), xi: inclination coordinate ('theta' | 'mu' | 'sin').
    fg / fj: optional (forward, inverse) transforms applied to g / Jn before / after blending, each called as
    f(values, a_node, theta_node)."""
    spins, thetas = tab["spins"], tab["thetas"]
    ca = risco(spins)[::-1] if xa == "risco" else spins
    # risco decreases with spin: work on an ascending coordinate
    if xa == "risco":
        idx_a, w_a = stencil(float(risco(a)), ca, order_a)
        idx_a = [len(spins) - 1 - k for k in idx_a]
    else:
        idx_a, w_a = stencil(a, ca, order_a)
    if xi == "theta":
        ci, x = thetas, theta
    elif xi == "mu":
        ci, x = -np.cos(thetas), -np.cos(theta)
    else:
        ci, x = np.sin(thetas), np.sin(theta)
    idx_i, w_i = stencil(x, ci, order_i)
    data = tab["data"]
    G = 0.0
    J = 0.0
    for m, wa in zip(idx_a, w_a):
        for n, wi in zip(idx_i, w_i):
            g = data[m, n, 0].astype(np.float64)
            j = data[m, n, 1].astype(np.float64)
            if fg:
                g = fg[0](g, spins[m], thetas[n])
            if fj:
                j = fj[0](j, spins[m], thetas[n])
            G = G + wa * wi * g
            J = J + wa * wi * j
    if fg:
        G = fg[1](G, a, theta)
    if fj:
        J = fj[1](J, a, theta)
    return G, J


def compare(tab, truth_a, truth_th, truth_data, **kw):
    """rows: per truth node (a, theta_deg, worst |flux-1| and |dc| for r >= 1.5 risco, and overall)"""
    out = []
    NR, dl = int(tab["NR"]), float(tab["dl"])
    for k in range(len(truth_a)):
        a, th = float(truth_a[k]), float(truth_th[k])
        G, J = blend(tab, a, th, **kw)
        F0, c0 = row_metrics(truth_data[k, 0], truth_data[k, 1])
        F1, c1 = row_metrics(G, J)
        rr = row_radii(a, NR, dl)
        ok = F0 > 0
        fe = np.where(ok, F1 / np.where(ok, F0, 1) - 1, 0)
        ce = np.where(ok, c1 - c0, 0)
        far = rr >= 1.5 * risco(a)
        out.append((a, np.rad2deg(th), np.abs(fe[far]).max(), np.abs(ce[far]).max(), np.abs(fe).max(), np.abs(ce).max()))
    return np.array(out)


def subsample(tab, sa, si):
    return dict(spins=tab["spins"][sa], thetas=tab["thetas"][si], NR=tab["NR"], NPHI=tab["NPHI"], dl=tab["dl"],
                data=tab["data"][np.ix_(sa, si)])


def summarize(name, res):
    print("%-44s far: flux %.2e  dc %.3f | all: flux %.2e  dc %.3f   (worst far flux at a=%.3f i=%.1f)" % (
        name, res[:, 2].max(), res[:, 3].max(), res[:, 4].max(), res[:, 5].max(),
        res[np.argmax(res[:, 2]), 0], res[np.argmax(res[:, 2]), 1]))


if __name__ == "__main__":
    dense = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "dense_tables.npz"))
    truth = np.load(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "truth_nodes.npz"))
    tab = {k: dense[k] for k in dense.files}
    for name, kw in [("dense bilinear", dict()), ("dense bicubic", dict(order_a=4, order_i=4))]:
        summarize(name, compare(tab, truth["a"], truth["theta"], truth["data"], **kw))
