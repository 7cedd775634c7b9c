"""Block-coordinate-descent oracle (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates `python bcd.py`:
  sidepsi :84-88, purepsi :98-99, bcd :101-257 (one Viterbi chain), ceoBCD :261-284 (phase order).

Recurrence exactly as the reference evaluates it (float64, same operation order):
  side neighbours of a chain pixel are the two pixels ALONG the chain axis, (ty+-yside, tx+-xside)
  with (yside,xside) = (0,1) for row chains and (1,0) for column chains (:108-113), read with the
  labels the chain had before this call (bestlabels is only rewritten by the backtrack, :238-253);
  dp0(l)  = (psi_plus + psi_minus) + lamda*lcost                                   (:118-120)
  U_i(l)  = (lamda*lcost + psi_plus) + psi_minus                                   (:161-162 / :196-197)
  S_l     = {k < nprop[prev] : L1(v_l, u_k) < tpsi}     (what packedksets caches, :131-142)
  S_l non-empty: m = min_{k in S_l}(dp_{i-1}(k) + L1), back = lowest such k       (:170-175 / :213-218)
  S_l empty    : m = min_k(tpsi + dp_{i-1}(k)),        back = lowest such k       (:152-157)   [quirk Q1]
  dp_i(l) = m + U_i(l)                                                              (:176 / :219)
  end: lowest-index argmin of dp_last (:231-237), backtrack (:238-253).
Chains of one phase touch disjoint pixels, so they are evaluated together (vectorised over chains);
the four phases and the sweeps are sequential as in ceoBCD.
"""
import numpy as np


def _psi_side(v, nb_vec, valid, tpsi):
    """min(tpsi, L1(v, nb_vec)) where the neighbour exists, else 0 (sidepsi :84-88).  int64."""
    l1 = np.abs(v[..., 0] - nb_vec[..., None, 0]) + np.abs(v[..., 1] - nb_vec[..., None, 1])
    return np.where(valid[..., None], np.minimum(tpsi, l1), 0)


def _phase(proposals, lcosts, nprop, labels, tpsi, lamda, ystep, xstep, starts):
    """Run all chains of one phase.  starts: list of (ty, tx).  labels modified in place."""
    H, W, K, _ = proposals.shape
    C = len(starts)
    if C == 0:
        return
    sy = np.array([s[0] for s in starts])
    sx = np.array([s[1] for s in starts])
    L = H if ystep != 0 else W
    yside, xside = (0, 1) if ystep == 0 else (1, 0)
    old = labels.copy()                                   # labels before this phase (chains disjoint)
    lab_idx = np.arange(K)
    dp_prev = None
    back = np.zeros((L, C, K), dtype=np.int64)
    for i in range(L):
        y = sy + i * ystep
        x = sx + i * xstep
        v = proposals[y, x]                               # (C,K,2)
        n = nprop[y, x]                                   # (C,)
        cost = lcosts[y, x]                               # (C,K)

        def side(dy, dx):
            yy, xx = y + dy, x + dx
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            yc, xc = np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)
            nb = proposals[yc, xc, old[yc, xc]]           # (C,2)
            return _psi_side(v, nb, ok, tpsi)
        psi_p = side(yside, xside)
        psi_m = side(-yside, -xside)
        if i == 0:
            dp = (psi_p + psi_m) + lamda * cost
        else:
            small = (lamda * cost + psi_p) + psi_m
            u = proposals[y - ystep, x - xstep]           # (C,K,2) previous pixel's vectors
            pn = nprop[y - ystep, x - xstep]
            l1 = np.abs(v[:, :, None, 0] - u[:, None, :, 0]) + np.abs(v[:, :, None, 1] - u[:, None, :, 1])
            kvalid = lab_idx[None, None, :] < pn[:, None, None]
            near = (l1 < tpsi) & kvalid
            cand = np.where(near, dp_prev[:, None, :] + l1, np.inf)
            m_near = cand.min(axis=2)
            a_near = cand.argmin(axis=2)
            trunc = np.where(lab_idx[None, :] < pn[:, None], tpsi + dp_prev, np.inf)
            a_tr = trunc.argmin(axis=1)
            m_tr = trunc[np.arange(C), a_tr]
            has = near.any(axis=2)
            m = np.where(has, m_near, m_tr[:, None])
            back[i] = np.where(has, a_near, a_tr[:, None])
            dp = m + small
        dp_prev = np.where(lab_idx[None, :] < n[:, None], dp, np.inf)
    # final label and backtrack
    lab = dp_prev.argmin(axis=1)
    for i in range(L - 1, -1, -1):
        y = sy + i * ystep
        x = sx + i * xstep
        labels[y, x] = lab
        if i > 0:
            lab = back[i, np.arange(C), lab]


def sweep(proposals, lcosts, nprop, labels, tpsi=8, lamda=0.05, chain_chunk=64):
    """One pass of ceoBCD's loop body (:265-277): phases A, B, C, D.  labels (int64) in place."""
    H, W = labels.shape

    def run(ystep, xstep, starts):
        for s in range(0, len(starts), chain_chunk):
            _phase(proposals, lcosts, nprop, labels, tpsi, lamda, ystep, xstep, starts[s:s + chain_chunk])
    run(1, 0, [(0, x) for x in range(0, W, 2)])                       # A: even columns, downwards
    run(0, -1, [(y, W - 1) for y in range(0, H, 2)])                  # B: even rows, right to left
    run(-1, 0, [(H - 1, x) for x in range((W // 2) * 2 - 1, -1, -2)])  # C: odd columns, upwards
    run(0, 1, [(y, 0) for y in range((H // 2) * 2 - 1, -1, -2)])      # D: odd rows, left to right
    return labels


def ceo_bcd(proposals, lcosts, nprop, labels, bcd_times, tpsi=8, lamda=0.05):
    """ceoBCD (:261-284).  Returns the list of label images after each sweep (copies)."""
    labels = labels.astype(np.int64).copy()
    out = []
    for _ in range(bcd_times):
        sweep(proposals, lcosts, nprop, labels, tpsi, lamda)
        out.append(labels.copy())
    return out


def energy(proposals, lcosts, labels, tphi_lambda=0.05, tpsi=8):
    """MRF energy lamda*sum(lcost) + sum over 4-connected edges min(L1, tpsi) (diagnostic only)."""
    H, W = labels.shape
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    v = proposals[ys, xs, labels]
    e = tphi_lambda * lcosts[ys, xs, labels].sum()
    e += np.minimum(tpsi, np.abs(v[1:] - v[:-1]).sum(-1)).sum()
    e += np.minimum(tpsi, np.abs(v[:, 1:] - v[:, :-1]).sum(-1)).sum()
    return float(e)
