"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference SpatialActionTokenizer grid arithmetic
(model/action_tokenizer.py).  Works on integer *local* ids (0..8193); the reference's `<ACTION%05d>` strings
map 1:1 to `action_token_begin_idx + local id`.  Pinned by tests/golden/tokenizer_*.npz, which hold outputs of
the live reference (oracle/gen_golden.py)."""
from __future__ import annotations

import numpy as np
from scipy.stats import norm

RANGE_BINS = {  # model/action_tokenizer.py:250-261
    "translation": {"theta_bins": (0.0, np.pi), "phi_bins": (-np.pi, np.pi), "r_bins": (0.0, np.sqrt(3))},
    "rotation": {"roll_bins": (-1.0, 1.0), "pitch_bins": (-1.0, 1.0), "yaw_bins": (-1.0, 1.0)},
}


def get_bin_policy(num_bins, gs_params=None, min_sigma=0.0):
    """model/action_tokenizer.py:343-370"""
    pol = {"translation": {}, "rotation": {}}
    for bt, d in RANGE_BINS.items():
        for bk, (lo, hi) in d.items():
            n = num_bins[bt][bk]
            if gs_params is None:
                pol[bt][bk] = np.linspace(lo, hi, n + 1)
            else:
                g = gs_params[bk.split("_")[0].lower()]
                mu, sigma = g["mu"], max(g["sigma"], min_sigma)
                prob = np.linspace(norm.cdf(lo, loc=mu, scale=sigma), norm.cdf(hi, loc=mu, scale=sigma), n + 1)
                pol[bt][bk] = np.clip(norm.ppf(prob, loc=mu, scale=sigma), lo, hi).tolist()
    return pol


def encode(actions, pol, num_bins, min_action=-1.0, max_action=1.0):
    """model/action_tokenizer.py:305-319 -> :105-119, :177-188, :227-233. (n,7) float64 -> (n,3) local ids"""
    a = np.clip(np.asarray(actions, dtype=np.float64).reshape(-1, 7), min_action, max_action)
    t, r = pol["translation"], pol["rotation"]
    nt, nr = num_bins["translation"], num_bins["rotation"]
    x, y, z = a[:, 0], a[:, 1], a[:, 2]
    theta = np.arctan2(np.sqrt(x ** 2 + y ** 2), z)
    phi = np.arctan2(y, x)
    rad = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    dt = np.digitize(theta, np.asarray(t["theta_bins"])[1:-1])
    dp = np.digitize(phi, np.asarray(t["phi_bins"])[1:-1])
    dr = np.digitize(rad, np.asarray(t["r_bins"])[1:-1])
    tid = dt * (nt["phi_bins"] * nt["r_bins"]) + dp * nt["r_bins"] + dr
    d0 = np.clip(np.digitize(a[:, 3], np.asarray(r["roll_bins"])) - 1, 0, nr["roll_bins"] - 1)
    d1 = np.clip(np.digitize(a[:, 4], np.asarray(r["pitch_bins"])) - 1, 0, nr["pitch_bins"] - 1)
    d2 = np.clip(np.digitize(a[:, 5], np.asarray(r["yaw_bins"])) - 1, 0, nr["yaw_bins"] - 1)
    n_trans = nt["theta_bins"] * nt["phi_bins"] * nt["r_bins"]
    n_rot = nr["roll_bins"] * nr["pitch_bins"] * nr["yaw_bins"]
    rid = d0 * (nr["pitch_bins"] * nr["yaw_bins"]) + d1 * nr["yaw_bins"] + d2 + n_trans
    gid = np.where(a[:, 6] >= 0.5, 1, 0) + n_trans + n_rot
    return np.stack([tid, rid, gid], 1).astype(np.int64)


def decode(local_ids, pol, num_bins):
    """model/action_tokenizer.py:321-333 -> :121-137, :190-202, :235-243. (n,3) local ids -> (n,7) float64"""
    ids = np.asarray(local_ids, dtype=np.int64).reshape(-1, 3)
    t, r = pol["translation"], pol["rotation"]
    nt, nr = num_bins["translation"], num_bins["rotation"]
    n_trans = nt["theta_bins"] * nt["phi_bins"] * nt["r_bins"]
    n_rot = nr["roll_bins"] * nr["pitch_bins"] * nr["yaw_bins"]
    i0 = np.clip(ids[:, 0], 0, n_trans - 1)
    NP = nt["phi_bins"] * nt["r_bins"]
    a, b, c = i0 // NP, (i0 % NP) // nt["r_bins"], i0 % nt["r_bins"]
    tb, pb, rb = (np.asarray(t[k]) for k in ("theta_bins", "phi_bins", "r_bins"))
    th = 0.5 * (tb[a] + tb[a + 1])
    ph = 0.5 * (pb[b] + pb[b + 1])
    rr = 0.5 * (rb[c] + rb[c + 1])
    x = rr * np.sin(th) * np.cos(ph)
    y = rr * np.sin(th) * np.sin(ph)
    z = rr * np.cos(th)
    x, y, z = np.clip([x, y, z], -1, 1)
    i1 = np.clip(ids[:, 1], n_trans, n_trans + n_rot - 1) - n_trans
    NP = nr["pitch_bins"] * nr["yaw_bins"]
    a, b, c = i1 // NP, (i1 % NP) // nr["yaw_bins"], i1 % nr["yaw_bins"]
    ro, pi_, ya = (np.asarray(r[k]) for k in ("roll_bins", "pitch_bins", "yaw_bins"))
    roll = 0.5 * (ro[a] + ro[a + 1])
    pitch = 0.5 * (pi_[b] + pi_[b + 1])
    yaw = 0.5 * (ya[c] + ya[c + 1])
    i2 = np.clip(ids[:, 2], n_trans + n_rot, n_trans + n_rot + num_bins["gripper"] - 1) - n_trans - n_rot
    g = np.where(i2 == 0, 0.0, 1.0)
    return np.stack([x, y, z, roll, pitch, yaw, g], 1)


def encode_angles_exact(actions, pol, num_bins, trig, phi_nonpos, phi_neg, min_action=-1.0, max_action=1.0):
    """CPU restatement of the atan2-free angular binning of csrc/tokenizer.cu (svla_tok_encode with an edge_trig table), used to
    check the decision rule, the table and the quadrant logic against the golden ids without a GPU: the fast path is the same
    rounded double expression with the same bound; the slow path (error-free transformations on the device) is done here in exact
    rational arithmetic.  Returns (theta bin, phi bin) int arrays."""
    from fractions import Fraction
    a = np.clip(np.asarray(actions, dtype=np.float64).reshape(-1, 7), min_action, max_action)
    te = np.asarray(pol["translation"]["theta_bins"], dtype=np.float64)[1:-1]
    pe = np.asarray(pol["translation"]["phi_bins"], dtype=np.float64)[1:-1]
    n_ti, n_pi = te.shape[0], pe.shape[0]
    trig = np.asarray(trig, dtype=np.float64).reshape(-1, 4)

    def ge(av, bv, row):
        ch, cl, sh, sl = (float(v) for v in trig[row])
        p1, p2 = av * ch, bv * sh
        d = p1 - p2
        if abs(d) > 1.8e-15 * (abs(p1) + abs(p2)):
            return d > 0.0
        tot = Fraction(av) * (Fraction(ch) + Fraction(cl)) - Fraction(bv) * (Fraction(sh) + Fraction(sl))
        return tot >= 0

    def count(av, bv, base, lo, hi):
        while lo < hi:
            mid = (lo + hi) >> 1
            if ge(av, bv, base + mid):
                lo = mid + 1
            else:
                hi = mid
        return lo

    dt, dp = np.zeros(a.shape[0], dtype=np.int64), np.zeros(a.shape[0], dtype=np.int64)
    for i, (x, y, z) in enumerate(a[:, :3]):
        x, y, z = float(x), float(y), float(z)
        rho = float(np.sqrt(x * x + y * y))
        if rho == 0.0 and z == 0.0:
            dt[i] = np.digitize(np.arctan2(rho, z), te)
        else:
            dt[i] = count(rho, z, 0, 0, n_ti)
        if y == 0.0:
            dp[i] = np.digitize(np.arctan2(y, x), pe)
        elif y > 0.0:
            dp[i] = count(y, x, n_ti, phi_nonpos, n_pi)
        else:
            dp[i] = count(y, x, n_ti, 0, phi_neg)
    return dt, dp
