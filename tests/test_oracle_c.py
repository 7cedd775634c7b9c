"""The C oracle (oracle/flow_oracle.c) against the golden outputs of the reference's own source and the
numpy restatements (CPU only)."""
import numpy as np
import pytest

from helpers import load_case, load_npz, oracle_params
from oracle import bcd as obcd
from oracle import consistency as ocons
from oracle import cport
from oracle import proposals as oprop


@pytest.mark.parametrize("name,dirs", [("pair_a", (0, 1)), ("pair_b", (0,))])
def test_c_generisi_golden(name, dirs):
    z = load_case(name)
    p = oracle_params(z["meta"])
    for b in dirs:
        d1, d2 = (z["desc1"], z["desc2"]) if b == 0 else (z["desc2"], z["desc1"])
        P, L, N, B = cport.generisi(d1, d2, p)
        assert np.array_equal(B, z[f"b{b}_labels00"])
        assert np.array_equal(oprop.final_flow(P, B), z[f"b{b}_flow00"])
        P2, L2, N2, B2 = oprop.generisi(d1, d2, p)
        assert np.array_equal(P, P2) and np.array_equal(L, L2) and np.array_equal(N, N2)
        P, L, N = cport.nasumicni(d1, d2, P, L, N, B, p, draws=z[f"b{b}_draws"])
        assert np.array_equal(P, z[f"b{b}_proposals"])
        assert np.array_equal(L, z[f"b{b}_lcosts"])
        assert np.array_equal(N, z[f"b{b}_nprop"])
        P3, L3, N3 = cport.nasumicni(d1, d2, P2, L2, N2, B2, p, seed=3)        # own RNG: plausibility only
        kept = N3 - N2
        assert kept.min() >= 0 and kept.max() <= p.n_gauss and abs(kept.mean() - (N - N2).mean()) < 1.5


@pytest.mark.parametrize("name,dirs", [("pair_a", (0, 1)), ("pair_b", (0,))])
def test_c_bcd_golden(name, dirs):
    z = load_case(name)
    sweeps = int(z["meta"][5])
    for b in dirs:
        got = cport.ceo_bcd(z[f"b{b}_proposals"], z[f"b{b}_lcosts"], z[f"b{b}_nprop"], z[f"b{b}_labels00"], sweeps)
        for w in range(sweeps):
            assert np.array_equal(got[w], z[f"b{b}_labels{w + 1:02d}"]), f"sweep {w + 1} dir {b}"


def test_c_bcd_quantised_and_thread_independent():
    z = load_npz("bcd_q12")
    sweeps, shift = int(z["meta"][5]), int(z["meta"][6])
    lq = 20.0 * z["m"].astype(np.float64) / float(1 << shift)
    n0 = cport.num_threads()
    try:
        for nt in (1, 3):
            cport.set_num_threads(nt)
            got = cport.ceo_bcd(z["proposals"], lq, z["nprop"], z["labels00"], sweeps)
            for w in range(sweeps):
                assert np.array_equal(got[w], z[f"labels{w + 1:02d}"])
    finally:
        cport.set_num_threads(n0)


def test_c_bcd_random_vs_numpy():
    rng = np.random.default_rng(5)
    H, W, K = 9, 13, 40
    prop = rng.integers(-10, 11, (H, W, K, 2))
    nprop = rng.integers(5, K + 1, (H, W))
    lc = rng.uniform(0, 2.5, (H, W, K)).astype(np.float32).astype(np.float64)
    lab0 = rng.integers(0, 5, (H, W))
    a = obcd.ceo_bcd(prop, lc, nprop, lab0, 2)
    b = cport.ceo_bcd(prop, lc, nprop, lab0, 2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_c_consistency_golden():
    z = load_npz("consistency")
    for k in ("const", "randint", "real"):
        out = cport.forward_backward_consistency(ocons.ucitaj_flow(z[k + "_fwd"]), ocons.ucitaj_flow(z[k + "_bwd"]),
                                                 float(z[k + "_thr"]))
        assert np.array_equal(out, z[k + "_out"]), k


def test_c_remove_small_segments_golden():
    """fo_remove_small_segments vs the reference's own removeSmallSegments (tests/golden/make_golden_segments.py)."""
    z = load_npz("segments")
    changed = 0
    for i in range(int(z["n"])):
        f, (tresh, ms) = z[f"c{i}_in"], z[f"c{i}_par"]
        out = cport.remove_small_segments(f, float(tresh), int(ms))
        assert np.array_equal(out[..., :2], f[..., :2])
        assert np.array_equal(out[..., 2], z[f"c{i}_valid_out"].astype(np.float32)), i
        changed += int((out[..., 2] != f[..., 2]).sum())
    assert changed > 3000     # the fixtures do remove segments
