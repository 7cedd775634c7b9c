"""The oracle restatements against outputs of the reference's own source (tests/golden, CPU only)."""
import numpy as np
import pytest

from helpers import load_case, load_npz, oracle_params
from oracle import bcd as obcd
from oracle import consistency as ocons
from oracle import proposals as oprop


@pytest.mark.parametrize("name,dirs", [("pair_a", (0, 1)), ("pair_b", (0,))])
def test_generisi_nasumicni_pakovanje(name, dirs):
    z = load_case(name)
    p = oracle_params(z["meta"])
    for b in dirs:
        d1, d2 = (z["desc1"], z["desc2"]) if b == 0 else (z["desc2"], z["desc1"])
        P, L, N, B = oprop.generisi(d1, d2, p)
        assert np.array_equal(B, z[f"b{b}_labels00"])
        assert np.array_equal(oprop.final_flow(P, B), z[f"b{b}_flow00"])
        oprop.nasumicni(d1, d2, P, L, N, B, p, z[f"b{b}_draws"])
        assert np.array_equal(P, z[f"b{b}_proposals"])
        assert np.array_equal(L, z[f"b{b}_lcosts"])
        assert np.array_equal(N, z[f"b{b}_nprop"])
        pk = oprop.pakovanje(P, N, p)
        assert oprop.ksets_masked_equal(pk, z[f"b{b}_packedksets"], N, p)


@pytest.mark.parametrize("name,dirs", [("pair_a", (0, 1)), ("pair_b", (0,))])
def test_bcd_fp64(name, dirs):
    z = load_case(name)
    sweeps = int(z["meta"][5])
    for b in dirs:
        got = obcd.ceo_bcd(z[f"b{b}_proposals"], z[f"b{b}_lcosts"], z[f"b{b}_nprop"], z[f"b{b}_labels00"], sweeps)
        for w in range(sweeps):
            assert np.array_equal(got[w], z[f"b{b}_labels{w + 1:02d}"]), f"sweep {w + 1} dir {b}"
            assert np.array_equal(oprop.final_flow(z[f"b{b}_proposals"], got[w]), z[f"b{b}_flow{w + 1:02d}"])


def test_bcd_quantised_costs():
    z = load_npz("bcd_q12")
    sweeps, shift = int(z["meta"][5]), int(z["meta"][6])
    lq = 20.0 * z["m"].astype(np.float64) / float(1 << shift)
    got = obcd.ceo_bcd(z["proposals"].astype(np.int64), lq, z["nprop"].astype(np.int64),
                       z["labels00"].astype(np.int64), sweeps)
    for w in range(sweeps):
        assert np.array_equal(got[w], z[f"labels{w + 1:02d}"])


def test_consistency_cases():
    z = load_npz("consistency")
    for k in ("const", "randint", "real"):
        out = ocons.post_processing(z[k + "_fwd"], z[k + "_bwd"], z[k + "_thr"].item())
        assert out.dtype == np.float32
        assert np.array_equal(out, z[k + "_out"]), k
    # quirk Q5: constant dx=+3 invalidates the bottom three ROWS
    out = z["const_out"]
    assert (out[-3:, :, 2] == 0).all() and (out[:-3, :, 2] == 1).all()


def test_consistency_pipeline_outputs():
    z = load_case("pair_a")
    for thr in (10, 2):
        out = ocons.post_processing(z["b0_flow02"], z["b1_flow02"], thr)
        assert np.array_equal(out, z[f"sparse_thr{thr}"])
