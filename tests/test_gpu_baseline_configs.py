"""Parity at the sizes the benchmark times (BASELINE.json configs[1] and configs[2]) against the C oracle
(oracle/flow_oracle.c, OpenMP; pinned bit for bit by tests/test_oracle_c.py against outputs of the reference's own
source), stage by stage over the FULL arrays:

  (i)   DAISY                      vs oracle.daisy                       <= 1e-4 relative (bit-identical expected)
  (ii)  proposal search            vs cport.generisi                     proposals, float32 costs, nprop, bestlabels
  (iii) random proposals           vs cport.nasumicni, replayed draws    proposals, costs, nprop
  (iv)  BCD, every mode            vs cport.ceo_bcd after EVERY sweep    int32 modes on costs quantised to 20*m/2^12,
                                                                         float64 mode on the unquantised costs
  (v)   the whole int32 pipeline (what bench.py times: INT32_F32COST, random proposals on) vs the oracle's float64
        pipeline on unquantised costs: EPE within 0.01 px (north star); the label mismatch fraction is printed.

The oracle needs about a minute per direction and configuration on 16 host threads (exact float64 search), so the
module takes a few minutes; FLOWB200_FULLSIZE=0 skips it.
"""
import os
import time

import numpy as np
import pytest

from helpers import pkg

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FLOWB200_FULLSIZE", "1") == "0", reason="FLOWB200_FULLSIZE=0")]

# name -> (H, W, K, sweeps the oracle runs with quantised costs, sweeps it runs with float64 costs, both directions)
CONFIGS = {
    "1024x436_K300_bcd4": (436, 1024, 300, 4, 4, True),     # BASELINE.json configs[1]
    "1242x375_K500_bcd8": (375, 1242, 500, 8, 3, False),    # BASELINE.json configs[2]
}
SHIFT = 12


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def make_draws(H, W, n, sigma, seed):
    """Accepted in-bounds (tgy, tgx) samples as nasumicni draws them (daisy i flann.py:216-222): int() of a normal
    around the pixel, redrawn until inside the image.  int16 (H, W, n, 2)."""
    rng = np.random.default_rng(seed)
    out = np.empty((H, W, n, 2), np.int16)
    for c, (size, centre) in enumerate(((H, np.arange(H)[:, None, None]), (W, np.arange(W)[None, :, None]))):
        v = np.trunc(rng.normal(centre, sigma, (H, W, n)))
        bad = (v < 0) | (v >= size)
        while bad.any():
            redo = np.trunc(rng.normal(np.broadcast_to(centre, bad.shape)[bad], sigma))
            v[bad] = redo
            bad = (v < 0) | (v >= size)
        out[..., c] = v
    return out


class Case:
    pass


@pytest.fixture(scope="module", params=list(CONFIGS))
def case(request):
    """Oracle and GPU stage 1 of one configuration (forward; backward too where the EPE gate needs it)."""
    from oracle import cport, daisy as od, proposals as oprop
    ops, params, synth, ioc = pkg("ops"), pkg("params"), pkg("synth"), pkg("io_contract")
    H, W, K, sw_q, sw_f, both = CONFIGS[request.param]
    c = Case()
    c.name, c.H, c.W, c.K, c.sw_q, c.sw_f, c.both = request.param, H, W, K, sw_q, sw_f, both
    c.p = params.for_k(K, H=H, W=W, knn_mode=1)
    c.op = oprop.Params(H, W, c.p.cellw, c.p.cellh, k_cell=c.p.k_cell, n_gauss=c.p.n_gauss, maxnprop=K)
    c.img = synth.make_pair(H, W, 5)[:2]
    c.gt = synth.make_pair(H, W, 5)[2]
    t0 = time.time()
    c.odesc = [od.daisy(c.img[0]), od.daisy(c.img[1])]
    c.gdesc = [ops.daisy(dev(c.img[0])), ops.daisy(dev(c.img[1]))]
    c.dirs = (0, 1) if both else (0,)
    c.o, c.g = {}, {}
    for d in c.dirs:
        s, t = c.odesc[d], c.odesc[1 - d]
        P, L, N, B = cport.generisi(s, t, c.op)
        draws = make_draws(H, W, c.p.n_gauss, c.p.sigma, 77 + d)
        P2, L2, N2 = cport.nasumicni(s, t, P, L, N, B, c.op, draws=draws)
        c.o[d] = dict(P=ioc.pack_proposals(P), L=L, N=N, B=B, P2=ioc.pack_proposals(P2), L2=L2, N2=N2, draws=draws)
        del P, P2
        gs, gt_ = c.gdesc[d], c.gdesc[1 - d]
        pvec, lcost, nprop, labels = ops.knn_proposals(gs, gt_, c.p)
        g = dict(P=pvec.cpu().numpy(), L=lcost.cpu().numpy(), N=nprop.cpu().numpy(), B=labels.cpu().numpy())
        ops.random_proposals(gs, gt_, c.p, pvec, lcost, nprop, labels, draws=dev(draws))
        g.update(pvec=pvec, lcost=lcost, nprop=nprop, labels=labels)
        c.g[d] = g
    print(f"\n[{c.name}] stage 1 of {len(c.dirs)} direction(s): oracle + GPU in {time.time() - t0:.1f} s "
          f"({cport.num_threads()} host threads)")
    return c


def test_daisy_full_size(case):
    for o, g in zip(case.odesc, case.gdesc):
        got = g.cpu().numpy()
        scale = np.abs(o).max(axis=-1, keepdims=True) + 1e-12
        rel = float((np.abs(got - o) / scale).max())
        assert rel <= 1e-4, rel
        assert ((o == 0) == (got == 0)).all()
        print(f"[{case.name}] DAISY max relative difference {rel:.2e}, bit-identical: {np.array_equal(got, o)}")


def test_search_full_arrays(case):
    """(ii) every proposal slot, data cost, nprop and bestlabels of the tcgen05 search equal the oracle's exact search."""
    for d in case.dirs:
        o, g = case.o[d], case.g[d]
        assert np.array_equal(g["N"], o["N"])
        assert np.array_equal(g["B"], o["B"])
        assert np.array_equal(g["P"], o["P"]), float((g["P"] != o["P"]).mean())
        assert np.array_equal(g["L"].astype(np.float64), o["L"])


def test_random_proposals_full_arrays(case):
    """(iii) nasumicni on replayed draws, quirks Q4 / Q6 included."""
    for d in case.dirs:
        o, g = case.o[d], case.g[d]
        assert np.array_equal(g["nprop"].cpu().numpy(), o["N2"])
        assert np.array_equal(g["pvec"].cpu().numpy(), o["P2"])
        assert np.array_equal(g["lcost"].cpu().numpy().astype(np.float64), o["L2"])
        kept = float((o["N2"] - o["N"]).mean())
        print(f"[{case.name}] direction {d}: nprop mean {o['N2'].mean():.1f}, {kept:.1f} of {case.p.n_gauss} draws kept")


def _oracle_inputs(case, d):
    ioc = pkg("io_contract")
    o = case.o[d]
    return ioc.unpack_proposals(o["P2"]), o["L2"], o["N2"], o["B"]


def test_bcd_int32_every_sweep(case):
    """(iv) both int32 entry points (integer costs m; float32 costs quantised on the fly) against the oracle's float64
    programme on costs 20*m/2^12, after every sweep."""
    from oracle import cport
    ops, lib = pkg("ops"), pkg("_lib")
    P, L, N, B = _oracle_inputs(case, 0)
    used = L != 1000.0
    m = np.where(used, np.rint(0.05 * L * float(1 << SHIFT)), 0.0)
    lq = np.where(used, 20.0 * m / float(1 << SHIFT), 1000.0)
    t0 = time.time()
    want = cport.ceo_bcd(P, lq, N, B, case.sw_q)
    t1 = time.time()
    g = case.g[0]
    lab = g["labels"].clone()
    snaps = ops.bcd(g["pvec"], dev(m.astype(np.int32)), g["nprop"], lab, case.sw_q, mode=lib.BCD_INT32,
                    cost_shift=SHIFT, per_sweep=True).cpu().numpy()
    lab2 = g["labels"].clone()
    snaps2 = ops.bcd(g["pvec"], g["lcost"], g["nprop"], lab2, case.sw_q, mode=lib.BCD_INT32_F32COST,
                     cost_shift=SHIFT, per_sweep=True).cpu().numpy()
    for w in range(case.sw_q):
        assert np.array_equal(snaps[w], want[w]), (w, float((snaps[w] != want[w]).mean()))
        assert np.array_equal(snaps2[w], want[w]), (w, float((snaps2[w] != want[w]).mean()))
    changed = float((want[-1] != B).mean())
    print(f"[{case.name}] int32 BCD: {case.sw_q} sweeps bit-exact; oracle {t1 - t0:.1f} s; {changed:.3f} of the labels moved")
    case.int_labels = want[-1]


def test_bcd_fp64_every_sweep(case):
    """(iv) the float64 mode on the unquantised float32 costs (what `bcd.py` runs on reference-written files)."""
    from oracle import cport
    ops, lib = pkg("ops"), pkg("_lib")
    P, L, N, B = _oracle_inputs(case, 0)
    want = cport.ceo_bcd(P, L, N, B, case.sw_f)
    g = case.g[0]
    lab = g["labels"].clone()
    snaps = ops.bcd(g["pvec"], g["lcost"], g["nprop"], lab, case.sw_f, mode=lib.BCD_FP64_F32COST,
                    per_sweep=True).cpu().numpy()
    for w in range(case.sw_f):
        assert np.array_equal(snaps[w], want[w]), (w, float((snaps[w] != want[w]).mean()))
    case.fp64_labels = {0: want[-1]}


def test_int32_pipeline_epe_against_float64_oracle(case):
    """(v) GPU: DAISY -> search -> random proposals -> INT32_F32COST BCD -> check, as bench.py runs it;
    oracle: the same stages in float64 on unquantised costs.  EPE (visualization.py:128-152) within 0.01 px."""
    if not case.both:
        pytest.skip("forward-only configuration")
    from oracle import consistency as ocons, cport, epe as oepe, proposals as oprop
    ops, lib, synth = pkg("ops"), pkg("_lib"), pkg("synth")
    gflow, oflow, mism = [], [], []
    for d in case.dirs:
        P, L, N, B = _oracle_inputs(case, d)
        have = getattr(case, "fp64_labels", {}).get(d) if case.sw_f == case.sw_q else None
        want = have if have is not None else cport.ceo_bcd(P, L, N, B, case.sw_q)[-1]
        oflow.append(oprop.final_flow(P, want))
        g = case.g[d]
        lab = g["labels"].clone()
        ops.bcd(g["pvec"], g["lcost"], g["nprop"], lab, case.sw_q, mode=lib.BCD_INT32_F32COST, cost_shift=SHIFT)
        gflow.append(ops.flow_from_labels(g["pvec"], lab, want_yx=False)[1])
        mism.append(float((lab.cpu().numpy() != want).mean()))
    got = ops.consistency(gflow[0].clone(), gflow[1], case.p.con_tresh).cpu().numpy()
    want = ocons.post_processing(oflow[0], oflow[1], case.p.con_tresh)
    gt = synth.gt_uvv(case.gt)
    e_got, e_want = oepe.error_image(got, gt), oepe.error_image(want, gt)
    e_raw_got = oepe.error_image(gflow[0].cpu().numpy(), gt)
    e_raw_want = oepe.error_image(ocons.ucitaj_flow(oflow[0]), gt)
    print(f"[{case.name}] int32 pipeline vs float64 oracle: label mismatch fwd {mism[0]:.5f} bwd {mism[1]:.5f}; "
          f"EPE checked {e_got[0]:.4f} vs {e_want[0]:.4f} px, raw forward {e_raw_got[0]:.4f} vs {e_raw_want[0]:.4f} px; "
          f"outliers {e_got[1]:.3f} vs {e_want[1]:.3f} %")
    assert abs(e_got[0] - e_want[0]) <= 0.01, (e_got, e_want)
    assert abs(e_raw_got[0] - e_raw_want[0]) <= 0.01, (e_raw_got, e_raw_want)
    # the device-side metric kernel agrees with the definition on this field
    dm = ops.epe(dev(got), dev(gt))
    assert dm[2] == e_got[2] and abs(dm[0] - e_got[0]) <= 1e-5 * max(1.0, e_got[0])


def test_flow_pair_equals_the_stages(case):
    """flowb200_flow_pair (one call, two streams, Philox draws) == the stage calls one after the other with the same
    seeds: what the benchmark times is the path the tests above checked piece by piece."""
    if not case.both:
        pytest.skip("forward-only configuration")
    ops, lib = pkg("ops"), pkg("_lib")
    g0, g1 = dev(case.img[0]), dev(case.img[1])
    out, raw_f, raw_b = ops.flow_pair(g0, g1, case.p, sweeps=case.sw_q, directions=2, seed=9,
                                      bcd_mode=lib.BCD_INT32_F32COST, want_raw=True)
    flows = []
    for d in (0, 1):
        s, t = case.gdesc[d], case.gdesc[1 - d]
        pvec, lcost, nprop, labels = ops.knn_proposals(s, t, case.p)
        ops.random_proposals(s, t, case.p, pvec, lcost, nprop, labels, seed=9 + d)
        ops.bcd(pvec, lcost, nprop, labels, case.sw_q, mode=lib.BCD_INT32_F32COST, cost_shift=case.p.cost_shift)
        flows.append(ops.flow_from_labels(pvec, labels, want_yx=False)[1])
    assert torch.equal(raw_f, flows[0]) and torch.equal(raw_b, flows[1])
    assert torch.equal(out, ops.consistency(flows[0].clone(), flows[1], case.p.con_tresh))
