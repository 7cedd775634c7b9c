"""The drop-in entry points (same argv / files / module API as the reference scripts) on the GPU, against the
golden outputs of the reference's own source.  Interop rule: the new `bcd.py` must accept
stage-1 files as the reference writes them, and the new stage 1 must write files the reference can load."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import load_case, load_npz, pkg

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(meta, **extra):
    H, W, cw, ch = (int(v) for v in meta[:4])
    e = dict(os.environ, FLOWB200_PICW=str(W), FLOWB200_PICH=str(H), FLOWB200_CELLW=str(cw), FLOWB200_CELLH=str(ch))
    e.update(extra)
    return e


def test_bcd_script_on_reference_stage1_files(tmp_path):
    """`python bcd.py <pair> <backward> <bcd_times>` on files laid out exactly as the reference writes them."""
    from oracle import ref_harness as rh
    ioc = pkg("io_contract")
    z = load_case("pair_a")
    pair, sweeps = int(z["meta"][4]), int(z["meta"][5])
    for b in (0, 1):
        rh.write_stage1_files(str(tmp_path), pair, b, z[f"b{b}_proposals"], z[f"b{b}_lcosts"], z[f"b{b}_nprop"],
                              z[f"b{b}_labels00"], packedksets=None)      # packedksets.npy is not needed
        for script in ("bcd.py", "python bcd.py"):
            r = subprocess.run([sys.executable, os.path.join(ROOT, script), str(pair), str(b), str(sweeps)],
                               cwd=tmp_path, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            for w in range(1, sweeps + 1):
                lab = np.load(tmp_path / ioc.labels_file(pair, b, w))
                flo = np.load(tmp_path / ioc.flow_file(pair, b, w))
                assert lab.dtype == np.int64 and flo.dtype == np.float64 and flo.shape == lab.shape + (2,)
                assert np.array_equal(lab, z[f"b{b}_labels{w:02d}"]), (script, b, w)
                assert np.array_equal(flo, z[f"b{b}_flow{w:02d}"])


def test_stage1_script_files_and_contract(tmp_path):
    """`python "daisy i flann.py" <pair> <backward> <dopython>`: same file names / shapes / dtypes / fills; the
    deterministic part (NN proposals, costs, posle-00 flow and labels) equals the reference's outputs."""
    import cv2
    ioc = pkg("io_contract")
    z = load_case("pair_a")
    pair = int(z["meta"][4])
    data = tmp_path / "data_scene_flow" / "training" / "image_2"
    work = tmp_path / "work"
    data.mkdir(parents=True)
    work.mkdir()
    nn = ioc.pad2(pair)
    cv2.imwrite(str(data / f"0001{nn}_10.png"), z["img1"])
    cv2.imwrite(str(data / f"0001{nn}_11.png"), z["img2"])
    for b, dopython in ((0, 1), (1, 0)):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "daisy i flann.py"), str(pair), str(b), str(dopython)],
                           cwd=work, capture_output=True, text=True, env=_env(z["meta"], FLOWB200_SEED="5"))
        assert r.returncode == 0, r.stderr[-2000:]
        P = np.load(work / ioc.stage1_file(pair, b, "proposals_nakon_gausa"))
        L = np.load(work / ioc.stage1_file(pair, b, "lcosts_nakon_gausa"))
        N = np.load(work / ioc.stage1_file(pair, b, "nprop"))
        lab0 = np.load(work / ioc.labels_file(pair, b, 0))
        flo0 = np.load(work / ioc.flow_file(pair, b, 0))
        H, W = z["img1"].shape[:2]
        assert P.shape == (H, W, 150, 2) and P.dtype == np.int64
        assert L.shape == (H, W, 150) and L.dtype == np.float64
        assert N.shape == (H, W) and N.dtype == np.int64 and lab0.dtype == np.int64 and flo0.dtype == np.float64
        assert np.array_equal(lab0, z[f"b{b}_labels00"]) and np.array_equal(flo0, z[f"b{b}_flow00"])
        # NN slots are identical; unused slots carry the reference's fills
        from oracle import proposals as oprop
        from helpers import oracle_params
        d1, d2 = (z["desc1"], z["desc2"]) if b == 0 else (z["desc2"], z["desc1"])
        P0, L0, N0, _ = oprop.generisi(d1, d2, oracle_params(z["meta"]))
        for y in range(0, H, 7):
            for x in range(0, W, 5):
                k = int(N0[y, x])
                assert np.array_equal(P[y, x, :k], P0[y, x, :k]) and np.array_equal(L[y, x, :k], L0[y, x, :k])
                assert (P[y, x, N[y, x]:] == -1).all() and (L[y, x, N[y, x]:] == 1000.0).all()
                assert k <= N[y, x] <= k + 25
        if dopython:
            pk = np.load(work / ioc.stage1_file(pair, b, "packedksets"))
            assert pk.shape == (H, W, 2, 2813) and pk.dtype == np.uint8
            assert oprop.ksets_masked_equal(pk, oprop.pakovanje(P, N, oracle_params(z["meta"])), N,
                                            oracle_params(z["meta"]))
        else:
            shapes = [((W + 1) // 2, H, 2813), ((H + 1) // 2, W, 2813), ((W + 1) // 2, H, 2813), ((H + 1) // 2, W, 2813)]
            for i, shp in enumerate(shapes):
                a = np.load(work / ioc.stage1_file(pair, b, f"pakovani za c {i}"))
                assert a.shape == shp and a.dtype == np.uint8


def test_postprocessing_module_dropin(tmp_path):
    """import postprocessing; postProcessing(fwd.npy, bwd.npy, thr, out) and the per-pixel / whole-image calls."""
    sys.path.insert(0, ROOT)
    pp = importlib.import_module("postprocessing")
    z = load_npz("consistency")
    for k in ("const", "randint", "real"):
        f, b, o = tmp_path / f"{k}_f.npy", tmp_path / f"{k}_b.npy", tmp_path / f"{k}_o.npy"
        np.save(f, z[k + "_fwd"])
        np.save(b, z[k + "_bwd"])
        fi = pp.postProcessing(str(f), str(b), z[k + "_thr"].item(), str(o))
        assert np.array_equal(np.load(o), z[k + "_out"]) and np.array_equal(fi.flow, z[k + "_out"])
        assert (fi.height, fi.width) == z[k + "_out"].shape[:2]
    # consistencyCheck on single pixels == the whole-image call
    f1, f2 = pp.FlowImage(), pp.FlowImage()
    f1.ucitajFlow(str(tmp_path / "real_f.npy"))
    f2.ucitajFlow(str(tmp_path / "real_b.npy"))
    # the reference's callers pass the arrays (postprocessing.py:133); a FlowImage is accepted as well
    for a in range(f1.height):
        for bb in range(0, f1.width, 3):
            assert pp.consistencyCheck(f1.flow, f2.flow, a, bb, z["real_thr"].item()) is None
    assert np.array_equal(f1.flow[:, ::3], z["real_out"][:, ::3])
    assert pp.consistencyCheck(f1, f2, 0, 1, z["real_thr"].item()) is None
    assert np.array_equal(f1.flow[0, 1], z["real_out"][0, 1])
    g1 = pp.FlowImage()
    g1.ucitajFlow(str(tmp_path / "real_f.npy"))
    assert pp.fowardBackwardConsistency(g1.flow, f2.flow, z["real_thr"].item()) is None
    assert np.array_equal(g1.flow, z["real_out"])
    with pytest.raises(IndexError):
        pp.consistencyCheck(f1.flow, f2.flow, f1.height, 0, 10)
    with pytest.raises(TypeError):
        pp.fowardBackwardConsistency(g1.flow.astype(np.float64), f2.flow, 10)


def test_consistency_flow2_of_another_shape():
    """flow2 larger than flow1 is indexed by its own shape (rows/columns beyond flow1's extent are never read); a
    smaller flow2 raises IndexError as soon as a target lies beyond it (postprocessing.py:80-97, numpy indexing)."""
    sys.path.insert(0, ROOT)
    pp = importlib.import_module("postprocessing")
    from oracle import consistency as ocons
    rng = np.random.default_rng(11)
    A, B = 33, 47
    f1 = np.zeros((A, B, 3), np.float32)
    f1[..., :2] = rng.integers(-4, 5, size=(A, B, 2))
    f1[..., 2] = rng.random((A, B)) < 0.9
    big = np.zeros((A + 5, B + 9, 3), np.float32)
    big[..., :2] = rng.integers(-4, 5, size=(A + 5, B + 9, 2))
    big[..., 2] = rng.random((A + 5, B + 9)) < 0.9
    got = f1.copy()
    pp.fowardBackwardConsistency(got, big, 3.0)
    assert np.array_equal(got, ocons.forward_backward_consistency(f1, big[:A, :B], 3.0))
    with pytest.raises(IndexError):
        pp.fowardBackwardConsistency(f1.copy(), big[:A - 6, :B - 6].copy(), 3.0)
    # a smaller flow2 that is never indexed beyond its extent works like the reference: the pixels that would
    # point beyond it are invalid, and invalid pixels return before flow2 is touched (:84-85)
    ok = f1.copy()
    a2 = np.trunc(ok[..., 0] + np.arange(A, dtype=np.float32)[:, None])
    b2 = np.trunc(ok[..., 1] + np.arange(B, dtype=np.float32)[None, :])
    ok[(a2 >= A - 3) | (b2 >= B - 3), 2] = 0
    small = big[:A - 3, :B - 3].copy()
    got = ok.copy()
    pp.fowardBackwardConsistency(got, small, 3.0)
    assert np.array_equal(got, ocons.forward_backward_consistency(ok, np.pad(small, ((0, 3), (0, 3), (0, 0))), 3.0))
