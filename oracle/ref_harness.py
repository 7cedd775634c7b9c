"""Run the reference's OWN source, unmodified except for its hard-coded size constants.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Works only where /root/reference exists (the build
container); it is how tests/golden/*.npz were produced (tests/golden/make_golden.py) and how the
numpy/C restatements in this package were pinned.  Nothing here is copied from the reference: the
script text is read from /root/reference at run time, the active `picw/pich/cellw/cellh[/maxnprop]`
assignments are substituted, and the text is exec()'d with

  * `pyflann` replaced by an exact brute-force shim (FLANN is approximate and unseeded),
  * `cv2.imread` / `cv2.xfeatures2d.DAISY_create` / `cv2.KeyPoint` replaced by stand-ins that hand
    the script our synthetic images and the descriptors of oracle.daisy (opencv-contrib is absent),
  * `numpy.random.normal` wrapped so the accepted Gaussian draws can be replayed on the GPU.
"""
import contextlib
import io
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("FLOWB200_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "python bcd.py"))


def _patched_source(fname, consts):
    with open(os.path.join(REF, fname), "r", encoding="utf-8") as f:
        src = f.read()
    for name, val in consts.items():
        pat = re.compile(r"^%s\s*=\s*\d+\s*$" % re.escape(name), re.M)
        if not pat.search(src):
            raise RuntimeError(f"constant {name} not found in {fname}")
        src = pat.sub(f"{name} = {int(val)}", src)
    return src


class _ExactFlann:
    """Stand-in for pyflann.FLANN: exact float64 brute force, ties -> lowest index."""

    def build_index(self, pts=None, **kw):
        self.pts = np.asarray(pts, dtype=np.float32).astype(np.float64)
        return {}

    def nn_index(self, qpts=None, num_neighbors=1, **kw):
        q = np.asarray(qpts, dtype=np.float32).astype(np.float64).reshape(1, -1)
        d = np.zeros(self.pts.shape[0], dtype=np.float64)
        for j in range(q.shape[1]):
            diff = q[0, j] - self.pts[:, j]
            d += diff * diff
        idx = np.argsort(d, kind="stable")[:num_neighbors]
        return idx.reshape(1, -1).astype(np.int32), d[idx].reshape(1, -1)


def _fake_pyflann():
    m = types.ModuleType("pyflann")
    m.FLANN = _ExactFlann
    return m


def _fake_cv2(images, daisy_fn):
    real = None
    try:
        import cv2 as real  # noqa
    except Exception:
        pass
    m = types.ModuleType("cv2")
    if real is not None:
        for k in ("cvtColor", "COLOR_BGR2GRAY"):
            setattr(m, k, getattr(real, k))

    def imread(path, *a):
        key = os.path.basename(path)
        if key not in images:
            return None
        return images[key].copy()

    class _Daisy:
        def compute(self, picture, kp):
            d = daisy_fn(np.ascontiguousarray(picture))
            return kp, d.reshape(-1, d.shape[-1])

    xf = types.SimpleNamespace(DAISY_create=lambda **kw: _Daisy())
    m.imread = imread
    m.xfeatures2d = xf
    m.KeyPoint = lambda x, y, s: (x, y, s)
    return m


@contextlib.contextmanager
def _patched_env(modules, argv, cwd):
    old_mods = {k: sys.modules.get(k) for k in modules}
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.modules.update(modules)
    sys.argv = argv
    os.chdir(cwd)
    try:
        yield
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        for k, v in old_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_stage1(workdir, img_src, img_tgt, pair, backward, dopython, cellw, cellh, seed=0,
               daisy_fn=None, maxnprop=None, quiet=True):
    """exec `daisy i flann.py <pair> <backward> <dopython>` on a synthetic pair.

    img_src is the image the flow starts from (frame `_1{backward}`), img_tgt the other frame.
    Returns dict with every file the script wrote plus 'draws' (H,W,25,2): the accepted (tgy,tgx)."""
    from . import daisy as odaisy
    daisy_fn = daisy_fn or odaisy.daisy
    H, W, _ = img_src.shape
    nn = f"{int(pair):02d}"
    b = str(int(backward))
    images = {f"0001{nn}_1{b}.png": img_src, f"0001{nn}_1{1 - int(b)}.png": img_tgt}
    consts = {"picw": W, "pich": H, "cellw": cellw, "cellh": cellh}
    if maxnprop is not None:
        consts["maxnprop"] = maxnprop
    src = _patched_source("daisy i flann.py", consts)
    raw = []
    real_normal = np.random.normal

    def rec_normal(loc=0.0, scale=1.0, size=None):
        v = real_normal(loc, scale, size)
        raw.append(v)
        return v
    np.random.seed(seed)
    np.random.normal = rec_normal
    out = io.StringIO()
    try:
        with _patched_env({"cv2": _fake_cv2(images, daisy_fn), "pyflann": _fake_pyflann()},
                          ["daisy i flann.py", str(pair), b, "1" if dopython else "0"], workdir):
            with contextlib.redirect_stdout(out if quiet else sys.stdout):
                exec(compile(src, "daisy i flann.py", "exec"), {"__name__": "__main__"})
    finally:
        np.random.normal = real_normal
    # replay the acceptance logic (:216-233) over the recorded stream to get the accepted draws
    draws = np.zeros((H, W, 25, 2), dtype=np.int16)
    it = iter(raw)
    for x in range(W):
        for y in range(H):
            i = 0
            while i < 25:
                tgy = int(next(it))
                if 0 <= tgy < H:
                    tgx = int(next(it))
                    if 0 <= tgx < W:
                        draws[y, x, i] = (tgy, tgx)
                        i += 1
    assert next(it, None) is None, "draw stream not fully consumed"
    res = {"draws": draws}
    pre = f"Daisy output slike 1{nn} backward={b} "
    for key, fn in (("flow00", f"Gotova flow slika 1{nn} backward={b} posle 00 BCD.npy"),
                    ("labels00", f"Bestlabels fajl slike 1{nn} backward={b} posle 00 BCD.npy"),
                    ("proposals", pre + "proposals_nakon_gausa.npy"),
                    ("lcosts", pre + "lcosts_nakon_gausa.npy"),
                    ("nprop", pre + "nprop.npy"),
                    ("packedksets", pre + "packedksets.npy")):
        p = os.path.join(workdir, fn)
        if os.path.exists(p):
            res[key] = np.load(p)
    for i in range(4):
        p = os.path.join(workdir, pre + f"pakovani za c {i}.npy")
        if os.path.exists(p):
            res[f"zac{i}"] = np.load(p)
    return res


def run_stage2(workdir, H, W, pair, backward, bcd_times, maxnprop=None, quiet=True):
    """exec `python bcd.py <pair> <backward> <bcd_times>` in workdir (which holds the stage-1 files).
    Returns (list of labels after each sweep, list of flows after each sweep).  Needs H >= 36."""
    assert H >= 36, "python bcd.py:269 prints bestlabels[35]"
    nn = f"{int(pair):02d}"
    b = str(int(backward))
    consts = {"picw": W, "pich": H}
    if maxnprop is not None:
        consts["maxnprop"] = maxnprop
    src = _patched_source("python bcd.py", consts)
    out = io.StringIO()
    with _patched_env({"pyflann": _fake_pyflann()}, ["python bcd.py", str(pair), b, str(bcd_times)], workdir):
        with contextlib.redirect_stdout(out if quiet else sys.stdout):
            exec(compile(src, "python bcd.py", "exec"), {"__name__": "__main__"})
    labels, flows = [], []
    for w in range(1, bcd_times + 1):
        labels.append(np.load(os.path.join(workdir, f"Bestlabels fajl slike 1{nn} backward={b} posle {w:02d} BCD.npy")))
        flows.append(np.load(os.path.join(workdir, f"Gotova flow slika 1{nn} backward={b} posle {w:02d} BCD.npy")))
    return labels, flows


def reference_postprocessing():
    """Import the reference's postprocessing module itself (numpy only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_postprocessing", os.path.join(REF, "postprocessing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_stage1_files(workdir, pair, backward, proposals, lcosts, nprop, labels00, packedksets=None):
    """Write arrays under the reference's stage-1 file names (daisy i flann.py:200-202, 249-253, 308)."""
    nn = f"{int(pair):02d}"
    b = str(int(backward))
    pre = os.path.join(workdir, f"Daisy output slike 1{nn} backward={b} ")
    np.save(pre + "proposals_nakon_gausa.npy", np.asarray(proposals, dtype=np.int64))
    np.save(pre + "lcosts_nakon_gausa.npy", np.asarray(lcosts, dtype=np.float64))
    np.save(pre + "nprop.npy", np.asarray(nprop, dtype=np.int64))
    np.save(os.path.join(workdir, f"Bestlabels fajl slike 1{nn} backward={b} posle 00 BCD.npy"),
            np.asarray(labels00, dtype=np.int64))
    if packedksets is not None:
        np.save(pre + "packedksets.npy", packedksets)
