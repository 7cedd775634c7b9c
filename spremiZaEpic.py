"""Drop-in for the reference's `spremiZaEpic.py` (:1-27), the caller on the far side of the hot path: it turns the
forward and backward flow files of stage 2 into EpicFlow's three inputs and starts EpicFlow.

    python spremiZaEpic.py <image1> <image2> <forward.npy> <backward.npy> <con_tresh> <sed|canny>      (README.md:55-67)

Same argv, same files in the working directory (`sparse_field.npy`, `parovi.txt`, `ivice.bin`, `epic.flo`), same
order of work.  The consistency check and the Canny edge map run on the GPU (postprocessing.postProcessing,
edge.canny_ivice); `sed` needs opencv-contrib and a trained model and raises NotImplementedError.  EpicFlow itself is an
external binary the reference expects at ../discrete_flow/external/EpicFlow_v1.00/epicflow-static (:27); the path can
be overridden with FLOWB200_EPICFLOW, and a missing binary raises FileNotFoundError exactly as `subprocess.run` does in
the reference.
"""
import os
import sys
from subprocess import run

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from edge import canny_ivice, sed_ivice          # noqa: E402
from napravi_parove import parovi                # noqa: E402
from postprocessing import postProcessing        # noqa: E402

EPICFLOW = os.environ.get("FLOWB200_EPICFLOW", "../discrete_flow/external/EpicFlow_v1.00/epicflow-static")


def main(argv):
    kitti1, kitti2 = argv[1], argv[2]
    foward, backward = argv[3], argv[4]
    con_tresh = int(argv[5])
    postProcessing(foward, backward, con_tresh, "sparse_field.npy")
    parovi("sparse_field.npy", "parovi.txt")
    if argv[6] == "sed":
        sed_ivice(kitti1, "ivice.bin")
    elif argv[6] == "canny":
        canny_ivice(kitti1, "ivice.bin")
    run([EPICFLOW, kitti1, kitti2, "ivice.bin", "parovi.txt", "epic.flo"])


if __name__ == "__main__":
    main(sys.argv)
