#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference checkout (run in the build container only).

The reference is C++ and cannot be imported; what this script does is
  * copy the sample frames the reference's own tests use (data, not source) from
    /root/reference/samples (sz3/cimbar-samples) and record the SHA-256 of their decoded RGB pixels,
  * record the reference's golden SHA-256s with the file:line they come from,
  * record cv2-derived pins for the OpenCV arithmetic the oracle restates
    (cvtColor / adaptiveThreshold / filter2D), so a cv2 change is detected rather than silently followed.
"""
import hashlib, json, os, shutil, sys
import cv2, numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

SAMPLES = [
    "b/tr_0.png", "b/tr_1.png", "b/tr_2.png", "b/tr_3.png", "b/ex2434.jpg", "b/ex380.jpg",
    "6bit/4color_ecc30_fountain_0.png", "6bit/4_30_f0_627_extract.jpg", "mycell.png",
    # camera pictures the reference's ScannerTest scans (not extracted: 960x1280 / 1280x960)
    "6bit/4_30_f0_627.jpg", "6bit/4_30_f2_734.jpg", "6bit/4_30_f2_246.jpg", "6bit/4_30_f1_360.jpg",
]

GOLDENS = [  # (sample, mode, ecc, bytes, sha256, source)
    ("b/tr_0.png", 68, False, 9300, "ddcb6cd47751df1402dcf2cffdace212bc9e4a4b6ef097ad4828913086309469", "src/lib/encoder/test/DecoderTest.cpp:26-36"),
    ("b/tr_0.png", 68, True, 7500, "a0e9fff8cd5b13807fae215b8b07e38091d3f533ff46243b53ee7f74fbbee0d5", "src/lib/encoder/test/DecoderTest.cpp:38-48"),
    ("6bit/4color_ecc30_fountain_0.png", 4, False, 9300, "7e1919b1210ccc332fc56e8b35cccd622d980f03c6c3b32338bb00aa4b6a22a2", "src/lib/encoder/test/DecoderTest.cpp:64-77"),
    ("6bit/4color_ecc30_fountain_0.png", 4, True, 7500, "382c76644a4dff475c5793c5fe061e35e47be252010d29aeaf8d93ee6a3f7045", "src/lib/encoder/test/DecoderTest.cpp:79-90"),
    ("6bit/4_30_f0_627_extract.jpg", 4, False, 9300, "2040c157884c476def842f7854621a7655182e5f11a34ade563616d93cb93455", "src/lib/encoder/test/DecoderTest.cpp:92-106"),
    ("b/scan2434.jpg", 68, False, 9300, "ccb39ac3511a8974a8e98d3ea321d576d974d28f0b9373fc611ce6a4c83b561c", "src/lib/encoder/test/DecoderTest.cpp:50-62"),
]


# colour-correction goldens: the CCM the reference's tests print (operator<< of cv::Matx<float>, "%.8g") and the colours of
# the first six cells in flood order with that CCM active
CCM_GOLDENS = [
    {"sample": "b/ex2434.jpg", "mode": 68, "source": "src/lib/cimb_translator/test/CimbReaderTest.cpp:181-214",
     "matrix": [[2.3991191, -0.41846275, -0.54654282], [-0.42976046, 2.632102, -0.76466882], [-0.54299992, -0.20199311, 2.2753253]],
     "first_colors": [0, 1, 1, 2, 2, 2]},
    {"sample": "b/ex2434.jpg", "mode": 68, "source": "src/lib/cimb_translator/test/CimbReaderTest.cpp:216-237 (CCM disabled)",
     "matrix": None, "first_colors": [0, 1, 1, 2, 2, 2]},
    {"sample": "b/ex380.jpg", "mode": 68, "source": "src/lib/cimb_translator/test/CimbReaderTest.cpp:239-272",
     "matrix": [[1.6250746, 0.0024788622, -0.45772526], [-0.29126319, 2.2922182, -0.67037439], [-1.2192062, -2.7447209, 5.0476217]],
     "first_colors": [0, 1, 1, 2, 2, 2]},
]
# color_correctionTest/testComputeMoorePenrose (.2): color_correction::get_moore_penrose_lsm on explicit inputs
MOORE_PENROSE_GOLDENS = [
    {"actual": [[0, 142.31060606, 0], [0, 148.75, 148.75], [148.75, 148.75, 0], [148.75, 0, 148.75], [255, 255, 255]],
     "desired": [[0, 255, 0], [0, 255, 255], [255, 255, 0], [255, 0, 255], [255, 255, 255]],
     "matrix_str": "[1.5223049, -0.10023587, -0.19198087;\n -0.20533442, 1.6441474, -0.20533434;\n -0.19198078, -0.10023584, 1.5223049]",
     "source": "src/lib/chromatic_adaptation/test/color_correctionTest.cpp:32-56"},
    {"actual": [[14.58901515, 115.74431818, 39.88320707], [19.34027778, 124.4375, 115.37152778], [140.70486111, 137.45833333, 65.50694444],
                [131.59722222, 41.22222222, 104.84027778], [171.625, 163.625, 158.875]],
     "desired": [[0, 255, 0], [0, 255, 255], [255, 255, 0], [255, 0, 255], [255, 255, 255]],
     "matrix_str": "[2.0261116, -0.21691091, -0.19806443;\n -0.43822661, 2.4562523, -0.41700464;\n -0.55769891, -1.1443435, 3.4819376]",
     "source": "src/lib/chromatic_adaptation/test/color_correctionTest.cpp:58-84"},
]
# CimbReaderTest/testCCM(.VeryNecessary): init_ccm after update_metadata(FountainMetadata(0, 23586, 7)) -> printed matrix
INIT_CCM_GOLDENS = [
    {"sample": "b/ex2434.jpg", "md": [0, 23586, 7],
     "matrix_str": "[2.3991191, -0.41846275, -0.54654282;\n -0.42976046, 2.632102, -0.76466882;\n -0.54299992, -0.20199311, 2.2753253]",
     "source": "src/lib/cimb_translator/test/CimbReaderTest.cpp:181-199"},
    {"sample": "b/ex380.jpg", "md": [0, 23586, 7],
     "matrix_str": "[1.6250746, 0.0024788622, -0.45772526;\n -0.29126319, 2.2922182, -0.67037439;\n -1.2192062, -2.7447209, 5.0476217]",
     "source": "src/lib/cimb_translator/test/CimbReaderTest.cpp:239-257"},
]
# ScannerTest: Anchor strings are "xavg+-xrange,yavg+-yrange" (Anchor.h:107-111)
SCAN_GOLDENS = [
    {"sample": "6bit/4_30_f0_627.jpg", "source": "src/lib/extractor/test/ScannerTest.cpp:16-72, :136-147",
     "t1_contains": ["210+-25,912+-0", "195+-25,64+-0", "1039+-23,880+-0"],
     "t2_contains": ["210+-0,914+-24", "195+-0,61+-25", "1039+-0,887+-23"],
     "t3_contains": ["210+-23,914+-24", "195+-25,61+-25", "1039+-22,887+-23"],
     "piecemeal_filtered": "195+-25,61+-25 210+-25,914+-24 1039+-23,887+-23",
     "cutoff": 1766, "primary": "210+-25,914+-24 195+-25,61+-25 1039+-23,887+-23",
     "scan": "210+-25,914+-24 195+-25,61+-25 1039+-23,887+-23 1035+-23,68+-24",
     "scan_adaptive": "210+-25,914+-24 195+-25,61+-25 1039+-23,887+-23 1035+-23,67+-24"},      # testExampleScan.Adaptive, :178-189
    {"sample": "6bit/4color_ecc30_fountain_0.png", "source": "src/lib/extractor/test/ScannerTest.cpp:74-92, :164-176",
     "cutoff": 2268, "primary": "29+-27,29+-27 993+-27,29+-27 29+-27,993+-27",
     "scan": "29+-27,29+-27 993+-27,29+-27 29+-27,993+-27 993+-27,993+-27"},
    {"sample": "6bit/4_30_f2_734.jpg", "source": "src/lib/extractor/test/ScannerTest.cpp:94-112",
     "cutoff": 1575, "primary": "56+-24,166+-25 870+-24,133+-25 137+-20,910+-18",
     "scan": "56+-24,166+-25 870+-24,133+-25 137+-20,910+-18 837+-18,897+-18"},
    {"sample": "6bit/4_30_f2_246.jpg", "source": "src/lib/extractor/test/ScannerTest.cpp:114-132",
     "cutoff": 1606, "primary": "189+-25,899+-23 157+-26,79+-25 924+-18,811+-19",
     "scan": "189+-25,899+-23 157+-26,79+-25 924+-18,811+-19 911+-18,122+-19"},
    {"sample": "6bit/4_30_f1_360.jpg", "source": "src/lib/extractor/test/ScannerTest.cpp:149-162",
     "scan": "41+-24,196+-25 909+-25,183+-25 69+-23,1036+-23 896+-23,1036+-23"},
]
# ScannerTest/testSortTopToBottom(.2, .3): Anchor(x, xmax, y, ymax) lists in, strings out
SORT_GOLDENS = [
    {"in": [[300, 360, 100, 160], [300, 360, 300, 360], [100, 160, 300, 360]], "out": "330+-30,330+-30 130+-30,330+-30 330+-30,130+-30",
     "source": "src/lib/extractor/test/ScannerTest.cpp:208-221"},
    {"in": [[966, 1020, 966, 1020], [966, 1020, 2, 56], [2, 56, 966, 1020]], "out": "993+-27,993+-27 29+-27,993+-27 993+-27,29+-27",
     "source": "src/lib/extractor/test/ScannerTest.cpp:223-236"},
    {"in": [[383, 437, 994, 1048], [395, 447, 107, 157], [1250, 1296, 124, 170]], "out": "421+-26,132+-25 1273+-23,147+-23 410+-27,1021+-27",
     "source": "src/lib/extractor/test/ScannerTest.cpp:238-251"},
]
# color_correctionTest/testTransform (src/lib/chromatic_adaptation/test/color_correctionTest.cpp:14-30)
ADAPTATION_GOLDEN = {"actual": [192, 255, 255], "desired": [255, 255, 255],
                     "matrix_str": "[1.0655777, 0.2109226, -0.013239831;\n 0.023168325, 0.98723376, -0.0046780901;\n 0, 0, 1]",
                     "transform_in": [180, 98, 255], "transform_out": [209.09822971, 99.72629027, 255.0],
                     "source": "src/lib/chromatic_adaptation/test/color_correctionTest.cpp:14-30"}


def load_rgb(path):
    img = cv2.imread(path, cv2.IMREAD_COLOR)  # TestHelpers.h:8-18: imread + BGR2RGB
    return np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    manifest = {"cv2": cv2.__version__, "samples": {}, "goldens": [], "cv_pins": {}, "ccm_goldens": CCM_GOLDENS,
                "adaptation_golden": ADAPTATION_GOLDEN, "moore_penrose_goldens": MOORE_PENROSE_GOLDENS,
                "init_ccm_goldens": INIT_CCM_GOLDENS, "scan_goldens": SCAN_GOLDENS, "sort_goldens": SORT_GOLDENS}
    for s in SAMPLES:
        dst = os.path.join(HERE, s.replace("/", "__"))
        shutil.copyfile(os.path.join(REF, "samples", s), dst)
        os.chmod(dst, 0o644)
        rgb = load_rgb(dst)
        manifest["samples"][s] = {"file": os.path.basename(dst), "shape": list(rgb.shape), "rgb_sha256": sha(rgb)}
    # scan2434.jpg (1280x960, not extracted): only its size and its top-left 8x8 patch matter
    # (CimbReader not good -> colour read at (0,0), Decoder.h:107-114); keep just that.
    scan = load_rgb(os.path.join(REF, "samples/b/scan2434.jpg"))
    np.save(os.path.join(HERE, "b__scan2434_topleft8.npy"), scan[:8, :8].copy())
    manifest["samples"]["b/scan2434.jpg"] = {"file": "b__scan2434_topleft8.npy", "shape": list(scan.shape), "patch": "top-left 8x8 only"}
    for g in GOLDENS:
        manifest["goldens"].append(dict(zip(("sample", "mode", "ecc", "bytes", "sha256", "source"), g)))
    # OpenCV arithmetic pins on a camera frame (clean frames do not discriminate rounding variants)
    for s in ("b/ex2434.jpg", "6bit/4_30_f0_627_extract.jpg"):
        rgb = load_rgb(os.path.join(REF, "samples", s))
        gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        thr5 = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, 5, 0)
        k = np.array([[-0, -1, -0], [-1, 4.5, -1], [-0, -1, -0]], dtype=np.float32)
        sharp = cv2.filter2D(gray, -1, k)
        thr7 = cv2.adaptiveThreshold(sharp, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, 7, 0)
        manifest["cv_pins"][s] = {"gray": sha(gray), "thr5": sha(thr5), "sharp": sha(sharp), "thr7_sharp": sha(thr7)}
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(manifest["samples"]), "samples")


if __name__ == "__main__":
    main()
