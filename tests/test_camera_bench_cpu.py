"""bench.py --camera's CPU leg (libcimbar_b200/camera_bench.py cpu_pipeline): the same per-picture pipeline on one host core -- the
restatement's scan, cv2's deskew, the oracle's decode with should_preprocess = true -- runs without a GPU and decodes the photographs."""
import libcimbar_b200.camera_bench as cam


def test_cpu_pipeline_of_the_camera_bench():
    pics = cam.load_pictures()
    assert len(pics) == 2 and pics[0].shape == (960, 1280, 3)
    out = cam.cpu_pipeline(pics, reps=2)
    assert out["cores"] == 1 and out["kind"] == "port" and out["value"] > 0
    assert set(out["stages_ms_per_picture"]) == {"scan", "deskew_cv2", "decode"}
    assert out["good_bytes_per_picture"] == 7500.0          # mode 4C: all ten chunks of both photographs survive RS
