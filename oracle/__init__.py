"""oracle/ -- CPU restatement of the reference's per-frame vision hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or the timed CPU
baseline; the product package laser_3d_reconstruction_b200 never does and fails loudly when its
CUDA library is missing.

* ``oracle/csrc``  plain-C restatement of the OpenCV arithmetic the reference calls
  (pinned bit-exactly against cv2 4.13.0 of this image for every integer stage;
  WLS: PARITY UNPINNED, cv2.ximgproc is not installed).
* ``oracle/ref_ops.py``  restatement of the reference's Python classes (extractors,
  reconstructors, get_frames depth path), pinned by tests/golden fixtures generated from
  /root/reference by tests/golden/make_golden.py.
"""
