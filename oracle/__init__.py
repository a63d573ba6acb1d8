"""CPU oracle for the panorama -> perspective remap hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU arm -- never as a fallback for the CUDA library.

What is restated here (reference = /root/reference, cited file:line):

* ``geometry.py``   float64 ray -> rotation -> lon/lat -> ERP pixel maps
                    (gs360_GUI.py:342-395, :419-424) and the equisolid + Brown
                    dual-fisheye maps
                    (cli_tools/gs360_DualFisheyeDistortionCalibration.py:975-1005,
                    :1759-1823, :1857-1907).
* ``sampler.py``    the arithmetic of ``cv2.remap`` -- the third-party routine
                    the reference calls at
                    gs360_DualFisheyeDistortionCalibration.py:2001-2008 and
                    :2031-2038 (opencv-python, requirements.txt; 4.13.0 in this
                    image): 1/32-px fraction quantisation, a=-0.75 cubic,
                    15-bit fixed-point u8 tables, per-tap constant border.
* ``remap_c.c``     the same sampler in plain C (fast enough for full frames).

Pinning status
--------------
* sampler: PINNED -- bit-exact against ``cv2.remap`` itself (which is present in
  this image and on the GPU box) and against committed fixtures in
  ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
* dual-fisheye maps: PINNED against maps produced by importing the reference's
  own ``build_direct_perspective_map_for_lens`` / ``build_perspective_spec_maps``
  (fixtures in ``tests/golden/``).
* view planner: PINNED against ``build_view_jobs`` outputs of the reference
  (fixtures in ``tests/golden/``).
* ERP maps: PARITY UNPINNED at the ffmpeg boundary.  The reference's ERP
  arithmetic lives in FFmpeg's ``v360`` filter (no pinned version, binary absent
  from this image); the oracle follows the in-repo
  float64 statement of the same geometry (gs360_GUI.py:377-424), and that code
  is pinned only through the shared rotation helper that the dual-fisheye
  fixtures exercise.
"""
