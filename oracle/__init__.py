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
* ``color.py``      the dual-fisheye input colour pipeline (DF:568-725, pinned bit-exactly) and the
                    formula of the cutter's video colour filter (PC:299-309).
* ``footprint.py``  distinct source pixels touched per view set (the algorithmic bytes of the roofline).

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
* ERP maps: PINNED against the reference's in-repo geometry -- ``direction_from_uv``
  / ``rotate_pitch`` / ``rotate_yaw`` / ``lonlat_to_xy`` (gs360_GUI.py:342-424) are
  taken out of gs360_GUI.py's syntax tree (the module itself needs tkinter) and run
  by ``tests/golden/make_golden.py``; ``tests/golden/gui_geometry.npz`` holds their
  outputs for ten views (presets, seam, poles) and both the oracle and the kernels
  are compared with it.  What stays UNPINNED is the boundary to FFmpeg's ``v360``
  filter itself (third party, no pinned version, binary absent from this image,
  SURVEY.md section 8c): its pixel-centre convention (available as ``convention="v360"``)
  and its own cubic kernel cannot be checked here.
* video colour step (``color.py``, formula level) and the v360 fisheye inputs /
  outputs follow published formulas; same ffmpeg caveat.
* ``remap_c.c`` is not written: the NumPy sampler finishes every test size in seconds.
"""
