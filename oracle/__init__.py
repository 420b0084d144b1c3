"""CPU oracle for the df2d + pyba hot path of DeepFly3D.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepfly3d_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline, never as the product.

The arithmetic of the path lives in two third-party packages that are *not*
vendored in the reference checkout (``nely-df2d >= 0.14`` and
``nely-pyba >= 0.13``, reference ``setup.py:30-31``; neither is version-pinned).
The oracle therefore restates their observable behaviour:

* 3-D half (``oracle.geometry``, ``oracle.procrustes``, ``oracle.pack``):
  **parity pinned** -- reproduces the reference's own golden pickle
  ``tests/data/reference_df3d/df3d_result_3d.pkl`` from ``df3d_result_2d.pkl``
  + ``data/calib.pkl`` at the tolerances of the reference's ``test_calibration``
  (``tests/test_df3d.py:198-244``); see ``tests/test_oracle_golden.py``.
* image ingest (``oracle.ingest``): the loader's resize, **pinned** bit for bit against
  ``cv2.resize(..., INTER_LINEAR)`` itself (``tests/test_oracle_ingest.py``) -- the definition
  this port's host loader has always used; df2d's own resize mode is not verifiable offline.
* 2-D half (``oracle.hourglass``, ``oracle.argmax``): **parity unpinned** -- the
  golden 2-D points need the pretrained ``sh8_deepfly.tar`` weights that df2d
  downloads at run time (reference ``df3d/config.py:30-32``); they are not in
  the checkout and there is no network.  The architecture is the published
  stacked hourglass (Newell et al. 2016, pre-activation bottlenecks) with the
  hyper-parameters hinted by ``df3d/config.py:18,33-36``; the read-out rule
  (hard argmax + peak value) is pinned by the golden 2-D grid (README.md:404).
"""
