"""Oracle for the heat-map read-out of df2d: hard arg-max + peak value (reference README.md:404;
the golden 2-D points sit exactly on the k/64, k/128 grid -- SURVEY.md Appendix A).

Test infrastructure (see ``oracle/__init__.py``).  First occurrence wins on ties, like
``numpy.argmax`` / ``torch.argmax`` on the flattened (row-major) map.
"""
import numpy as np


def heatmap_argmax(hm):
    """hm (B,K,H,W) -> idx (B,K) int32 flat index, conf (B,K) float32 peak."""
    hm = np.asarray(hm, dtype=np.float32)
    B, K, H, W = hm.shape
    flat = hm.reshape(B, K, H * W)
    idx = flat.argmax(axis=-1)
    conf = np.take_along_axis(flat, idx[..., None], axis=-1)[..., 0]
    return idx.astype(np.int32), conf.astype(np.float32)
