"""Loads tests/golden/ref_llvm.npz (outputs of the unmodified reference, see make_golden.py)."""
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_llvm.npz")
_cache = None


def load():
    global _cache
    if _cache is None:
        z = np.load(_PATH)
        digests = dict(zip(z["digest_keys"].tolist(), z["digest_vals"].tolist()))
        arrays = {k: z[k] for k in z.files if k not in ("digest_keys", "digest_vals")}
        _cache = (digests, arrays)
    return _cache
