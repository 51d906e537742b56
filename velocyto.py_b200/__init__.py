"""velocyto.py_b200 -- B200-native (sm_100a) numerical core behind the VelocytoLoom hot path.

Drop-in for the numerical layer of velocyto.py 0.17.16 that sits under
``VelocytoLoom.knn_imputation / fit_gammas / predict_U / estimate_transition_prob``:

* ``estimation``  -- ``colDeltaCor*`` and ``fit_slope*`` with the reference's names and signatures
                     (velocyto/estimation.py), executed by hand-written CUDA kernels.
* ``neighbors``   -- ``connectivity_to_weights`` / ``convolve_by_sparse_weights`` (velocyto/neighbors.py).
* ``analysis``    -- ``VelocytoLoom`` with the hot methods of velocyto/analysis.py, data resident in HBM.
* ``device``      -- the cell-major fp32 device containers the kernels work on.
* ``_cabi``       -- ctypes binding of ``libvelo_b200.so`` (``include/velo_b200.h``).

Host code is Python; torch tensors are used only as device-memory containers and for
``torch.distributed``.  There is no CPU fallback: without the built library and a B200 the
compute entry points raise ``VeloError``.
"""
from . import _cabi
from ._cabi import VeloError

__all__ = ["_cabi", "VeloError"]
__version__ = "0.1.0"
