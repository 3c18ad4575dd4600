"""B200-native implementation of the RRDB hot path of SamSweere/xmm-superres-denoise.

Same class / function names as the reference package ``xmm_superres_denoise`` for the path in
scope: ``models.GeneratorRRDB_SR`` / ``GeneratorRRDB_DN``,
``transforms.Normalize`` / ``ImageUpsample``, ``utils.loss_functions.create_loss``,
``metrics.PoissonNLLLoss``.  Everything executes in ``libxmm_b200.so`` (hand-written sm_100a
CUDA behind a C ABI, ``include/xmm_b200.h``); there is no CPU, eager-PyTorch or cuDNN fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
