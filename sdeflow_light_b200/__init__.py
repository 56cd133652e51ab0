"""sdeflow_light_b200 -- B200-native (sm_100a) drop-in for the hot path of vressegu/sdeflow-light (MSGM).

Mirrors the reference's Python surface:  ``SDEs`` (SGMsde, MSGMsde, forward_SDE, PluginReverseSDE),
``sde_scheme`` (EM / Heun / RK4-Stratonovich samplers) and ``NN`` (MLP score net); all heavy lifting is done by
hand-written CUDA kernels in ``csrc/`` behind the C ABI of ``include/msgm_b200.h``.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from . import sde_scheme, SDEs, NN, NNUnet1D, NNUnet  # noqa: F401
from .sde_scheme import euler_maruyama_sampler, heun_sampler, rk4_stratonovich_sampler  # noqa: F401
from .SDEs import SGMsde, MSGMsde, forward_SDE, PluginReverseSDE  # noqa: F401
from .NN import MLP  # noqa: F401
from .NNUnet1D import UNet1D  # noqa: F401
from .NNUnet import VorticityUNet  # noqa: F401

__version__ = "0.1.0"
