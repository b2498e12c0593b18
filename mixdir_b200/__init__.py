"""mixdir_b200: B200-native engine for the variational-inference hot path of const-ae/mixdir.

The product is the C-ABI CUDA library (include/mixdir_b200.h, mixdir_b200/csrc/).  This
package is the host-side mirror of the reference's R interface for that path.
"""
from .api import find_defining_features, mixdir, predict  # noqa: F401
from .coding import code_columns, recode_for_predict  # noqa: F401
from .engine import Comm, Engine, MixdirError, ZetaUnderflow  # noqa: F401
