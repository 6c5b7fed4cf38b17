"""Fixed-point additive secret sharing (2 workers + crypto provider) for the encrypted demo.

Re-creates the PySyft contracts the reference relies on ([EXTERNAL]):
``fix_precision(precision_fractional=3)`` = ``round(x * 10**3)`` in the int64 ring;
``.share(w1, w2, crypto_provider)`` = additive sharing mod 2**64; multiplications consume Beaver
triples dealt by the crypto provider (SPDZ); ``.refresh()`` re-randomises shares;
``.get().float_precision()`` reconstructs and decodes.

Simplification (documented, off the metric path): the reference's ReLU uses SecureNN
comparisons; here the sign bit comes from a helper-aided comparison in which the two workers
blind the value with a common positive random scalar before the crypto provider sees it (the
helper learns the sign and a scaled magnitude, nothing about the scale).  Semi-honest model.
"""
from .sharing import (CryptoProvider, SharedTensor, fix_precision, float_precision, share, BASE,  # noqa: F401
                      PRECISION_FRACTIONAL)
from .mlp import SharedMLP, encrypted_sgd_step  # noqa: F401
from .dist import DistShared, PartyContext, train_encrypted_dist  # noqa: F401,E402
