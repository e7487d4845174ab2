"""vorta_b200 — B200-native (sm_100a) implementation of VORTA's routed sparse attention hot path.

Host side mirrors the reference's interface for this path (``vorta/attention``, ``vorta/patch/router.py``,
``vorta/patch/utils.py``, ``vorta/ulysses``); all compute goes through the C ABI in ``include/vorta_b200.h``.
"""
__version__ = "0.1.0"
