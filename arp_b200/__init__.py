"""arp_b200 — B200-native (sm_100a) reward labeling for ARP-DT.

Drop-in for the reference's `arp_dt.label_reward` hot path: hand-written CUDA behind a C ABI
(include/arp_b200.h), bound with ctypes (arp_b200.capi), driven by `arp_b200.label_reward`.
"""
__version__ = "0.1.0"
